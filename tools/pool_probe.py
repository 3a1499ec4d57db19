"""CUDA-event timing of the 3x3 s2 max-pool forward / backward at the stems' shape (N64 128x128x64 bf16)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge   # noqa: E402

b2 = ge.load_package()
dev = torch.device("cuda", 0)
x = torch.randn(64, 128, 128, 64, device=dev).bfloat16().requires_grad_(True)
veil = (torch.rand(64, 128, 128, device=dev) > 0.3).float()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
y, v = b2.ops.MaxPoolFn.apply(x, veil)
g = torch.randn_like(y)
ts = {"fwd": [], "bwd": []}
for it in range(7):
    flush.fill_(it)
    e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    e0.record()
    y, v = b2.ops.MaxPoolFn.apply(x, veil)
    e1.record()
    flush.fill_(it + 1)
    e2.record()
    y.backward(g)
    e3.record()
    torch.cuda.synchronize()
    ts["fwd"].append(e0.elapsed_time(e1) * 1e3)
    ts["bwd"].append(e2.elapsed_time(e3) * 1e3)
    x.grad = None
for k, t in ts.items():
    t.sort()
    print("maxpool %s: %.1f us (median of %d)" % (k, t[len(t) // 2], len(t)))
