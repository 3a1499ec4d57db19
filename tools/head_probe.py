#!/usr/bin/env python
"""Head (softmax + soft-argmax, forward and backward) and depth unprojection at the bench shapes: CUDA-event timings when
run plain, the target of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` otherwise."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

b2 = ge.load_package()
dev = torch.device("cuda:0")
N, J, D, S = 64, 17, 16, 16
feat = (torch.randn(N, S, S, D * J, device=dev) * 3).bfloat16().permute(0, 3, 1, 2).requires_grad_()
img = torch.rand(N, 256, 256, device=dev) + 0.05
K = [[365.0, 0, 128.0], [0, 365.0, 128.0], [0, 0, 1]]
flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)


def timed(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.max()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


coords = b2.heatmap_coords(feat, D, J, 1000.0)
g = torch.randn_like(coords)
coords.backward(g, retain_graph=True)
b2.utils.to_depth if hasattr(b2, "utils") else None
out = b2.ops.unproject_depth(img, K)
torch.cuda.synchronize()
logit_bytes = feat.numel() * 2
print("head_fwd  N=%d J=%d D=%d %dx%d bf16: %.1f us  (%.2f MB of logits -> %.0f GB/s)" % (
    N, J, D, S, S, (t := timed(lambda: b2.heatmap_coords(feat, D, J, 1000.0))), logit_bytes / 1e6, logit_bytes / t / 1e3))
t = timed(lambda: coords.backward(g, retain_graph=True))
print("head_bwd  (read logits, write dlogits): %.1f us  (%.2f MB -> %.0f GB/s)" % (t, 2 * logit_bytes / 1e6, 2 * logit_bytes / t / 1e3))
t = timed(lambda: b2.ops.unproject_depth(img, K))
ub = img.numel() * 8
print("unproject N=%d 256x256 fp32 (read + write): %.1f us  (%.1f MB -> %.0f GB/s)" % (N, t, ub / 1e6, ub / t / 1e3))
big = torch.rand(1024, 256, 256, device=dev) + 0.05
t = timed(lambda: b2.ops.unproject_depth(big, K), reps=5)
print("unproject N=1024 256x256 fp32 (read + write): %.1f us  (%.1f MB -> %.0f GB/s)" % (t, big.numel() * 8 / 1e6, big.numel() * 8 / t / 1e3))
