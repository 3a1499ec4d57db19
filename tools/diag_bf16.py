#!/usr/bin/env python
"""Diagnostic: train-mode forward of a net in bf16 / fp32 on the device against the CPU oracle, for a given size."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pose_oracle as po
import __graft_entry__ as ge
from conftest import rel_err

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="fusionnet"); ap.add_argument("--model", default="resnet50")
ap.add_argument("--side", type=int, default=128); ap.add_argument("--n", type=int, default=8)
ap.add_argument("--fp32", action="store_true"); ap.add_argument("--eval", action="store_true")
a = ap.parse_args()
b2 = ge.load_package()
dev = torch.device("cuda:0")
fused = "fusion" in a.kind
cfg = po.net_config(side_in=a.side, num_joints=17, depth_only=not fused)
sd = po.init_state(a.kind, a.model, cfg, seed=41)
for k, v in sd.items():
    if v.dim() == 4:
        sd[k] = v.bfloat16().float()
color, depth, tc, tv = po.synth_batch(a.n, a.side, 17, seed=9)
color, depth = color.bfloat16().float(), depth.bfloat16().float()
with torch.no_grad():
    zr, lr = po.net_forward({k: v.clone() for k, v in sd.items()}, a.kind, a.model, cfg, color if fused else depth,
                            depth if fused else None, training=not a.eval)
net = getattr(getattr(b2, a.kind), a.model)(cfg, False)
net.load_state_dict(sd)
net = net.to(dev)
net = net.eval() if a.eval else net.train()
if not a.fp32:
    net = net.half()
with torch.no_grad():
    z, l = net(color.to(dev), depth.to(dev)) if fused else net(depth.to(dev))
torch.cuda.synchronize()
env = {k: v for k, v in os.environ.items() if k.startswith("B2POSE")}
print("%s %s n%d s%d %s %s env=%s: rel err z %.4f last %.4f" % (a.kind, a.model, a.n, a.side, "fp32" if a.fp32 else "bf16",
      "eval" if a.eval else "train", env, rel_err(z, zr), rel_err(l, lr)), flush=True)
