#!/usr/bin/env python
"""Runs ONE training step of the bench workload between cudaProfilerStart/Stop (after eager
warm-up steps), for `ncu --profile-from-start off ...`.  Not a benchmark: numbers printed under
a profiler are never bench values."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="fusionnet")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--side", type=int, default=256)
    ap.add_argument("--joints", type=int, default=17)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--steps", type=int, default=1)
    a = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    b2 = ge.load_package()
    dev = torch.device("cuda:0")
    fused = "fusion" in a.workload
    cfg = b2.train_args(model="resnet50", num_joints=a.joints, side_in=a.side, depth_only=not fused, do_fusion=fused,
                        half_acc=a.dtype == "bf16")
    torch.manual_seed(0)
    net = getattr(getattr(b2, a.workload), "resnet50")(cfg, False).to(dev).train()
    tr = b2.Trainer(cfg, net, dict(key_index=a.joints - 1), use_graph=False)
    batch = b2.synthetic_batch(a.batch, a.side, a.joints, dev, seed=1)
    for _ in range(2):
        tr.train_step(batch)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(a.steps):
        out = tr.train_step(batch)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("loss", float(out["loss"]))


if __name__ == "__main__":
    main()
