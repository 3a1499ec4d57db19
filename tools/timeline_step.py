#!/usr/bin/env python
"""Kernel timeline of graph-replayed training steps (torch.profiler / CUPTI): how much of the step is any kernel
running, how much do the streams overlap, where are the gaps.  Diagnostic only (profiler overhead included)."""
import argparse, collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="partial_fusionnet")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.json"))
a = ap.parse_args()
b2 = ge.load_package()
dev = torch.device("cuda:0")
fused = "fusion" in a.workload
cfg = b2.train_args(model="resnet50", num_joints=17, side_in=256, depth_only=not fused, do_fusion=fused, half_acc=True)
torch.manual_seed(0)
net = getattr(getattr(b2, a.workload), "resnet50")(cfg, False).to(dev).train()
tr = b2.Trainer(cfg, net, dict(key_index=16), use_graph=True)
batch = b2.synthetic_batch(a.batch, 256, 17, dev, seed=1)
for _ in range(8):
    tr.train_step(batch)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        tr.train_step(batch)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ks = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "stream", None)) for e in ev), key=lambda t: t[0])
print("kernels recorded:", len(ks))
if not ks:
    sys.exit(0)
t0, t1 = ks[0][0], max(k[1] for k in ks)
span = (t1 - t0) / a.steps
# union busy time
busy, cur_s, cur_e = 0.0, None, None
for s, e, _, _ in ks:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
total = sum(e - s for s, e, _, _ in ks)
print("per step: span %.3f ms, some kernel running %.3f ms (%.1f%%), sum of kernel durations %.3f ms (mean concurrency %.2f)"
      % (span / 1e3, busy / a.steps / 1e3, 100 * busy / (t1 - t0), total / a.steps / 1e3, total / busy))
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n, _ in ks:
    k = n.replace("void ", "").replace("(anonymous namespace)::", "").split("(")[0][:60]
    agg[k][0] += 1
    agg[k][1] += e - s
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print("  %-60s %5d launches/step %8.3f ms/step" % (k, v[0] // a.steps, v[1] / a.steps / 1e3))
# gap histogram (no kernel running)
gaps, cur_e = [], None
for s, e, _, _ in ks:
    if cur_e is not None and s > cur_e:
        gaps.append(s - cur_e)
    cur_e = e if cur_e is None else max(cur_e, e)
gaps.sort()
if gaps:
    print("idle gaps per step: %d, total %.3f ms; median %.2f us, p90 %.2f us, max %.1f us" % (
        len(gaps) // a.steps, sum(gaps) / a.steps / 1e3, gaps[len(gaps) // 2], gaps[int(len(gaps) * 0.9)], gaps[-1]))
json.dump([(s - t0, e - t0, n[:80], st) for s, e, n, st in ks], open(a.out, "w"))
