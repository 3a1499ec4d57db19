"""Where do the ATen fill kernels of a training step come from?  One eager step under torch.profiler with Python
stacks; prints the call sites of aten::fill_ / aten::zero_ / aten::zeros / aten::ones events and their sizes."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge   # noqa: E402
import pose_oracle as po       # noqa: E402

b2 = ge.load_package()
dev = torch.device("cuda", 0)
cfg = po.net_config(side_in=128, num_joints=17, depth_only=False)
net = b2.partial_fusionnet.resnet50(cfg, False)
args = b2.train_args(model="resnet50", num_joints=17, side_in=128, depth_only=False, do_fusion=True, half_acc=True)
tr = b2.Trainer(args, net.to(dev).train(), dict(key_index=16), use_graph=False)
batch = tuple(t.to(dev) for t in po.synth_batch(8, 128, 17, seed=3))
for _ in range(2):
    tr.train_step(batch)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
    tr.train_step(batch)
    torch.cuda.synchronize()
sites = collections.Counter()
for ev in prof.events():
    if ev.name in ("aten::fill_", "aten::zero_", "aten::zeros", "aten::ones", "aten::zeros_like", "aten::cat", "aten::add",
                   "aten::copy_", "aten::mul"):
        stack = [s for s in ev.stack if "b200" in s or "autograd" in s][:3]
        sites[(ev.name, str(ev.input_shapes)[:60], " <- ".join(s.split("/")[-1] for s in stack))] += 1
for (name, shapes, where), n in sorted(sites.items(), key=lambda x: -x[1]):
    print("%3d  %-16s %-60s %s" % (n, name, shapes, where))
