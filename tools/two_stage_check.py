"""Single-GPU check of the staged backward pass (Trainer._fwd_bwd_deep / _bwd_shallow): the gradients of one step
must equal those of the plain backward pass.  Run with B2POSE_BN_TOTALS=0 for bit-reproducible BatchNorm sums.

    B2POSE_BN_TOTALS=0 python tools/two_stage_check.py [resnet18|resnet50] [side] [batch]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge   # noqa: E402
import pose_oracle as po       # noqa: E402

b2 = ge.load_package()
model = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
side = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)


def grads(half, two, kind):
    cfg = po.net_config(side_in=side, num_joints=17, depth_only=kind != "partial_fusionnet")
    net = getattr(getattr(b2, kind), model)(cfg, False)
    net.load_state_dict(po.init_state(kind, model, cfg, seed=5))
    args = b2.train_args(model=model, num_joints=17, side_in=side, depth_only=kind != "partial_fusionnet",
                         do_fusion=kind == "partial_fusionnet", half_acc=half)
    tr = b2.Trainer(args, net.to(dev).train(), dict(key_index=16), use_graph=False)
    tr._force_two = two
    out = tr.train_step(tuple(t.to(dev) for t in po.synth_batch(n, side, 17, seed=20)))
    torch.cuda.synchronize()
    return tr, float(out["loss"]), tr.flat.g.clone()


for kind in ("partial_fusionnet", "partial_depthnet"):
    for half in (False, True):
        tr, l0, g0 = grads(half, False, kind)
        _, l0b, g0b = grads(half, False, kind)
        _, l1, g1 = grads(half, True, kind)
        rel = lambda a, b: float((a - b).norm() / b.norm())
        print("%s %s %dx%d batch %d half=%s: loss %.6f / %.6f / %.6f   plain-vs-plain %.2e   staged-vs-plain %.2e"
              % (kind, model, side, side, n, half, l0, l0b, l1, rel(g0b, g0), rel(g1, g0)))
        worst = []
        for name, p, o in zip(tr.list_names, tr.flat.params, tr.flat.offsets):
            a, b = g1[o:o + p.numel()], g0[o:o + p.numel()]
            worst.append((float((a - b).norm() / (b.norm() + 1e-30)), name))
        worst.sort(reverse=True)
        print("   worst parameters:", ", ".join("%s %.1e" % (nm, e) for e, nm in worst[:6]))
