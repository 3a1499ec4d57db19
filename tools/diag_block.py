#!/usr/bin/env python
"""Diagnostic: one residual block of the device net fed the oracle's input / output gradient (see tests/test_gpu_bf16_blocks.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pose_oracle as po
import __graft_entry__ as ge
import test_gpu_bf16_blocks as tb
from conftest import rel_err

b2 = ge.load_package()
dev = torch.device("cuda:0")
kind, model, side, n = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
names = sys.argv[5].split(",")
import numpy as np
def q(d, want, f):
    return float(np.quantile(d.flatten().numpy(), f)) / float(want.abs().max())
cfg, sd, batch, orc = tb._oracle_trace(kind, model, side, n, 41, 9)
net = getattr(getattr(b2, kind), model)(cfg, False)
net.load_state_dict(sd)
net = net.to(dev).train().half()
for rec in orc.trace:
    if names != ["all"] and rec["name"] not in names:
        continue
    if "." not in rec["name"]:
        continue
    lname, idx = rec["name"].split(".")
    blk = getattr(net, lname)[int(idx)]
    for mode in ("holder",):
        if mode == "no-holder":
            blk._holder = lambda x: None
        net.zero_grad(set_to_none=True)
        x = tb._nhwc(rec["x"], dev).requires_grad_()
        veil = None if rec["veil"] is None else rec["veil"][:, 0].contiguous().to(dev)
        out, vout = blk.forward_nhwc(x, veil)
        out.backward(tb._nhwc(rec["out"].grad, dev))
        got = x.grad.permute(0, 3, 1, 2).float().cpu()
        want = rec["x"].grad
        d = (got - want).abs()
        i = int(d.argmax())
        print(rec["name"], mode, "out err %.4f dx err %.4f | max|want| %.4g, worst at flat %d: got %.5g want %.5g | mean|d| %.3g mean|want| %.3g | frac elems with |d| > 0.05 max: %.2e"
              % (rel_err(out.permute(0, 3, 1, 2), rec["out"]), rel_err(got, want), float(want.abs().max()), i, float(got.flatten()[i]),
                 float(want.flatten()[i]), float(d.mean()), float(want.abs().mean()), float((d > 0.05 * want.abs().max()).float().mean())))
        print("    dx: L2-rel %.4f  q99.9 %.4f q99.99 %.4f" % (float((got - want).norm() / want.norm()), q(d, want, 0.999), q(d, want, 0.9999)))
        for pname, p in blk.named_parameters():
            w = orc.sd["%s.%s" % (rec["name"], pname)].grad
            g = p.grad.float().cpu()
            dd = (g - w).abs()
            print("    %-22s norm err %.4f elem err %.4f L2-rel %.4f q99.9 %.4f" % (pname, abs(float(g.norm()) - float(w.norm())) / float(w.norm()), rel_err(g, w), float((g - w).norm() / w.norm()), q(dd, w, 0.999)))
