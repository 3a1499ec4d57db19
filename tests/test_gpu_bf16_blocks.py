"""Block-level parity of the benched bf16 training path at the contract tolerance, forward AND backward.

Training-mode BatchNorm makes the whole ResNet-50 an expanding map (tests/golden/bf16_sensitivity.npz: the reference's
own bf16-autocast forward is 0.1-0.6 relative away from its fp32 forward), so end-to-end elementwise bounds cannot
separate a kernel bug from amplified rounding.  This test removes the amplification instead of loosening the bound:
the oracle (CPU, reference algorithm under the bf16 storage contract, ``po.round_bf16``) runs one forward + backward of
the whole net and records, for every stem, residual block, the fusion unit and the regressor, its input, its output
and their gradients; every unit of the device net -- same module objects, same fused conv+BN(+ReLU)(+residual)(x veil)
autograd nodes, tcgen05 kernels, BatchNorm totals path, side streams that the benched step uses -- is then fed the
ORACLE's input and output gradient and must reproduce

  * its output within 2e-2 relative (max|a-b| / max|b|, the bf16 bound of north_star; measured <= 0.7e-2),
  * the updated veil bit-exactly,
  * its input gradient within 5e-2 in relative L2 norm (measured <= 2.9e-2: the L2 norm is dominated by the outliers below),
    with at least 99.9 % of the elements within 3e-2 of max|ref| (measured: 99.9 % quantile <= 2.8e-2).  The remaining <= 1e-4 of the elements are ReLU-gate flips: a
    pre-activation within rounding distance of zero is gated differently on the two sides, which switches one whole
    term of that pixel's sum on or off (the device and the oracle share every formula, not the summation order),
  * every parameter gradient within 3e-2 in norm (measured <= 0.3e-2) and 8e-2 in relative L2 norm (measured <= 5.2e-2,
    the stems' BatchNorm bias at 256x256: ties of equal bf16 values in the 3x3 max-pool route gradients differently).

Reference chain: partial_depthnet.py:140-157 / fusionnet.py:97-127 (blocks), :213-229 (stems), fusionnet.py:130-140
(fusion), depth_train.py:397-405 (head + loss, tests/test_gpu_head.py).
"""
import numpy as np
import pytest
import torch

import pose_oracle as po
from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL, TOL_GRAD, TOL_L2, TOL_DX = 2e-2, 3e-2, 8e-2, 5e-2


def _fused(kind):
    return kind in ("fusionnet", "partial_fusionnet")


def _nhwc(t, dev):
    """Oracle NCHW fp32 tensor -> device NHWC bf16 (the values are bf16-exact under the storage contract)."""
    return t.detach().permute(0, 2, 3, 1).contiguous().to(dev).bfloat16()


def _oracle_trace(kind, model, side, n, seed_w, seed_b):
    cfg = po.net_config(side_in=side, num_joints=17, depth_only=not _fused(kind))
    sd = po.init_state(kind, model, cfg, seed=seed_w)
    for k, v in sd.items():
        if v.dim() == 4:
            sd[k] = v.bfloat16().float()
    color, depth, true_cam, true_val = po.synth_batch(n, side, 17, seed=seed_b, invalid_frac=0.25)
    batch = (color.bfloat16().float(), depth.bfloat16().float(), true_cam, true_val)
    orc = po.StepOracle({k: v.clone() for k, v in sd.items()}, kind, model, cfg, key_index=16, act_round=po.round_bf16)
    orc.trace = []
    loss, spec, z = orc.forward_loss(batch)
    loss.backward()                       # no optimizer step: parameter gradients stay unclipped
    return cfg, sd, batch, orc


def _check_params(unit_name, module, orc, prefix, worst):
    for pname, p in module.named_parameters():
        want = orc.sd["%s.%s" % (prefix, pname) if prefix else pname].grad
        got = p.grad
        assert got is not None, (unit_name, pname)
        got = got.detach().float().cpu()
        nerr = abs(float(got.norm()) - float(want.norm())) / max(float(want.norm()), 1e-12)
        l2 = float((got - want).norm()) / max(float(want.norm()), 1e-12)
        worst["gnorm"] = max(worst["gnorm"], nerr)
        worst["gelem"] = max(worst["gelem"], l2)
        assert nerr < TOL_GRAD and l2 < TOL_L2, (unit_name, pname, nerr, l2)


def _check_dx(unit_name, got_nhwc, want, worst):
    got = got_nhwc.permute(0, 3, 1, 2).float().cpu()
    want = want.detach()
    d = (got - want).abs()
    l2 = float((got - want).norm()) / max(float(want.norm()), 1e-30)
    frac = float((d > TOL_GRAD * float(want.abs().max())).float().mean())
    worst["dx"] = max(worst["dx"], l2)
    worst["dx_out"] = max(worst["dx_out"], frac)
    assert l2 < TOL_DX and frac < 1e-3, (unit_name, "dx", l2, frac)


def _run_units(b2pose, dev, kind, model, side, n, seed_w=41, seed_b=9):
    cfg, sd, batch, orc = _oracle_trace(kind, model, side, n, seed_w, seed_b)
    net = getattr(getattr(b2pose, kind), model)(cfg, False)
    net.load_state_dict(sd)
    net = net.to(dev).train().half()
    worst = dict(out=0.0, dx=0.0, dx_out=0.0, gnorm=0.0, gelem=0.0)
    for rec in orc.trace:
        name = rec["name"]
        net.zero_grad(set_to_none=True)
        dout = _nhwc(rec["out"].grad, dev)
        if name in ("conv1", "conv2"):                                  # stems: conv + BN + ReLU + max-pool (+ veil)
            conv, bn = getattr(net, name), getattr(net, "bn" + name[-1])
            out, vout = net._stem(rec["x"].to(dev), conv, bn)
            out.backward(dout)
            _check_params(name, conv, orc, name, worst)
            _check_params(name, bn, orc, "bn" + name[-1], worst)
            dx = None
        elif name == "fusion":
            a, b = _nhwc(rec["x"], dev).requires_grad_(), _nhwc(rec["x2"], dev).requires_grad_()
            out, vout = net.fusion.forward_nhwc(a, b), None
            out.backward(dout)
            _check_params(name, net.fusion, orc, "fusion", worst)
            _check_dx(name, a.grad, rec["x"].grad, worst)
            _check_dx(name, b.grad, rec["x2"].grad, worst)
            dx = None
        elif name == "regressor":
            x = _nhwc(rec["x"], dev).requires_grad_()
            out, vout = net.regressor.forward_nhwc(x), None
            out.backward(dout)
            _check_params(name, net.regressor, orc, "regressor", worst)
            dx = x.grad
        else:                                                            # residual block "layerL.i"
            lname, idx = name.split(".")
            blk = getattr(net, lname)[int(idx)]
            x = _nhwc(rec["x"], dev).requires_grad_()
            veil = None if rec["veil"] is None else rec["veil"][:, 0].contiguous().to(dev)
            out, vout = blk.forward_nhwc(x, veil)
            out.backward(dout)
            _check_params(name, blk, orc, name, worst)
            dx = x.grad
        e = rel_err(out.permute(0, 3, 1, 2), rec["out"])
        worst["out"] = max(worst["out"], e)
        assert e < TOL, (name, "out", e)
        if rec["veil_out"] is not None:
            assert vout is not None and torch.equal(vout.cpu(), rec["veil_out"][:, 0]), (name, "veil")
        if dx is not None:
            _check_dx(name, dx, rec["x"].grad, worst)
    print("%s %s %dx%d batch %d: %d units; worst  out (max-rel) %.4f  dx (L2-rel) %.4f, gate-flip outliers %.1e  "
          "param-grad norm %.4f  param-grad L2-rel %.4f" % (kind, model, side, side, n, len(orc.trace), worst["out"],
                                                          worst["dx"], worst["dx_out"], worst["gnorm"], worst["gelem"]))
    return worst


@pytest.mark.parametrize("kind", ["fusionnet", "partial_fusionnet", "partial_depthnet"])
def test_bf16_units_resnet50_128(b2pose, dev, kind):
    _run_units(b2pose, dev, kind, "resnet50", 128, 8)


def test_bf16_units_resnet50_full_size(b2pose, dev):
    """The benched workload's shapes: 256x256, batch 16 (the oracle's forward + backward takes ~30 s of CPU)."""
    _run_units(b2pose, dev, "partial_fusionnet", "resnet50", 256, 16, seed_w=43, seed_b=11)


def test_bf16_units_resnet18_odd_size(b2pose, dev):
    """BasicBlock nets (un-premasked 3x3 PartialConv first in every block) at the reference's default 257-style odd
    size (65x65: ragged tiles everywhere)."""
    _run_units(b2pose, dev, "partial_fusionnet", "resnet18", 65, 4, seed_w=7, seed_b=5)
    _run_units(b2pose, dev, "partial_depthnet", "resnet18", 65, 4, seed_w=7, seed_b=5)
