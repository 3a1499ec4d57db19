"""GPU parity of the distillation ("privileged information") step -- SURVEY §8f rank 1: the fused
feature-mimic loss (Trainer.distill, depth_train.py:115-129), the attention map (utils.py:14-42) and the
whole distill_train step (depth_train.py:179-283) against fixtures produced by the reference itself."""
import numpy as np
import pytest
import torch

import pose_oracle as po
from conftest import rel_err

pytestmark = pytest.mark.gpu

MIMIC_KW = {"l2": dict(sigmoid=False, bin_dist=False), "sigmoid": dict(sigmoid=True, bin_dist=False),
            "bce": dict(sigmoid=False, bin_dist=True)}
DISTILL_CASES = {
    "dist_pf18_l2": ("partial_fusionnet", {}, dict(depth_only=False), dict()),
    "dist_f18skip_sig": ("fusionnet", dict(skip_relu=True, early_dist=True),
                         dict(depth_only=False, skip_relu=True, early_dist=True), dict(sigmoid=True)),
    "dist_pf18_bce_frz": ("partial_fusionnet", {}, dict(depth_only=False), dict(bin_dist=True, do_freeze=True)),
}


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_mimic_loss_golden(b2pose, dev, golden_dir, layout):
    g = np.load(golden_dir + "/distill.npz")
    for key in g["loss_names"]:
        mode = str(key).rsplit("_", 1)[1]
        t = torch.tensor(g[f"{key}_t"], device=dev)
        s = torch.tensor(g[f"{key}_s"], device=dev)
        if layout == "nhwc":
            t, s = t.contiguous(memory_format=torch.channels_last), s.contiguous(memory_format=torch.channels_last)
        s.requires_grad_(True)
        a = torch.tensor(g[f"{key}_a"], device=dev)
        loss = b2pose.mimic_loss(t, s, a, **MIMIC_KW[mode])
        (loss * 1.5).backward()
        np.testing.assert_allclose(float(loss), float(g[f"{key}_loss"]), rtol=1e-5)
        assert rel_err(s.grad, 1.5 * g[f"{key}_ds"]) < 1e-4, key
    with pytest.raises(ValueError):
        b2pose.mimic_loss(torch.zeros(2, 4, 3, 3, device=dev), torch.zeros(2, 4, 3, 2, device=dev),
                          torch.ones(2, 1, 3, 3, device=dev))


def test_mimic_loss_bf16_and_full_size(b2pose, dev):
    """bf16 features within 2e-2 of the fp32 oracle; full-size [64, 2048, 16, 16] properties: the loss of
    identical features is 0 with zero gradient, and scaling the attention map scales the L2 loss."""
    g = torch.Generator().manual_seed(2)
    t, s = torch.randn(4, 64, 9, 9, generator=g), torch.randn(4, 64, 9, 9, generator=g)
    a = torch.rand(4, 1, 9, 9, generator=g)
    for mode, kw in MIMIC_KW.items():
        sr = s.clone().requires_grad_(True)
        want = po.distill_loss(t.bfloat16().float(), sr, a, **kw)
        want.backward()
        sd = s.to(dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        got = b2pose.mimic_loss(t.to(dev).bfloat16().contiguous(memory_format=torch.channels_last), sd, a.to(dev), **kw)
        got.backward()
        assert abs(float(got) - float(want)) / abs(float(want)) < 2e-2, mode
        assert rel_err(sd.grad.float(), sr.grad) < 2e-2, mode
    N, C, H = 64, 2048, 16
    f = torch.randn(N, H, H, C, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2)
    att = torch.rand(N, 1, H, H, device=dev)
    fs = f.clone().requires_grad_(True)
    z = b2pose.mimic_loss(f, fs, att)
    z.backward()
    assert float(z) == 0.0 and float(fs.grad.abs().max()) == 0.0
    f2 = torch.randn_like(f)
    l1, l2 = float(b2pose.mimic_loss(f, f2, att)), float(b2pose.mimic_loss(f, f2, att * 2))
    assert abs(l2 / l1 - 2.0) < 1e-4
    ref = float(torch.linalg.norm(((f.float() - f2.float()) * att).reshape(N, -1), dim=-1).mean())
    assert abs(l1 - ref) / ref < 1e-4


def test_attention_map(b2pose, dev, golden_dir):
    g = np.load(golden_dir + "/distill.npz")
    for name in ("att_257", "att_64", "att_48s8"):
        side, stride = (int(v) for v in g[name + "_cfg"])
        got = b2pose.get_attention(side, stride, g[name + "_coords"], True)          # numpy in -> numpy out
        assert got.shape == g[name + "_map"].shape
        np.testing.assert_allclose(got, g[name + "_map"], rtol=2e-5, atol=1e-7)
        assert np.array_equal(b2pose.get_attention(side, stride, g[name + "_coords"], False), g[name + "_ones"])
        c = torch.tensor(np.stack([g[name + "_coords"]] * 3), dtype=torch.float32, device=dev)
        batched = b2pose.get_attention(side, stride, c, True)
        assert tuple(batched.shape) == (3, 1) + g[name + "_map"].shape[1:]
        np.testing.assert_allclose(batched[2, 0].cpu().numpy(), g[name + "_map"][0], rtol=2e-5, atol=1e-7)


def _pair(b2pose, dev, tag, half=False, use_graph=False):
    tkind, textra, sextra, dkw = DISTILL_CASES[tag]
    tcfg = po.net_config(side_in=64, num_joints=17, **textra)
    scfg = po.net_config(side_in=64, num_joints=17, **sextra)
    teacher = getattr(getattr(b2pose, tkind), "resnet18")(tcfg, False)
    teacher.load_state_dict(po.init_state(tkind, "resnet18", tcfg, seed=21))
    student = b2pose.depthnet.resnet18(scfg, False)
    student.load_state_dict(po.init_state("depthnet", "resnet18", scfg, seed=11))
    teacher, student = teacher.to(dev).train(), student.to(dev).train()
    args = b2pose.train_args(model="resnet18", num_joints=17, side_in=64, stride=16, depth_only=False, do_fusion=False,
                             do_teach=True, half_acc=half, alpha_init=0.3, alpha_dest=0.3, **dkw)
    tr = b2pose.Trainer(args, student, dict(key_index=16), use_graph=use_graph)
    tr.set_teacher(teacher)
    if dkw.get("do_freeze"):
        tr.freeze_batchnorm()
    tr.alpha = tr.get_dist_weight(1)
    batch = tuple(t.to(dev) for t in po.synth_distill_batch(2, 64, 17, stride=16, seed=3)[:5])
    return tr, teacher, student, batch


@pytest.mark.parametrize("tag", sorted(DISTILL_CASES))
def test_distill_step_fp32(b2pose, dev, golden_dir, tag):
    g = np.load(golden_dir + "/distill.npz")
    tr, teacher, student, batch = _pair(b2pose, dev, tag)
    cams, dists, gns = [], [], []
    for it in range(2):
        out = tr.train_step(batch)
        cams.append(float(out["cam_loss"])); dists.append(float(out["dist_loss"]))
        gns.append(float(out["grad_sumsq"].sqrt()))
        assert abs(float(out["loss"]) - (dists[-1] * 0.3 + cams[-1])) < 1e-4 * abs(float(out["loss"]))
        if it == 0:
            assert np.abs(out["spec_cam"].cpu().numpy() - g[f"{tag}_spec"]).max() < 0.1            # mm
            coef = min(1.0, 5.0 / (g[f"{tag}_gn"][0] + 1e-6))     # the reference recorded clipped grads
            grads = dict(student.named_parameters())
            got = grads["layer4.1.conv2.weight"].grad.detach().contiguous().reshape(-1)[:64].cpu().numpy() * coef
            assert np.abs(got - g[f"{tag}_g_l4"]).max() / max(np.abs(g[f"{tag}_g_l4"]).max(), 1e-12) < 3e-2
            assert abs(float(grads["layer1.0.conv1.weight"].grad.norm()) * coef - float(g[f"{tag}_gn_l1"])) \
                / float(g[f"{tag}_gn_l1"]) < 3e-2
    np.testing.assert_allclose(cams[0], g[f"{tag}_cam"][0], rtol=1e-4)
    np.testing.assert_allclose(dists[0], g[f"{tag}_dist"][0], rtol=1e-4)
    np.testing.assert_allclose(gns[0], g[f"{tag}_gn"][0], rtol=5e-3)
    # second step: starts from Adam-updated weights (see test_gpu_nets.py for why the bound is looser)
    np.testing.assert_allclose(cams[1], g[f"{tag}_cam"][1], rtol=1e-2)
    np.testing.assert_allclose(dists[1], g[f"{tag}_dist"][1], rtol=1e-2)
    np.testing.assert_allclose(teacher.state_dict()["bn1.running_mean"].cpu().numpy(), g[f"{tag}_teacher_bn1_rm"],
                               rtol=1e-3, atol=1e-5)
    assert all(p.grad is None for p in teacher.parameters())


def test_distill_step_bf16_graph_and_schedule(b2pose, dev, golden_dir):
    """bf16 tensor-core mode through the captured CUDA graph: the losses stay near the fp32 golden values,
    alpha moves without re-capture, and the epoch loop reports the reference's dictionary."""
    g = np.load(golden_dir + "/distill.npz")
    tag = "dist_pf18_l2"
    tr, teacher, student, batch = _pair(b2pose, dev, tag, half=True, use_graph=True)
    out = tr.train_step(batch)
    # (bf16 train-mode BatchNorm on a batch of 2 is chaotic in the last bits of the statistics: see the bound
    # discussion in test_gpu_nets.py::test_net_bf16)
    assert abs(float(out["cam_loss"]) - g[f"{tag}_cam"][0]) / g[f"{tag}_cam"][0] < 5e-2
    assert abs(float(out["dist_loss"]) - g[f"{tag}_dist"][0]) / g[f"{tag}_dist"][0] < 5e-2
    for _ in range(4):                                  # warm-ups, capture, replay
        out = tr.train_step(batch)
    vals = {}
    for alpha in (0.0, 1.0):          # graph outputs are static tensors: read them before the next replay
        tr.alpha = alpha
        o = tr.train_step(batch)
        vals[alpha] = {k: float(o[k]) for k in ("loss", "cam_loss", "dist_loss")}
    assert abs(vals[0.0]["loss"] - vals[0.0]["cam_loss"]) < 1e-3 * vals[0.0]["loss"]
    assert abs(vals[1.0]["loss"] - vals[1.0]["cam_loss"] - vals[1.0]["dist_loss"]) < 1e-3 * vals[1.0]["loss"]
    assert vals[1.0]["cam_loss"] < g[f"{tag}_cam"][0]                # the fixed batch is being fitted
    # semi-supervised mimic term on a second (unlabelled) batch + the epoch loop
    semi = tuple(t.to(dev) for t in po.synth_distill_batch(2, 64, 17, stride=16, seed=9)[:5])
    o2 = tr.train_step(batch, semi)
    assert "semi_dist_loss" in o2 and float(o2["semi_dist_loss"]) > 0
    tr.alpha_init, tr.alpha_dest, tr.alpha_span = 0.5, 0.1, 5
    np.testing.assert_allclose([tr.get_dist_weight(e) for e in range(1, 9)], g["alpha_sched"], rtol=1e-12)
    res = tr.distill_train(2, [batch, batch])
    assert set(res) == {"dist_train_loss", "cam_train_loss"} and abs(tr.alpha - 0.4) < 1e-12
