"""CPU: pin oracle/pose_oracle.py against fixtures produced by executing the reference
(oracle/make_golden.py).  No product code involved."""
import os

import numpy as np
import pytest
import torch

import pose_oracle as po


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_known_answers(golden_dir):
    g = load(golden_dir, "ka")
    x = torch.tensor(g["ka1_x"], requires_grad=True)
    w = torch.ones(1, 1, 3, 3, requires_grad=True)
    b = torch.full((1,), 0.5, requires_grad=True)
    y, mo = po.partial_conv(x, torch.tensor(g["ka1_mask"]), w, b, 1, 1, 1)
    y.sum().backward()
    assert np.array_equal(mo.numpy(), g["ka1_mask_out"])
    np.testing.assert_allclose(y.detach().numpy(), g["ka1_out"], rtol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), g["ka1_dx"], rtol=1e-6)
    np.testing.assert_allclose(w.grad.numpy(), g["ka1_dw"], rtol=1e-6)
    assert float(b.grad) == float(g["ka1_db"][0]) == 15.0
    assert np.all(x.grad.numpy()[g["ka1_mask"] == 0] == 0)
    # SURVEY KA1 literal values
    np.testing.assert_allclose(g["ka1_out"][0, 0, 0], [0, 63.49993896, 67.99996948, 67.99996948], rtol=1e-7)
    np.testing.assert_allclose(float(g["ka1_dw"][0, 0, 1, 1]), 202.83210754, rtol=1e-7)
    # KA2 ratios bit-exact
    for (win, cnt), r in zip(g["ka2_pairs"], g["ka2_ratio"]):
        assert np.float32(po.renorm_ratio(int(win), int(cnt))) == r, (win, cnt)
    assert np.all(g["ka3_out"] == 0) and np.all(g["ka3_mask_out"] == 0)


def test_pconv_cases(golden_dir):
    g = load(golden_dir, "pconv")
    for name, spec in zip(g["names"], g["specs"]):
        N, C, K, H, W, k, s, p, d, has_bias = [int(v) for v in spec]
        x = torch.tensor(g[f"{name}_x"], requires_grad=True)
        w = torch.tensor(g[f"{name}_w"], requires_grad=True)
        b = torch.tensor(g[f"{name}_b"], requires_grad=True) if has_bias else None
        m = torch.tensor(g[f"{name}_mask"])
        y, mo = po.partial_conv(x, m, w, b, s, p, d)
        (y * torch.tensor(g[f"{name}_cot"])).sum().backward()
        assert np.array_equal(mo.numpy(), g[f"{name}_mask_out"]), name
        assert rel_err(y.detach(), g[f"{name}_out"]) < 1e-6, name
        assert rel_err(x.grad, g[f"{name}_dx"]) < 1e-6, name
        assert rel_err(w.grad, g[f"{name}_dw"]) < 1e-6, name
        if has_bias:
            assert rel_err(b.grad, g[f"{name}_db"]) < 1e-6, name
        # independent loop implementation (no ATen conv)
        if N * K * H * W * C * k * k < 3e6:
            yl, ml = po.partial_conv_loops(g[f"{name}_x"], g[f"{name}_mask"], g[f"{name}_w"],
                                           g[f"{name}_b"] if has_bias else None, s, p, d)
            assert np.array_equal(ml, g[f"{name}_mask_out"]), name
            assert rel_err(yl, g[f"{name}_out"]) < 2e-6, name
        # dtype rule (KA4)
        yb, mob = po.partial_conv(x.detach().bfloat16(), m, w.detach().bfloat16(),
                                  b.detach().bfloat16() if has_bias else None, s, p, d)
        assert mob.dtype == torch.float32
        assert yb.dtype == (torch.float32 if has_bias else torch.bfloat16)
        assert rel_err(yb.float(), g[f"{name}_out_bf16"]) < 1e-6, name


def test_head(golden_dir):
    g = load(golden_dir, "head")
    for name, spec in zip(g["names"], g["specs"]):
        N, J, D, H, W = [int(v) for v in spec]
        feat = torch.tensor(g[f"{name}_feat"], requires_grad=True)
        heat = po.to_heatmap(feat, D, J, H, W)
        coords = po.decode(heat, 1000.0)
        (coords * torch.tensor(g[f"{name}_cot"])).sum().backward()
        assert heat.shape == (N, J, H, W, D)
        assert rel_err(coords.detach(), g[f"{name}_coords"]) < 1e-6
        assert rel_err(feat.grad, g[f"{name}_dfeat"]) < 1e-5
        np.testing.assert_allclose(heat.sum(dim=(2, 3, 4)).detach().numpy(), 1.0, atol=1e-5)
    np.testing.assert_allclose(g["ka6_uniform"], 1000.0, rtol=1e-6)
    # one-hot voxel (h,w,d) -> (2w/(W-1), 2h/(H-1), 2d/(D-1)) * range   (H=5, W=6, D=16)
    np.testing.assert_allclose(g["ka6_onehot"][0, 1], [2 * 4 / 5 * 1000, 2 * 3 / 4 * 1000, 2 * 7 / 15 * 1000], rtol=1e-5)
    np.testing.assert_allclose(g["ka6_onehot"][0, 0], [2 * 5 / 5 * 1000, 0.0, 2 * 2 / 15 * 1000], rtol=1e-5, atol=1e-3)


def test_to_depth(golden_dir):
    g = load(golden_dir, "to_depth")
    for name in g["names"]:
        out = po.to_depth(g[f"{name}_img"], g[f"{name}_K"])
        assert out.dtype == np.float32
        np.testing.assert_allclose(out, g[f"{name}_out"], rtol=1e-7)


def test_shapes(golden_dir):
    g = load(golden_dir, "shapes")
    for kind, n in (("partial_depthnet", 28515536), ("partial_fusionnet", 30485776), ("fusionnet", 30485776)):
        cfg = po.net_config(side_in=256, num_joints=17)
        shapes = po.param_shapes(kind, "resnet50", cfg)
        assert list(shapes.keys()) == [str(k) for k in g[f"{kind}_keys"]]
        train = [k for k in shapes if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
        assert sum(int(np.prod(shapes[k])) for k in train) == int(g[f"{kind}_nparams"]) == n


NET_CASES = {
    "pdepth18_s65": ("partial_depthnet", "resnet18", 65, 2, 17, {}),
    "pdepth50_s64": ("partial_depthnet", "resnet50", 64, 2, 17, {}),
    "pfusion50_s64": ("partial_fusionnet", "resnet50", 64, 2, 17, {}),
    "pfusion18_s49_j25": ("partial_fusionnet", "resnet18", 49, 2, 25, {}),
    "fusion50_s64": ("fusionnet", "resnet50", 64, 2, 17, {}),
    "fusion18_skip": ("fusionnet", "resnet18", 64, 2, 17, dict(skip_relu=True, early_dist=True)),
    "depth50_rgb_s64": ("depthnet", "resnet50", 64, 2, 19, dict(depth_only=False)),
    "depth18_d_s33": ("depthnet", "resnet18", 33, 3, 17, {}),
    "legacy50_s64": ("resnet", "resnet50", 64, 2, 19, {}),
    "pdepth50_stride8": ("partial_depthnet", "resnet50", 64, 2, 17, dict(stride=8)),
}


@pytest.mark.parametrize("tag", sorted(NET_CASES))
def test_net_and_step(golden_dir, tag):
    g = load(golden_dir, "nets")
    kind, model, side, N, J, extra = NET_CASES[tag]
    cfg = po.net_config(side_in=side, num_joints=J, **extra)
    sd = po.init_state(kind, model, cfg, seed=11)
    batch = po.synth_batch(N, side, J, seed=3, invalid_frac=0.25)
    with torch.no_grad():
        if kind in ("fusionnet", "partial_fusionnet"):
            z_eval = po.net_forward(sd, kind, model, cfg, batch[0], batch[1], training=False)[0]
        elif kind == "resnet":
            z_eval = po.net_forward(sd, kind, model, cfg, batch[0], None, training=False)
        else:
            inp = batch[1] if (kind == "partial_depthnet" or cfg.depth_only) else batch[0]
            z_eval = po.net_forward(sd, kind, model, cfg, inp, None, training=False)[0]
    assert rel_err(z_eval, g[f"{tag}_z_eval"]) < 1e-5
    orc = po.StepOracle(sd, kind, model, cfg, key_index=J - 1)
    losses, gns = [], []
    for it in range(2):
        loss, gn, spec, z = orc.step(batch)
        losses.append(loss)
        gns.append(gn)
        if it == 0:
            assert rel_err(z, g[f"{tag}_z"]) < 1e-5
            assert rel_err(spec, g[f"{tag}_spec"]) < 1e-5
    np.testing.assert_allclose(losses, g[f"{tag}_loss"], rtol=2e-4)
    np.testing.assert_allclose(gns, g[f"{tag}_gradnorm"], rtol=5e-3)
    np.testing.assert_allclose(sd["bn1.running_mean"].detach().numpy(), g[f"{tag}_bn1_running_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sd["bn1.running_var"].detach().numpy(), g[f"{tag}_bn1_running_var"], rtol=1e-4)
    np.testing.assert_allclose(sd["conv1.weight"].detach().reshape(-1)[:64].numpy(), g[f"{tag}_conv1_after"],
                               rtol=1e-3, atol=2e-5)
    assert sum(sd[k].numel() for k in po.trainable_names(sd)) == int(g[f"{tag}_nparams"])


# ------------------------------------------------------------------ distillation step
DISTILL_CASES = {
    "dist_pf18_l2": ("partial_fusionnet", {}, dict(depth_only=False), dict()),
    "dist_f18skip_sig": ("fusionnet", dict(skip_relu=True, early_dist=True),
                         dict(depth_only=False, skip_relu=True, early_dist=True), dict(sigmoid=True)),
    "dist_pf18_bce_frz": ("partial_fusionnet", {}, dict(depth_only=False), dict(bin_dist=True, freeze=True)),
}
MIMIC_KW = {"l2": dict(sigmoid=False, bin_dist=False), "sigmoid": dict(sigmoid=True, bin_dist=False),
            "bce": dict(sigmoid=False, bin_dist=True)}


def test_distill_loss_attention_schedule(golden_dir):
    g = load(golden_dir, "distill")
    for key in g["loss_names"]:
        mode = str(key).rsplit("_", 1)[1]
        s = torch.tensor(g[f"{key}_s"], requires_grad=True)
        loss = po.distill_loss(torch.tensor(g[f"{key}_t"]), s, torch.tensor(g[f"{key}_a"]), **MIMIC_KW[mode])
        loss.backward()
        np.testing.assert_allclose(float(loss), float(g[f"{key}_loss"]), rtol=1e-6)
        assert rel_err(s.grad, g[f"{key}_ds"]) < 1e-5, key
    for name in ("att_257", "att_64", "att_48s8"):
        side, stride = (int(v) for v in g[name + "_cfg"])
        np.testing.assert_allclose(po.get_attention(side, stride, g[name + "_coords"], True), g[name + "_map"], rtol=1e-12)
        assert np.array_equal(po.get_attention(side, stride, g[name + "_coords"], False), g[name + "_ones"])
        assert g[name + "_map"].shape == (1, (side - 1) // stride + 1, (side - 1) // stride + 1)
        assert g[name + "_map"].max() == 1.0
    sched = [po.dist_weight_at(e, 0.5, 0.1, 5) for e in range(1, 9)]
    np.testing.assert_allclose(sched, g["alpha_sched"], rtol=0, atol=0)


@pytest.mark.parametrize("tag", sorted(DISTILL_CASES))
def test_distill_step(golden_dir, tag):
    g = load(golden_dir, "distill")
    tkind, textra, sextra, dkw = DISTILL_CASES[tag]
    tcfg = po.net_config(side_in=64, num_joints=17, **textra)
    scfg = po.net_config(side_in=64, num_joints=17, **sextra)
    orc = po.DistillOracle(po.init_state("depthnet", "resnet18", scfg, seed=11), "depthnet", "resnet18", scfg,
                           po.init_state(tkind, "resnet18", tcfg, seed=21), tkind, tcfg, key_index=16, **dkw)
    batch = po.synth_distill_batch(2, 64, 17, stride=16, seed=3)
    cams, dists, gns = [], [], []
    for it in range(2):
        cam, dist, gn, spec, last = orc.step(batch, 0.3)
        cams.append(cam); dists.append(dist); gns.append(gn)
        if it == 0:
            assert rel_err(spec, g[f"{tag}_spec"]) < 1e-5
            assert rel_err(last[:, :8], g[f"{tag}_last_slice"]) < 1e-5
    np.testing.assert_allclose(cams, g[f"{tag}_cam"], rtol=2e-4)
    np.testing.assert_allclose(dists, g[f"{tag}_dist"], rtol=2e-4)
    np.testing.assert_allclose(gns, g[f"{tag}_gn"], rtol=5e-3)
    np.testing.assert_allclose(orc.tsd["bn1.running_mean"].numpy(), g[f"{tag}_teacher_bn1_rm"], rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------ input pipeline
def test_pipeline(golden_dir):
    g = load(golden_dir, "pipeline")
    for name in g["names"]:
        hom = po.homography(g[f"{name}_K_old"], g[f"{name}_R_old"], g[f"{name}_K_new"], g[f"{name}_R_new"])
        assert np.array_equal(hom, g[f"{name}_hom"])
        side = g[f"{name}_color_crop"].shape[0]
        crop = po.remap_bilinear(g[f"{name}_color"], hom, (side, side))
        assert np.array_equal(crop, g[f"{name}_color_crop"]), name                   # cv2.remap, bit exact (uint8)
        np.testing.assert_allclose(po.normalize_rgb(crop).numpy(), g[f"{name}_color_out"], rtol=1e-6, atol=1e-7)
        dcrop = po.remap_bilinear(g[f"{name}_depth"], hom, (side, side))
        np.testing.assert_allclose(dcrop, g[f"{name}_depth_crop"], rtol=1e-6, atol=1e-9)
        d = g[f"{name}_depth_crop"]
        np.testing.assert_allclose(po.enhance(d, True, "ntu"), g[f"{name}_ntu_exp"], rtol=1e-6)
        np.testing.assert_allclose(po.enhance(d, False, "ntu"), g[f"{name}_ntu_lin"], rtol=1e-6)
        np.testing.assert_allclose(po.enhance(d, True, "pku"), g[f"{name}_pku_exp"], rtol=1e-6)
        np.testing.assert_allclose(po.enhance(po.to_depth(d, g[f"{name}_K"]), True, "ntu"), g[f"{name}_todepth_ntu_exp"],
                                   rtol=1e-5)
        assert g[f"{name}_ntu_exp"].shape == (1, side, side) and (g[f"{name}_ntu_exp"] == 0).any()
    assert (g["corner_flip_color_crop"] == 0).all(-1).any()          # the crop leaves the frame: border pixels


# ------------------------------------------------------------------ evaluation metrics
def test_metrics(golden_dir):
    g = load(golden_dir, "metrics")
    thresh = dict(zip(("solid", "close", "rough"), g["thresh"]))
    keys = ("solid", "close", "depth", "jitter", "switch", "fail", "score_pck", "score_auc", "cam_mean", "batch_size")
    stats = []
    for b in range(int(g["n_batches"])):
        s = po.analyze(g[f"b{b}_spec"], g[f"b{b}_true"], g[f"b{b}_valid"], g["mirror"], thresh, g[f"b{b}_rot"])
        for k in keys:
            np.testing.assert_allclose(s[k], g[f"b{b}_{k}"], rtol=1e-6), (b, k)
        assert abs(sum(s[k] for k in keys[:6]) - 1.0) < 1e-12            # the taxonomy partitions the valid joints
        stats.append(s)
    ep = po.parse_epoch(stats)
    for k in keys[:-1]:
        np.testing.assert_allclose(ep[k], g[f"epoch_{k}"], rtol=1e-6)


# ------------------------------------------------------------------ 2-D head + projection
def test_head2d_and_projection(golden_dir):
    g = load(golden_dir, "head2d")
    for name in ("sq", "rect"):
        feat = torch.tensor(g[f"{name}_feat"], requires_grad=True)
        N, J, H, W = feat.shape
        heat = po.mat_to_heatmap(feat, J, H, W)
        coords = po.mat_decode(heat, 257.0)
        (coords * torch.tensor(g[f"{name}_cot"])).sum().backward()
        assert rel_err(heat.detach(), g[f"{name}_heat"]) < 1e-6 and rel_err(coords.detach(), g[f"{name}_coords"]) < 1e-6
        assert rel_err(feat.grad, g[f"{name}_dfeat"]) < 1e-5
    cam = dict(K=g["proj_K"], R=g["proj_R"], t=g["proj_t"], distCoef=g["proj_Kd"])
    np.testing.assert_allclose(po.project_points(g["proj_X"], cam), g["proj_out"], rtol=1e-12)


# ------------------------------------------------------------------ property: the two partial-conv restatements agree
def test_partial_conv_restatements_agree_on_random_geometry():
    """The torch-op oracle (library convolution) and the explicit-loop oracle (no library) must agree on random
    small geometries: ragged sizes, strides, dilations, all-invalid windows, with and without bias."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 2), st.integers(1, 3), st.integers(1, 4), st.integers(3, 9), st.integers(3, 9),
           st.sampled_from([1, 3, 5]), st.integers(1, 2), st.integers(0, 2), st.integers(1, 2), st.booleans(),
           st.floats(0.0, 1.0), st.integers(0, 10 ** 6))
    def check(N, C, K, H, W, k, stride, pad, dil, with_bias, invalid, seed):
        if H + 2 * pad - dil * (k - 1) < 1 or W + 2 * pad - dil * (k - 1) < 1:
            return
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(N, C, H, W, generator=g)
        mask = (torch.rand(N, 1, H, W, generator=g) >= invalid).float()
        w = torch.randn(K, C, k, k, generator=g) * 0.3
        b = torch.randn(K, generator=g) if with_bias else None
        y, mo = po.partial_conv(x, mask, w, b, stride, pad, dil)
        yl, ml = po.partial_conv_loops(x.numpy(), mask.numpy(), w.numpy(), None if b is None else b.numpy(), stride, pad, dil)
        assert np.array_equal(mo.numpy(), ml)                       # updated masks bit-exact
        np.testing.assert_allclose(y.numpy(), yl, rtol=2e-5, atol=2e-5)
        assert np.all(y.numpy()[np.broadcast_to(ml == 0, y.shape)] == 0)      # all-invalid windows give exactly 0

    check()
