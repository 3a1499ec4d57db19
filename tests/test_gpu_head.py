"""GPU parity of the volumetric heat-map head, the loss glue, the unprojection and the fused
clip+Adam step against golden fixtures (reference outputs) and the CPU oracle."""
import numpy as np
import pytest
import torch

import pose_oracle as po
from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_head_golden(b2pose, dev, golden_dir, layout):
    g = np.load(golden_dir + "/head.npz")
    for name, spec in zip(g["names"], g["specs"]):
        N, J, D, H, W = [int(v) for v in spec]
        cot = torch.tensor(g[f"{name}_cot"], device=dev)
        for fused in (True, False):
            feat = torch.tensor(g[f"{name}_feat"], device=dev)
            if layout == "nhwc":
                feat = feat.contiguous(memory_format=torch.channels_last)
            feat.requires_grad_(True)
            if fused:
                coords = b2pose.heatmap_coords(feat, D, J, 1000.0)
            else:
                heat = b2pose.to_heatmap(feat, D, J, H, W)
                assert tuple(heat.shape) == (N, J, H, W, D)
                np.testing.assert_allclose(heat.sum(dim=(2, 3, 4)).detach().cpu().numpy(), 1.0, atol=1e-5)
                if f"{name}_heat" in g.files:
                    assert rel_err(heat, g[f"{name}_heat"]) < 1e-4
                coords = b2pose.decode(heat, 1000.0)
            (coords * cot).sum().backward()
            # joint positions within 0.1 mm (BASELINE MPJPE tolerance), gradients 1e-4 relative
            assert float((coords.detach().cpu() - torch.tensor(g[f"{name}_coords"])).abs().max()) < 0.1
            assert rel_err(coords, g[f"{name}_coords"]) < 1e-4
            assert rel_err(feat.grad, g[f"{name}_dfeat"]) < 1e-4, (name, fused)


def test_head_known_answers(b2pose, dev, golden_dir):
    g = np.load(golden_dir + "/head.npz")
    uni = b2pose.heatmap_coords(torch.zeros(1, 16 * 3, 5, 6, device=dev), 16, 3, 1000.0)
    np.testing.assert_allclose(uni.cpu().numpy(), g["ka6_uniform"], rtol=1e-5)
    hot = torch.full((1, 16 * 2, 5, 6), -1e4, device=dev)
    hot[0, 7 * 2 + 1, 3, 4] = 50.0
    hot[0, 2 * 2 + 0, 0, 5] = 50.0
    one = b2pose.heatmap_coords(hot, 16, 2, 1000.0)
    np.testing.assert_allclose(one.cpu().numpy(), g["ka6_onehot"], rtol=1e-5, atol=1e-3)


def test_head_bf16_and_full_size(b2pose, dev):
    gen = torch.Generator().manual_seed(3)
    for (N, J, D, S) in [(64, 17, 16, 16), (8, 25, 16, 17)]:
        feat = (torch.randn(N, D * J, S, S, generator=gen) * 2).bfloat16()
        ref = po.decode(po.to_heatmap(feat.float(), D, J, S, S), 1000.0)
        out = b2pose.heatmap_coords(feat.to(dev).contiguous(memory_format=torch.channels_last), D, J, 1000.0)
        assert float((out.cpu() - ref).abs().max()) < 0.1


@pytest.mark.parametrize("crit", ["SmoothL1", "L1", "MSE"])
def test_pose_loss(b2pose, dev, crit):
    gen = torch.Generator().manual_seed(9)
    N, J, key = 6, 17, 16
    coords = (torch.rand(N, J, 3, generator=gen) * 2000).requires_grad_(True)
    true_cam = torch.randn(N, J, 3, generator=gen) * 300
    coords.data[0, 2] = coords.data[0, key] - true_cam[0, key] + true_cam[0, 2] + 3.0    # |diff| < 1 branch
    valid = torch.rand(N, J, generator=gen) < 0.8
    valid[:, key] = True
    rel = coords - coords[:, key:key + 1]
    spec = rel + true_cam[:, key:key + 1]
    sel = valid.reshape(-1)
    fn = dict(SmoothL1=torch.nn.functional.smooth_l1_loss, L1=torch.nn.functional.l1_loss,
              MSE=torch.nn.functional.mse_loss)[crit]
    loss = fn(spec.reshape(-1, 3)[sel] / 10.0, true_cam.reshape(-1, 3)[sel] / 10.0)
    loss.backward()
    cg = coords.detach().to(dev).requires_grad_(True)
    lg, sg = b2pose.pose_loss(cg, true_cam.to(dev), valid.to(dev), key, 10.0, crit)
    lg.backward()
    assert abs(float(lg) - float(loss)) / float(loss) < 1e-5
    assert rel_err(sg, spec) < 1e-6
    assert rel_err(cg.grad, coords.grad) < 1e-5


def test_to_depth(b2pose, dev, golden_dir):
    g = np.load(golden_dir + "/to_depth.npz")
    for name in g["names"]:
        out = b2pose.to_depth(g[f"{name}_img"], g[f"{name}_K"])
        assert out.dtype == np.float32
        np.testing.assert_allclose(out, g[f"{name}_out"], rtol=2e-6)      # fp32 rounding of 5 ops
        assert np.all(out[g[f"{name}_img"] == 0] == 0)
    # batched device form + an odd width (scalar tail path), against the oracle
    img = torch.rand(3, 17, 23) * 4000
    K = np.array([[365.0, 0, 11.5], [0, 365.0, 8.5], [0, 0, 1]])
    out = b2pose.to_depth(img.to(dev), K).cpu().numpy()
    for i in range(3):
        np.testing.assert_allclose(out[i], po.to_depth(img[i].numpy(), K), rtol=2e-6)


def test_adam_clip(b2pose, dev):
    """Fused clip_grad_norm_ + Adam(L2 weight decay) == torch.optim.Adam after clip_grad_norm_."""
    L = b2pose._lib
    gen = torch.Generator().manual_seed(2)
    n = 10007
    w0 = torch.randn(n, generator=gen)
    p = w0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], 5e-5, weight_decay=4e-5)
    w = w0.to(dev)
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    w16 = torch.zeros(n, dtype=torch.bfloat16, device=dev)
    for step in range(1, 4):
        gr = torch.randn(n, generator=gen) * (10.0 if step == 2 else 0.01)
        p.grad = gr.clone()
        total = torch.nn.utils.clip_grad_norm_([p], 5.0)
        opt.step()
        gd = gr.to(dev)
        ss = torch.zeros(1, dtype=torch.float64, device=dev)
        L.call("b2_grad_sumsq", L.ptr(gd), n, L.ptr(ss), L.stream())
        assert abs(float(ss.sqrt()) - float(total)) / float(total) < 1e-5
        L.call("b2_adam_step", L.ptr(w), L.ptr(gd), L.ptr(m), L.ptr(v), L.ptr(w16), n, 5e-5, 0.9, 0.999, 1e-8,
               4e-5, step, L.ptr(ss), 5.0, 1.0, None, L.stream())
        assert rel_err(w, p) < 1e-6
        assert torch.equal(w16, w.bfloat16())
    # non-finite gradient norm -> the update is skipped (depth_train.py:435-438)
    before = w.clone()
    ss = torch.full((1,), float("inf"), dtype=torch.float64, device=dev)
    L.call("b2_adam_step", L.ptr(w), L.ptr(gd), L.ptr(m), L.ptr(v), None, n, 5e-5, 0.9, 0.999, 1e-8, 4e-5, 4,
           L.ptr(ss), 5.0, 1.0, None, L.stream())
    assert torch.equal(w, before)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 64, 16, 16), (1, 64, 17, 13), (3, 8, 9, 9), (2, 12, 6, 7)])
def test_maxpool_with_veil(b2pose, dev, shape, dtype):
    """MaxPool2d(3,2,1) on x and the veil in one launch (partial_depthnet.py:219-220), fwd + bwd."""
    N, C, H, W = shape
    gen = torch.Generator().manual_seed(N * 1000 + C + H)
    x = torch.randn(N, C, H, W, generator=gen).to(dtype).float()
    veil = (torch.rand(N, 1, H, W, generator=gen) > 0.6).float()
    xr = x.clone().requires_grad_(True)
    yr = torch.nn.functional.max_pool2d(xr, 3, 2, 1)
    vr = torch.nn.functional.max_pool2d(veil, 3, 2, 1)
    cot = torch.randn(yr.shape, generator=gen).to(dtype).float()
    (yr * cot).sum().backward()
    xg = x.permute(0, 2, 3, 1).contiguous().to(dev).to(dtype).requires_grad_(True)
    yg, vg = b2pose.ops.MaxPoolFn.apply(xg, veil[:, 0].contiguous().to(dev))
    (yg.float() * cot.permute(0, 2, 3, 1).to(dev)).sum().backward()
    assert torch.equal(yg.float().cpu().permute(0, 3, 1, 2), yr.detach())        # selection: exact
    assert torch.equal(vg.cpu(), vr[:, 0])
    assert rel_err(xg.grad.float().permute(0, 3, 1, 2), xr.grad) < (1e-6 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("C,rows,relu,res", [(64, 5000, 1, False), (256, 1237, 1, True), (2048, 64, 0, False)])
def test_bn_totals_path_matches_slot_path(b2pose, dev, C, rows, relu, res):
    """The totals BatchNorm path (fp32 reductions into one float[2C], finalize folded into the consumers)
    against the slot path (deterministic partials + fp64 finalize kernels): statistics agree to fp32
    rounding, outputs and gradients to one bf16 ulp."""
    L = b2pose._lib
    P, st = L.ptr, L.stream()
    assert L.lib().b2_bn_totals_supported(C, L.BF16) == 1
    assert L.lib().b2_bn_totals_supported(C, L.F32) == 0 and L.lib().b2_bn_totals_supported(48, L.BF16) == 0
    g = torch.Generator(device="cpu").manual_seed(C + rows)
    y = (torch.randn(rows, C, generator=g) * 1.7 + 0.3).to(dev).bfloat16()
    dz = torch.randn(rows, C, generator=g).to(dev).bfloat16()
    resid = torch.randn(rows, C, generator=g).to(dev).bfloat16() if res else None
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev)
    beta = (torch.randn(C, generator=g) * 0.1).to(dev)
    row_mask = (torch.rand(rows, generator=g) > 0.2).float().to(dev)
    ratio = (torch.rand(rows, generator=g) + 0.5).to(dev)
    out = {}
    for mode in ("slots", "totals"):
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        mean, invstd = torch.empty(C, device=dev), torch.empty(C, device=dev)
        z, dy = torch.empty_like(y), torch.empty_like(y)
        dres = torch.empty_like(y) if res else None
        dgamma, dbeta = torch.full((C,), 2.0, device=dev), torch.full((C,), -1.0, device=dev)
        zsave = z if (relu and res) else None
        if mode == "slots":
            parts = torch.empty(L.BN_PARTS * 2 * C, device=dev)
            L.call("b2_bn_stats", P(y), rows, C, L.BF16, P(parts), st)
            L.call("b2_bn_finalize", P(parts), rows, C, P(rm), P(rv), 0.1, 1e-5, 1, P(mean), P(invstd), st)
            L.call("b2_bn_apply", P(y), P(mean), P(invstd), P(gamma), P(beta), P(resid), P(row_mask), relu, P(z), rows, C,
                   L.BF16, st)
            parts2, gsum = torch.empty(L.BN_PARTS * 2 * C, device=dev), torch.empty(2 * C, device=dev)
            L.call("b2_bn_bwd_reduce", P(dz), P(zsave), P(y), P(mean), P(invstd), P(gamma), P(beta), P(row_mask), relu,
                   P(parts2), rows, C, L.BF16, st)
            L.call("b2_bn_bwd_finalize", P(parts2), C, P(gsum), P(dgamma), P(dbeta), st)
            L.call("b2_bn_bwd_apply", P(dz), P(zsave), P(y), P(mean), P(invstd), P(gamma), P(beta), P(gsum), P(row_mask),
                   P(ratio), relu, 1, P(dy), P(dres), rows, C, L.BF16, st)
        else:
            totals, gsum = torch.zeros(2 * C, device=dev), torch.zeros(2 * C, device=dev)
            # residual + ReLU layers hand the gate over as a bitmask instead of the saved output z
            gate = torch.empty((rows * C // 8 + 15) // 16 * 16, dtype=torch.uint8, device=dev) if (relu and res) else None
            L.call("b2_bn_stats_totals", P(y), rows, C, L.BF16, P(totals), st)
            L.call("b2_bn_apply_totals", P(y), P(totals), rows, P(rm), P(rv), 0.1, 1e-5, 1, P(gamma), P(beta), P(resid),
                   P(row_mask), relu, P(z), P(mean), P(invstd), P(gate), C, L.BF16, st)
            L.call("b2_bn_bwd_reduce_totals", P(dz), None, P(y), P(mean), P(invstd), P(gamma), P(beta), P(row_mask),
                   relu, P(gsum), P(gate), rows, C, L.BF16, st)
            L.call("b2_bn_bwd_apply_totals", P(dz), None, P(y), P(mean), P(invstd), P(gamma), P(beta), P(gsum),
                   P(row_mask), P(ratio), relu, 1, P(dy), P(dres), P(dgamma), P(dbeta), P(gate), rows, C, L.BF16, st)
        torch.cuda.synchronize()
        out[mode] = dict(mean=mean, invstd=invstd, rm=rm, rv=rv, z=z.float(), dy=dy.float(), gsum=gsum, dgamma=dgamma,
                         dbeta=dbeta, dres=None if dres is None else dres.float())
    a, b = out["slots"], out["totals"]
    for k in ("mean", "invstd", "rm", "rv", "gsum", "dgamma", "dbeta"):
        assert rel_err(b[k], a[k]) < 2e-6, k
    for k in ("z", "dy"):
        assert rel_err(b[k], a[k]) < 1e-2 and float((b[k] != a[k]).float().mean()) < 1e-3, k
    if res:
        assert torch.equal(a["dres"], b["dres"])
    # frozen statistics (eval): running stats in, no totals
    mean2, invstd2, z2 = torch.empty(C, device=dev), torch.empty(C, device=dev), torch.empty_like(y)
    rm, rv = a["rm"].clone(), a["rv"].clone()
    L.call("b2_bn_apply_totals", P(y), None, rows, P(rm), P(rv), 0.1, 1e-5, 0, P(gamma), P(beta), P(resid), P(row_mask),
           relu, P(z2), P(mean2), P(invstd2), None, C, L.BF16, st)
    mean3, invstd3, z3 = torch.empty(C, device=dev), torch.empty(C, device=dev), torch.empty_like(y)
    L.call("b2_bn_finalize", None, rows, C, P(rm), P(rv), 0.1, 1e-5, 0, P(mean3), P(invstd3), st)
    L.call("b2_bn_apply", P(y), P(mean3), P(invstd3), P(gamma), P(beta), P(resid), P(row_mask), relu, P(z3), rows, C,
           L.BF16, st)
    assert torch.equal(rm, a["rm"]) and torch.equal(mean2, mean3) and rel_err(invstd2, invstd3) < 1e-6
    assert rel_err(z2.float(), z3.float()) < 1e-2


def test_head2d_and_projection(b2pose, dev, golden_dir):
    """mat_utils.to_heatmap / decode (the D = 1 case of the head kernels) and
    back_project.projectPoints against outputs of the reference's own functions."""
    g = np.load(golden_dir + "/head2d.npz")
    M = b2pose.mat_utils
    for name in ("sq", "rect"):
        cot = torch.tensor(g[f"{name}_cot"], device=dev)
        for fused in (False, True):
            feat = torch.tensor(g[f"{name}_feat"], device=dev, requires_grad=True)
            N, J, H, W = feat.shape
            if fused:
                coords = M.heatmap_coords(feat, J, 257.0)
            else:
                heat = M.to_heatmap(feat, J, H, W)
                assert tuple(heat.shape) == (N, J, H, W) and rel_err(heat, g[f"{name}_heat"]) < 1e-4
                coords = M.decode(heat, 257.0)
            (coords * cot).sum().backward()
            assert tuple(coords.shape) == (N, J, 2)
            assert float((coords.detach().cpu() - torch.tensor(g[f"{name}_coords"])).abs().max()) < 1e-2      # pixels
            assert rel_err(coords, g[f"{name}_coords"]) < 1e-4 and rel_err(feat.grad, g[f"{name}_dfeat"]) < 1e-4
    cam = dict(K=g["proj_K"], R=g["proj_R"], t=g["proj_t"], distCoef=g["proj_Kd"])
    got = b2pose.back_project.projectPoints(g["proj_X"], cam)
    np.testing.assert_allclose(got, g["proj_out"], rtol=2e-5, atol=1e-3)
    got_t = b2pose.back_project.projectPoints(torch.tensor(g["proj_X"], device=dev, dtype=torch.float32), cam)
    np.testing.assert_allclose(got_t.cpu().numpy(), g["proj_out"], rtol=2e-5, atol=1e-3)
