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
