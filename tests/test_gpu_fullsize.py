"""Size-independent properties at BASELINE.json's full sizes (batch 64, 256x256 inputs) on the
tensor-core path -- where the CPU oracle would take minutes -- plus spot checks of random output
pixels against a direct fp64 evaluation of the reference formula (partial_conv.py:39-53)."""
import numpy as np
import pytest
import torch

import pose_oracle as po
from conftest import rel_err

pytestmark = pytest.mark.gpu


def _mask(n, side, frac, seed):
    return po.blob_mask(n, side, frac, torch.Generator().manual_seed(seed))[:, 0].contiguous()


@pytest.mark.parametrize("shape", [(64, 64, 64, 64, 64, 3, 1, 1, 1), (64, 32, 32, 128, 128, 3, 2, 1, 1),
                                   (64, 64, 64, 64, 256, 1, 1, 0, 1)])
def test_pconv_full_size_properties(b2pose, dev, shape):
    N, H, W, C, K, k, s, p, d = shape
    gen = torch.Generator().manual_seed(7)
    conv = b2pose.PartialConv(C, K, kernel_size=k, stride=s, padding=p, dilation=d, bias=False).to(dev)
    x = torch.randn(N, H, W, C, generator=gen).to(dev).bfloat16()
    m = _mask(N, H, 0.25, 3).to(dev)
    xm = x * m.unsqueeze(-1).bfloat16()
    y, mo = conv.forward_nhwc(xm, m, premasked=True)
    # (1) mask algebra: mask_out is the binary dilation of the mask (max-pool with the conv geometry)
    ref_mo = torch.nn.functional.max_pool2d(m.unsqueeze(1), k, s, p, d)[:, 0] if d == 1 else None
    if ref_mo is not None:
        assert torch.equal(mo, ref_mo)
    assert set(np.unique(mo.cpu().numpy()).tolist()) <= {0.0, 1.0}
    # (2) outputs are exactly zero where the updated mask is zero
    assert float(y[mo == 0].abs().max()) == 0.0 if bool((mo == 0).any()) else True
    # (3) homogeneity: scaling the input by a power of two scales the output exactly (bf16)
    y2, _ = conv.forward_nhwc(xm * 4, m, premasked=True)
    assert torch.equal(y2, y * 4)
    # (4) with an all-ones mask the layer is the plain convolution times k*k/(#in-bounds taps + 1e-6)
    ones = torch.ones_like(m)
    ya, moa = conv.forward_nhwc(x, ones, premasked=True)
    assert bool((moa == 1).all())
    plain = b2pose.Conv2d(C, K, kernel_size=k, stride=s, padding=p, dilation=d, bias=False).to(dev)
    plain.weight.data.copy_(conv.weight.data)
    yp = plain.forward_nhwc(x)
    cnt = torch.nn.functional.avg_pool2d(ones.unsqueeze(1), k, s, p, divisor_override=1)[:, 0] if d == 1 else None
    if cnt is not None:
        ratio = (float(k * k) / (cnt + 1e-6)).unsqueeze(-1)
        assert rel_err(ya.float(), yp.float() * ratio) < 1e-2
    # (5) spot check 64 random outputs against an fp64 evaluation of the reference formula
    w = conv.weight.detach().double().cpu()
    xc, mc = xm.double().cpu(), m.double().cpu()
    rs = np.random.RandomState(0)
    Ho, Wo = y.shape[1], y.shape[2]
    for _ in range(64):
        n, oh, ow, kk = rs.randint(N), rs.randint(Ho), rs.randint(Wo), rs.randint(K)
        acc, c = 0.0, 0.0
        for r in range(k):
            for t in range(k):
                ih, iw = oh * s - p + r * d, ow * s - p + t * d
                if 0 <= ih < H and 0 <= iw < W:
                    c += float(mc[n, ih, iw])
                    acc += float((xc[n, ih, iw] * w[kk, :, r, t]).sum())
        want = acc * (k * k / (c + 1e-6)) * min(max(c, 0.0), 1.0)
        got = float(y[n, oh, ow, kk])
        assert abs(got - want) <= 2e-2 * max(1.0, abs(want)), (n, oh, ow, kk, got, want)


def test_head_full_size_properties(b2pose, dev):
    gen = torch.Generator().manual_seed(2)
    N, J, D, S = 64, 17, 16, 16
    feat = (torch.randn(N, S, S, D * J, generator=gen) * 3).to(dev).bfloat16().permute(0, 3, 1, 2)
    c0 = b2pose.heatmap_coords(feat, D, J, 1000.0)
    assert float(c0.min()) >= 0.0 and float(c0.max()) <= 2000.0
    # softmax is shift invariant: adding a constant to one joint's logits changes nothing
    heat = b2pose.to_heatmap(feat.float(), D, J, S, S)
    np.testing.assert_allclose(heat.sum(dim=(2, 3, 4)).cpu().numpy(), 1.0, atol=1e-4)
    assert float((b2pose.decode(heat, 1000.0) - c0).abs().max()) < 0.05
    shifted = feat.float() + 5.0
    assert float((b2pose.heatmap_coords(shifted, D, J, 1000.0) - c0).abs().max()) < 0.05
    # flipping the volume along W mirrors x: x -> 2000 - x
    flipped = torch.flip(feat, dims=[3]).contiguous(memory_format=torch.channels_last)
    cf = b2pose.heatmap_coords(flipped, D, J, 1000.0)
    assert float((cf[..., 0] - (2000.0 - c0[..., 0])).abs().max()) < 0.05
    assert float((cf[..., 1:] - c0[..., 1:]).abs().max()) < 0.05


@pytest.mark.parametrize("workload", ["partial_fusionnet", "partial_depthnet"])
def test_full_size_training_step_sanity(b2pose, dev, workload):
    """Batch 16 at 256x256 (and 257x257), bf16 + CUDA graph: finite loss that decreases on a fixed
    batch, BN-normalised activations, veil fully healed by layer2, weights move."""
    fused = "fusion" in workload
    for side, J in ((256, 17), (257, 25)):
        cfg = b2pose.train_args(model="resnet50", num_joints=J, side_in=side, depth_only=not fused, do_fusion=fused,
                                half_acc=True)
        torch.manual_seed(0)
        net = getattr(getattr(b2pose, workload), "resnet50")(cfg, False).to(dev).train()
        w0 = net.layer3[0].conv1.weight.detach().clone()
        tr = b2pose.Trainer(cfg, net, dict(key_index=J - 1), use_graph=True)
        batch = b2pose.synthetic_batch(16, side, J, dev, seed=4)
        losses = [float(tr.train_step(batch)["loss"]) for _ in range(8)]
        assert all(np.isfinite(losses)), losses
        assert losses[-1] < losses[0], losses
        assert float((net.layer3[0].conv1.weight.detach() - w0).abs().max()) > 0
        so = (side - 1) // 16 + 1
        z, feat = net(batch[0], batch[1]) if fused else net(batch[1])
        assert tuple(z.shape) == (16, 16 * J, so, so) and tuple(feat.shape) == (16, 2048, so, so)
