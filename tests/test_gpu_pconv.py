"""GPU parity of PartialConv (forward + backward) through the C ABI against (i) the golden
fixtures produced by the reference's own code and (ii) the CPU oracle on seeded inputs.

Tolerances (BASELINE.json): updated masks bit-exact; fp32 outputs 1e-4 relative; bf16 2e-2 relative.
"""
import numpy as np
import pytest
import torch

import pose_oracle as po
from conftest import rel_err

pytestmark = pytest.mark.gpu
FP32_TOL, BF16_TOL = 1e-4, 2e-2


def _module(b2pose, dev, C, K, k, s, p, d, w, b=None, dtype=torch.float32):
    conv = b2pose.PartialConv(C, K, kernel_size=k, stride=s, padding=p, dilation=d, bias=b is not None).to(dev)
    with torch.no_grad():
        conv.weight.copy_(torch.as_tensor(w))
        if b is not None:
            conv.bias.copy_(torch.as_tensor(b))
    return conv


def test_known_answers(b2pose, dev, golden_dir):
    g = np.load(golden_dir + "/ka.npz")
    conv = _module(b2pose, dev, 1, 1, 3, 1, 1, 1, np.ones((1, 1, 3, 3), np.float32), np.full(1, 0.5, np.float32))
    x = torch.tensor(g["ka1_x"], device=dev, requires_grad=True)
    y, mo = conv(x, torch.tensor(g["ka1_mask"], device=dev))
    y.sum().backward()
    assert np.array_equal(mo.cpu().numpy(), g["ka1_mask_out"])
    np.testing.assert_allclose(y.detach().cpu().numpy(), g["ka1_out"], rtol=2e-6)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["ka1_dx"], rtol=2e-6)
    np.testing.assert_allclose(conv.weight.grad.cpu().numpy(), g["ka1_dw"], rtol=2e-6)
    assert float(conv.bias.grad) == 15.0
    assert np.all(x.grad.cpu().numpy()[g["ka1_mask"] == 0] == 0)
    # KA2: the renormalisation ratios are bit-exact (mask-only entry point, all-ones / partial windows)
    # KA3: all-invalid window with bias -> exactly 0 and mask_out 0
    y3, mo3 = conv(torch.randn(1, 1, 4, 4, device=dev), torch.zeros(1, 1, 4, 4, device=dev))
    assert float(y3.abs().max()) == 0.0 and float(mo3.abs().max()) == 0.0


def test_ratio_bit_exact(b2pose, dev, golden_dir):
    """KA2 through b2_pconv_mask_update: window/(count+1e-6)*clamp(count,0,1) in fp32, bit for bit."""
    import ctypes as C
    L = b2pose._lib
    g = np.load(golden_dir + "/ka.npz")
    for (win, cnt), want in zip(g["ka2_pairs"], g["ka2_ratio"]):
        k = int(round(win ** 0.5))
        m = torch.zeros(1, k, k, device=dev)
        m.view(-1)[:int(cnt)] = 1
        desc = L.ConvDesc(1, k, k, 1, 1, k, k, 1, 0, 1, 1, 1, L.F32, L.CONV_PARTIAL)
        mo = torch.empty(1, 1, 1, device=dev)
        ratio = torch.empty(1, 1, 1, device=dev)
        L.call("b2_pconv_mask_update", C.byref(desc), L.ptr(m), L.ptr(mo), L.ptr(ratio), L.stream())
        assert np.float32(ratio.item()) == np.float32(want), (win, cnt)
        assert mo.item() == (1.0 if cnt > 0 else 0.0)


def test_golden_cases_fp32(b2pose, dev, golden_dir):
    g = np.load(golden_dir + "/pconv.npz")
    for name, spec in zip(g["names"], g["specs"]):
        N, C, K, H, W, k, s, p, d, has_bias = [int(v) for v in spec]
        conv = _module(b2pose, dev, C, K, k, s, p, d, g[f"{name}_w"], g[f"{name}_b"] if has_bias else None)
        x = torch.tensor(g[f"{name}_x"], device=dev, requires_grad=True)
        y, mo = conv(x, torch.tensor(g[f"{name}_mask"], device=dev))
        (y * torch.tensor(g[f"{name}_cot"], device=dev)).sum().backward()
        assert mo.dtype == torch.float32 and tuple(mo.shape) == g[f"{name}_mask_out"].shape
        assert np.array_equal(mo.cpu().numpy(), g[f"{name}_mask_out"]), name          # bit-exact
        assert rel_err(y, g[f"{name}_out"]) < FP32_TOL, name
        assert rel_err(x.grad, g[f"{name}_dx"]) < FP32_TOL, name
        assert rel_err(conv.weight.grad, g[f"{name}_dw"]) < FP32_TOL, name
        if has_bias:
            assert rel_err(conv.bias.grad, g[f"{name}_db"]) < FP32_TOL, name
        assert np.all(x.grad.cpu().numpy()[np.broadcast_to(g[f"{name}_mask"] == 0, x.shape)] == 0), name


def test_golden_cases_bf16(b2pose, dev, golden_dir):
    g = np.load(golden_dir + "/pconv.npz")
    for name, spec in zip(g["names"], g["specs"]):
        N, C, K, H, W, k, s, p, d, has_bias = [int(v) for v in spec]
        conv = _module(b2pose, dev, C, K, k, s, p, d, g[f"{name}_w"], g[f"{name}_b"] if has_bias else None)
        x = torch.tensor(g[f"{name}_x"], device=dev).bfloat16()
        y, mo = conv(x, torch.tensor(g[f"{name}_mask"], device=dev))
        assert mo.dtype == torch.float32                                             # KA4
        assert y.dtype == (torch.float32 if has_bias else torch.bfloat16)
        assert np.array_equal(mo.cpu().numpy(), g[f"{name}_mask_out"]), name
        assert rel_err(y, g[f"{name}_out_bf16"]) < BF16_TOL, name
        assert rel_err(y, g[f"{name}_out"]) < BF16_TOL, name


SWEEP = [
    # N, C,   K,   H,  W,  k, s, p, d, invalid
    (2, 64, 64, 32, 32, 3, 1, 1, 1, 0.25),
    (2, 64, 256, 16, 16, 1, 1, 0, 1, 0.5),
    (2, 128, 128, 17, 17, 3, 2, 1, 1, 0.1),
    (1, 128, 128, 16, 16, 3, 1, 2, 2, 0.9),
    (2, 256, 64, 9, 9, 1, 1, 0, 1, 0.0),
    (3, 1, 64, 33, 33, 7, 2, 3, 1, 0.25),
    (1, 512, 128, 8, 8, 1, 1, 0, 1, 0.5),
    (1, 32, 48, 5, 7, 3, 1, 1, 1, 1.0),      # everything invalid
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", SWEEP)
def test_sweep_vs_oracle(b2pose, dev, case, dtype):
    N, C, K, H, W, k, s, p, d, inv = case
    gen = torch.Generator().manual_seed(hash(case) % (1 << 31))
    w = torch.randn(K, C, k, k, generator=gen) * (2.0 / (k * k * K)) ** 0.5
    x = torch.randn(N, C, H, W, generator=gen)
    m = po.blob_mask(N, max(H, W), min(inv, 0.95), gen)[:, :, :H, :W].contiguous() if inv < 1.0 \
        else torch.zeros(N, 1, H, W)
    if dtype == torch.bfloat16:                     # same rounded operands on both sides
        w, x = w.bfloat16().float(), x.bfloat16().float()
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    yr, mr = po.partial_conv(xr, m, wr, None, s, p, d)
    cot = torch.randn(yr.shape, generator=gen)
    if dtype == torch.bfloat16:
        cot = cot.bfloat16().float()
    (yr * cot).sum().backward()

    conv = _module(b2pose, dev, C, K, k, s, p, d, w)
    xg = x.to(dev).to(dtype).requires_grad_(True)
    yg, mg = conv(xg, m.to(dev))
    (yg.float() * cot.to(dev)).sum().backward()
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    assert torch.equal(mg.cpu(), mr)
    if float(yr.abs().max()) == 0:
        assert float(yg.abs().max()) == 0 and float(xg.grad.abs().max()) == 0
        return
    assert rel_err(yg, yr) < tol
    assert rel_err(xg.grad, xr.grad) < tol
    assert rel_err(conv.weight.grad, wr.grad) < tol


def test_plain_conv_and_bias(b2pose, dev):
    """The mask-disabled instantiation that serves every nn.Conv2d of the nets (regressor has a bias)."""
    gen = torch.Generator().manual_seed(5)
    for (C, K, k, s, p, d) in [(64, 272, 3, 1, 1, 1), (3, 64, 7, 2, 3, 1), (256, 512, 1, 2, 0, 1), (64, 64, 3, 1, 2, 2)]:
        conv = b2pose.Conv2d(C, K, kernel_size=k, stride=s, padding=p, dilation=d, bias=True).to(dev)
        x = torch.randn(2, C, 13, 13, generator=gen)
        ref = torch.nn.functional.conv2d
        xr = x.clone().requires_grad_(True)
        w = conv.weight.detach().cpu().contiguous().requires_grad_(True)
        b = conv.bias.detach().cpu().clone().requires_grad_(True)
        yr = ref(xr, w, b, s, p, d)
        cot = torch.randn(yr.shape, generator=gen)
        (yr * cot).sum().backward()
        xg = x.to(dev).requires_grad_(True)
        yg = conv(xg)
        (yg * cot.to(dev)).sum().backward()
        assert rel_err(yg, yr) < FP32_TOL
        assert rel_err(xg.grad, xr.grad) < FP32_TOL
        assert rel_err(conv.weight.grad, w.grad) < FP32_TOL
        assert rel_err(conv.bias.grad, b.grad) < FP32_TOL


def test_errors(b2pose, dev):
    conv = b2pose.PartialConv(4, 4, kernel_size=3, padding=1, bias=False)
    with pytest.raises(RuntimeError):                  # CPU tensors: no fallback
        conv(torch.randn(1, 4, 5, 5), torch.ones(1, 1, 5, 5))
    with pytest.raises(NotImplementedError):
        b2pose.PartialConv(4, 4, kernel_size=3, multi_channel=True)
    conv = conv.to(dev)
    with pytest.raises(ValueError):
        conv(torch.randn(1, 4, 5, 5, device=dev), torch.ones(1, 1, 4, 5, device=dev))
    out = b2pose.PartialConv(4, 4, kernel_size=3, padding=1, return_mask=False).to(dev)(
        torch.randn(1, 4, 5, 5, device=dev), torch.ones(1, 1, 5, 5, device=dev))
    assert torch.is_tensor(out)
