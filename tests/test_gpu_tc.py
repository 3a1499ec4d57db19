"""GPU parity of the tcgen05 tensor-core convolution path (bf16): fprop, dgrad and wgrad through
the module API on shapes the dispatcher sends to the tensor cores, against the CPU oracle
evaluated in fp32 on the same bf16-rounded operands.  Tolerance: 2e-2 relative (BASELINE bf16)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import pose_oracle as po
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-2

CASES = [
    # name            N  C     K    H   W   k  s  p  d  partial premasked bias
    ("p3x3_64",       2, 64,   64,  32, 32, 3, 1, 1, 1, False, False, False),
    ("p3x3_d2",       3, 128,  128, 16, 16, 3, 1, 2, 2, False, False, False),
    ("p1x1_odd",      2, 256,  64,  9,  9,  1, 1, 0, 1, False, False, False),
    ("p3x3_s2",       2, 128,  128, 17, 17, 3, 2, 1, 1, False, False, False),
    ("p1x1_s2",       2, 256,  512, 16, 16, 1, 2, 0, 1, False, False, False),
    ("p1x1_s2_odd",   2, 64,   128, 33, 33, 1, 2, 0, 1, False, False, False),
    ("regressor",     2, 128,  272, 16, 16, 3, 1, 1, 1, False, False, True),
    ("pc1x1",         2, 64,   256, 16, 16, 1, 1, 0, 1, True,  False, False),
    ("pc3x3_pre",     2, 64,   64,  16, 16, 3, 1, 1, 1, True,  True,  False),
    ("pc3x3_s2_pre",  2, 128,  128, 32, 32, 3, 2, 1, 1, True,  True,  False),
    ("tiny_sp",       2, 512,  2048, 4, 4,  1, 1, 0, 1, False, False, False),
    ("ragged65",      1, 64,   64,  65, 65, 3, 1, 1, 1, False, False, False),
    ("deepK",         1, 2048, 512, 8,  8,  1, 1, 0, 1, False, False, False),
    ("l4_3x3",        2, 512,  512, 16, 16, 3, 1, 2, 2, False, False, False),
    ("s2_odd",        2, 128,  128, 33, 33, 3, 2, 1, 1, False, False, False),
    ("stem_rgb",      2, 3,    64,  64, 64, 7, 2, 3, 1, False, False, False),
    ("stem_depth_pc", 2, 1,    64,  65, 65, 7, 2, 3, 1, True,  False, False),
    ("c_tail",        2, 72,   80,  12, 12, 3, 1, 1, 1, False, False, False),
    ("stem_rgb_odd",  2, 3,    64,  65, 65, 7, 2, 3, 1, False, False, False),
    ("stem_c4",       1, 4,    64,  32, 32, 7, 2, 3, 1, False, False, False),
    # stems at widths that fill / exceed / straddle the 128-pixel output-row tiles of the row-stem kernel
    ("stem_depth_256", 2, 1,   64,  256, 256, 7, 2, 3, 1, True,  True,  False),
    ("stem_rgb_wide", 1, 3,    64,  38, 300, 7, 2, 3, 1, False, False, False),
    ("stem_pc_257",   1, 1,    64,  41, 257, 7, 2, 3, 1, True,  False, False),
    # 3x3 s1 p1, C = 64 at the widths of layer1 / layer5 (W % 64 == 0); K = 64 (second dY atom never loaded) and K = 128
    ("wg_halo_64",    2, 64,   64,  20, 64, 3, 1, 1, 1, False, False, False),
    ("wg_halo_128pc", 1, 64,   128, 9, 128, 3, 1, 1, 1, True,  True,  False),
    ("smallc_5x5",    2, 3,    32,  20, 20, 5, 1, 2, 1, False, False, False),
    # halo-tile mode (3x3 stride 1, Ho % 16 == 0, Wo % 8 == 0): several tiles per image, two channel blocks
    ("halo_48x40",    3, 128,  64,  48, 40, 3, 1, 1, 1, False, False, False),
    ("halo_pc_d2",    2, 64,   128, 32, 24, 3, 1, 2, 2, True,  True,  False),
]


def _uses_tc(b2pose, x_shape, K, k, s, p, d, flags):
    L = b2pose._lib
    desc = b2pose.ops.make_desc(x_shape, K, k, k, s, p, d, L.BF16, flags)
    return [L.lib().b2_conv_uses_tensor_cores(C.byref(desc), op) for op in (0, 1, 2)]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tc_conv(b2pose, dev, case):
    name, N, Cin, K, H, W, k, s, p, d, partial, premasked, bias = case
    L = b2pose._lib
    # poison every scratch buffer (0xFF bytes are NaNs in bf16 and fp32): a kernel that reads workspace memory it did not
    # write (a TMA box running past a padded row, an unwritten staging tile) must not get away with zeros
    for ws in b2pose.ops._workspaces.values():
        ws.fill_(0xFF)
    torch.empty(64 << 20, dtype=torch.uint8, device=dev).fill_(0xFF)        # and what the allocator hands out next
    gen = torch.Generator().manual_seed(abs(hash(name)) % (1 << 31))
    bf = lambda t: t.bfloat16().float()
    w = bf(torch.randn(K, Cin, k, k, generator=gen) * (2.0 / (k * k * K)) ** 0.5)
    b = torch.randn(K, generator=gen) * 0.1 if bias else None
    x = bf(torch.randn(N, Cin, H, W, generator=gen))
    m = po.blob_mask(N, max(H, W), 0.35, gen)[:, :, :H, :W].contiguous() if partial else None
    if premasked:
        x = x * m
    flags = (L.CONV_PARTIAL if partial else 0) | (L.CONV_X_PREMASKED if premasked else 0)
    use = _uses_tc(b2pose, (N, H, W, Cin), K, k, s, p, d, flags)
    assert use[0] == 1 and use[2] == 1, use            # these shapes must run on tcgen05
    assert use[1] == (0 if Cin <= 4 else 1), use       # (network inputs have no dgrad on the tensor cores)

    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True) if bias else None
    if partial:
        yr, mr = po.partial_conv(xr, m, wr, br, s, p, d)
    else:
        yr, mr = F.conv2d(xr, wr, br, s, p, d), None
    cot = bf(torch.randn(yr.shape, generator=gen))
    (yr * cot).sum().backward()

    Conv = b2pose.PartialConv if partial else b2pose.Conv2d
    conv = Conv(Cin, K, kernel_size=k, stride=s, padding=p, dilation=d, bias=bias).to(dev)
    with torch.no_grad():
        conv.weight.copy_(w)
        if bias:
            conv.bias.copy_(b)
    xg = x.permute(0, 2, 3, 1).contiguous().to(dev).bfloat16().requires_grad_(True)      # NHWC
    if partial:
        yg, mg = conv.forward_nhwc(xg, m[:, 0].contiguous().to(dev), premasked=premasked)
        assert torch.equal(mg.cpu(), mr[:, 0])
    else:
        yg = conv.forward_nhwc(xg)
    (yg.float() * cot.permute(0, 2, 3, 1).to(dev)).sum().backward()
    errs = dict(y=rel_err(yg.permute(0, 3, 1, 2), yr), dw=rel_err(conv.weight.grad, wr.grad))
    dx_ref = xr.grad * m if premasked else xr.grad     # premasked flow: the mask is applied upstream
    dxg = xg.grad.permute(0, 3, 1, 2).float().cpu()
    errs["dx"] = rel_err(dxg * m if premasked else dxg, dx_ref)
    if bias:
        errs["db"] = rel_err(conv.bias.grad, br.grad)
    print(name, "tc(fprop,dgrad,wgrad)=", use, {k_: "%.2e" % v for k_, v in errs.items()})
    bad = {k_: v for k_, v in errs.items() if not v < TOL}
    assert not bad, (name, errs)


def test_tc_matches_ffma_bits_of_mask_and_ratio(b2pose, dev):
    """The tensor-core epilogue computes mask_out / ratio exactly like the CUDA-core kernel."""
    L = b2pose._lib
    gen = torch.Generator().manual_seed(4)
    N, Cin, K, H, W = 2, 64, 64, 24, 24
    x = torch.randn(N, H, W, Cin, generator=gen).to(dev).bfloat16()
    m = po.blob_mask(N, H, 0.4, gen)[:, 0].contiguous().to(dev)
    x = x * m.unsqueeze(-1).bfloat16()
    w = (torch.randn(K, 3, 3, Cin, generator=gen) * 0.05).to(dev).bfloat16()
    outs = []
    for force in (0, L.CONV_FORCE_FFMA):
        desc = b2pose.ops.make_desc((N, H, W, Cin), K, 3, 3, 1, 1, 1, L.BF16,
                                    L.CONV_PARTIAL | L.CONV_X_PREMASKED | force)
        y, mo, ratio, _ = b2pose.ops._conv_fprop(desc, x, m, w, None, True, False)
        outs.append((y, mo, ratio))
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    assert rel_err(outs[0][0], outs[1][0]) < 1e-2


def test_dgrad_reduce_add(b2pose, dev):
    """B2_CONV_DX_ACCUMULATE: dx += dgrad(...) through the TMA reduce-add epilogue equals a separate
    bf16 add (this folds the identity-shortcut gradient of a residual block into its first conv)."""
    L = b2pose._lib
    gen = torch.Generator().manual_seed(12)
    for (N, H, W, Cin, K, k, p, d, partial) in [(2, 16, 16, 256, 64, 1, 0, 1, False), (2, 16, 16, 64, 64, 3, 1, 1, False),
                                                (2, 12, 12, 128, 128, 3, 2, 2, False), (2, 16, 16, 256, 64, 1, 0, 1, True)]:
        flags = L.CONV_PARTIAL if partial else 0
        desc = b2pose.ops.make_desc((N, H, W, Cin), K, k, k, 1, p, d, L.BF16, flags)
        dy = torch.randn(N, desc.Ho, desc.Wo, K, generator=gen).to(dev).bfloat16()
        w = (torch.randn(K, k, k, Cin, generator=gen) * 0.05).to(dev).bfloat16()
        mask = (torch.rand(N, H, W, generator=gen) > 0.3).float().to(dev) if partial else None
        ratio = torch.rand(N, desc.Ho, desc.Wo, generator=gen).to(dev) if partial else None
        addend = torch.randn(N, H, W, Cin, generator=gen).to(dev).bfloat16()
        plain = b2pose.ops._conv_dgrad(desc, dy, ratio, w, mask)
        fused = b2pose.ops._conv_dgrad(desc, dy, ratio, w, mask, addend.clone())
        want = plain + addend
        assert rel_err(fused, want) < 4e-3, (Cin, K, k)           # one bf16 rounding at most
    # a shape the reduce-add epilogue does not take falls back to a separate add
    desc = b2pose.ops.make_desc((2, 17, 17, 128), 128, 3, 3, 2, 1, 1, L.BF16, 0)
    dy = torch.randn(2, desc.Ho, desc.Wo, 128, generator=gen).to(dev).bfloat16()
    w = (torch.randn(128, 3, 3, 128, generator=gen) * 0.05).to(dev).bfloat16()
    addend = torch.randn(2, 17, 17, 128, generator=gen).to(dev).bfloat16()
    got = b2pose.ops._conv_dgrad(desc, dy, None, w, None, addend.clone())
    assert rel_err(got, b2pose.ops._conv_dgrad(desc, dy, None, w, None) + addend) < 4e-3


@pytest.mark.parametrize("cin,k,ksz,pad", [(512, 512, 1, 0), (256, 64, 3, 1)])
def test_concat_input_matches_materialised_cat(b2pose, dev, cin, k, ksz, pad):
    """B2_CONV_X_CONCAT (the fusion unit, fusionnet.py:137): conv + BN + ReLU over [x, x2] read through two tensor maps
    must equal the same layer over torch.cat([x, x2]) -- output, both input gradients, filter and BatchNorm gradients."""
    from b2pose.layers import conv_bn
    gen = torch.Generator().manual_seed(31)
    N, H, W = 3, 24, 20
    conv = b2pose.Conv2d(2 * cin, k, kernel_size=ksz, padding=pad, bias=False).to(dev)
    bn = b2pose.BatchNorm2d(k).to(dev)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=gen) * (1.0 / (2 * cin * ksz * ksz)) ** 0.5)
        bn.weight.copy_(torch.rand(k, generator=gen) + 0.5)
        bn.bias.copy_(torch.randn(k, generator=gen) * 0.1)
    a = torch.randn(N, H, W, cin, generator=gen).to(dev).bfloat16()
    b = torch.randn(N, H, W, cin, generator=gen).to(dev).bfloat16()
    cot = torch.randn(N, H, W, k, generator=gen).to(dev).bfloat16()
    L = b2pose._lib
    desc = b2pose.ops.make_desc((N, H, W, 2 * cin), k, ksz, ksz, 1, pad, 1, L.BF16, L.CONV_X_CONCAT)
    assert all(L.lib().b2_conv_uses_tensor_cores(C.byref(desc), op) == 1 for op in (0, 1, 2))
    outs = []
    for split in (True, False):
        conv.weight.grad = bn.weight.grad = bn.bias.grad = None
        xa, xb = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        if split:
            z, _ = conv_bn(xa, None, conv, bn, relu=True, x2=xb)
        else:
            z, _ = conv_bn(torch.cat([xa, xb], dim=3), None, conv, bn, relu=True)
        (z.float() * cot.float()).sum().backward()
        outs.append((z.detach(), xa.grad, xb.grad, conv.weight.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone()))
    for name, got, want in zip(("z", "dx", "dx2", "dw", "dgamma", "dbeta"), outs[0], outs[1]):
        e = rel_err(got, want)
        print("concat %dx%d k%d %s: %.2e" % (2 * cin, k, ksz, name, e))
        assert e < 5e-3, (name, e)           # same MMAs; only the order of the fp32 atomics (statistics, dW) differs
