#!/usr/bin/env python
"""BASELINE config 5: isolated PartialConv2d layer sweep (C = 64..2048, 8..128 px, mask sparsity 0-90 %), bf16,
batch 64: this repo's PartialConv module (tcgen05 kernels through the C ABI) next to an eager-PyTorch / cuDNN
restatement of partial_conv.py:32-58 on the same GPU.  Prints a markdown table (CUDA events, median of 5)."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import __graft_entry__ as ge  # noqa: E402


def eager_partial_conv(x, mask, w, k, pad, dil):
    """partial_conv.py:32-58 with torch ops (single-channel mask, no bias, return_mask=True)."""
    with torch.no_grad():
        ones = torch.ones(1, 1, k, k, device=x.device, dtype=mask.dtype)
        upd = F.conv2d(mask, ones, None, 1, pad, dil)
        ratio = (k * k) / (upd + 1e-6)
        upd = torch.clamp(upd, 0, 1)
        ratio = ratio * upd
    raw = F.conv2d(x * mask.to(x.dtype), w, None, 1, pad, dil)
    return raw * ratio.to(x.dtype), upd


def blob_mask(n, side, frac, gen):
    m = torch.ones(n, 1, side, side)
    lo, hi = max(1, side // 16), max(2, (side * 3) // 8)
    for i in range(n):
        guard = 0
        while float(1 - m[i].mean()) < frac and guard < 2000:
            h, w = (int(v) for v in torch.randint(lo, hi + 1, (2,), generator=gen))
            t, l = int(torch.randint(0, max(1, side - h + 1), (1,), generator=gen)), int(torch.randint(0, max(1, side - w + 1), (1,), generator=gen))
            m[i, :, t:t + h, l:l + w] = 0
            guard += 1
    return m


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def main():
    b2 = ge.load_package()
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(0)
    N = 64
    print("| C | side | k | invalid % | ours fwd ms | ours fwd+bwd ms | eager fwd ms | eager fwd+bwd ms | speed-up fwd+bwd | TFLOP/s (ours, fwd+bwd) | mask_out equal |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|:-:|")
    for Cc in (64, 128, 256, 512, 1024, 2048):
        for side in (8, 16, 32, 64, 128):
            if N * Cc * side * side > (1 << 28):
                continue
            for k in (1, 3):
                for frac in (0.0, 0.5, 0.9):
                    if frac and (Cc not in (64, 512) or side not in (16, 64)):
                        continue                     # sparsity does not change the work: sample it on a few shapes
                    pad = k // 2
                    x = torch.randn(N, Cc, side, side, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
                    mask = blob_mask(N, side, frac, gen).to(dev)
                    conv = b2.PartialConv(Cc, Cc, kernel_size=k, padding=pad, bias=False).to(dev).to(torch.bfloat16)
                    w = conv.weight.detach().clone().requires_grad_(True)
                    xo = x.clone().requires_grad_(True)
                    xe = x.clone().requires_grad_(True)

                    def ours_fwd():
                        with torch.no_grad():
                            return conv(x, mask)

                    def ours_fb():
                        y, _ = conv(xo, mask)
                        y.backward(y.detach())
                        conv.weight.grad = None
                        xo.grad = None

                    def eager_fwd():
                        with torch.no_grad():
                            return eager_partial_conv(x, mask, w, k, pad, 1)

                    def eager_fb():
                        y, _ = eager_partial_conv(xe, mask, w, k, pad, 1)
                        y.backward(y.detach())
                        w.grad = None
                        xe.grad = None

                    (yo, mo), (ye, me) = ours_fwd(), eager_fwd()
                    same = bool(torch.equal(mo.reshape(-1), me.reshape(-1)))
                    t = [timed(f) for f in (ours_fwd, ours_fb, eager_fwd, eager_fb)]
                    flops = 3 * 2.0 * N * side * side * Cc * Cc * k * k
                    print("| %d | %d | %d | %d | %.3f | %.3f | %.3f | %.3f | %.2fx | %.0f | %s |"
                          % (Cc, side, k, int(frac * 100), t[0], t[1], t[2], t[3], t[3] / t[1], flops / t[1] / 1e9,
                             "yes" if same else "NO"), flush=True)
                    del conv, x, xo, xe, w


if __name__ == "__main__":
    main()
