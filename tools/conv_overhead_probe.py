#!/usr/bin/env python
"""Separates the fixed per-launch cost of the tcgen05 convolution kernel from its work: times fprop of a few shapes
cold (L2 evicted before every launch), warm (20 launches back to back) and at batch 8 vs 64.  Run with
B2POSE_TC_DEBUG=9 (no operand loads, no epilogue: MMA issue + launch overhead only) and =0."""
import ctypes as C
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

b2 = ge.load_package()
L = b2._lib
dev = torch.device("cuda:0")
SHAPES = {"3x3_64": (64, 64, 64, 64, 3, 1, 1), "1x1_64_256": (64, 64, 64, 256, 1, 1, 0), "1x1_256_64": (64, 64, 256, 64, 1, 1, 0),
          "3x3_512": (16, 16, 512, 512, 3, 1, 1)}
flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


for name, (H, W, Cin, K, k, s, p) in SHAPES.items():
    for batch in (8, 64):
        d = b2.ops.make_desc((batch, H, W, Cin), K, k, k, s, p, 1, L.BF16, 0)
        x = torch.randn(batch, H, W, Cin, device=dev).bfloat16()
        w = (torch.randn(K, k, k, Cin, device=dev) * 0.05).bfloat16()
        y = torch.empty(batch, d.Ho, d.Wo, K, device=dev, dtype=torch.bfloat16)
        sums = torch.empty(L.BN_PARTS * 2 * K, dtype=torch.float32, device=dev)
        ws = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
        st = L.stream()
        fn = lambda: L.call("b2_pconv_fprop", C.byref(d), L.ptr(x), None, L.ptr(w), None, L.ptr(y), None, None,
                            L.ptr(sums), L.ptr(ws), ws.numel(), st)
        fn()
        torch.cuda.synchronize()
        cold = []
        for _ in range(5):
            flush.max()
            e0, e1 = ev(), ev()
            e0.record(); fn(); e1.record(); e1.synchronize()
            cold.append(e0.elapsed_time(e1) * 1e3)
        e0, e1 = ev(), ev()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); e1.synchronize()
        warm = e0.elapsed_time(e1) * 1e3 / 20
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        graph = e0.elapsed_time(e1) * 1e3 / 20
        print("dbg=%s %-11s batch %2d: cold %.1f us  warm back-to-back %.1f us  in a graph %.1f us" % (
            os.environ.get("B2POSE_TC_DEBUG", "0"), name, batch, statistics.median(cold), warm, graph), flush=True)
