#!/usr/bin/env python
"""Times the four BatchNorm stream kernels alone (CUDA events, L2 evicted by a 256 MB read) on the
activation shapes that dominate the step, and prints achieved HBM GB/s (algorithmic bytes)."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import __graft_entry__ as ge
    b2 = ge.load_package()
    L = b2._lib
    dev = torch.device("cuda:0")
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    shapes = [(64 * 64 * 64, 64), (64 * 64 * 64, 256), (64 * 32 * 32, 512), (64 * 16 * 16, 2048)]
    for rows, C in shapes:
        bf = torch.bfloat16
        y = torch.randn(rows, C, device=dev).to(bf)
        z = torch.empty_like(y)
        res = torch.randn(rows, C, device=dev).to(bf)
        dz = torch.randn(rows, C, device=dev).to(bf)
        dy = torch.empty_like(y)
        dres = torch.empty_like(y)
        gamma = torch.ones(C, device=dev)
        beta = torch.zeros(C, device=dev)
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        mean, invstd = torch.empty(C, device=dev), torch.empty(C, device=dev)
        parts = torch.empty(L.BN_PARTS * 2 * C, dtype=torch.float32, device=dev)
        parts2 = torch.empty_like(parts)
        gsum = torch.zeros(2 * C, device=dev)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        st = L.stream()
        esz = 2
        n = rows * C * esz
        P = L.ptr
        calls = {
            "stats": (lambda: L.call("b2_bn_stats", P(y), rows, C, L.BF16, P(parts), st), n),
            "finalize": (lambda: L.call("b2_bn_finalize", P(parts), rows, C, P(rm), P(rv), 0.1, 1e-5, 1, P(mean), P(invstd), st), 0),
            "apply": (lambda: L.call("b2_bn_apply", P(y), P(mean), P(invstd), P(gamma), P(beta), None, None, 1, P(z), rows, C,
                                     L.BF16, st), 2 * n),
            "apply+res": (lambda: L.call("b2_bn_apply", P(y), P(mean), P(invstd), P(gamma), P(beta), P(res), None, 1, P(z), rows,
                                         C, L.BF16, st), 3 * n),
            "bwd_reduce(regate)": (lambda: L.call("b2_bn_bwd_reduce", P(dz), None, P(y), P(mean), P(invstd), P(gamma), P(beta),
                                                  None, 1, P(parts2), rows, C, L.BF16, st), 2 * n),
            "bwd_reduce(z)": (lambda: L.call("b2_bn_bwd_reduce", P(dz), P(z), P(y), P(mean), P(invstd), P(gamma), P(beta),
                                             None, 1, P(parts2), rows, C, L.BF16, st), 3 * n),
            "bwd_finalize": (lambda: L.call("b2_bn_bwd_finalize", P(parts2), C, P(gsum), P(dg), P(db), st), 0),
            "bwd_apply(regate)": (lambda: L.call("b2_bn_bwd_apply", P(dz), None, P(y), P(mean), P(invstd), P(gamma), P(beta),
                                                 P(gsum), None, None, 1, 1, P(dy), None, rows, C, L.BF16, st), 3 * n),
            "bwd_apply(z,dres)": (lambda: L.call("b2_bn_bwd_apply", P(dz), P(z), P(y), P(mean), P(invstd), P(gamma), P(beta),
                                                 P(gsum), None, None, 1, 1, P(dy), P(dres), rows, C, L.BF16, st), 5 * n),
        }
        line = []
        for name, (fn, nbytes) in calls.items():
            fn()
            ts = []
            for _ in range(5):
                flush.max()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = statistics.median(ts)
            line.append("%s %.0fus %.0fGB/s" % (name, ms * 1e3, nbytes / ms / 1e6))
        print("rows=%d C=%d | " % (rows, C) + " | ".join(line))


if __name__ == "__main__":
    main()
