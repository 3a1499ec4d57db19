#!/usr/bin/env python
"""Launches fprop / dgrad / wgrad of a few named convolution shapes once each (after one warm-up
launch) -- the target of `ncu --set full -k regex:...`; also prints CUDA-event timings when run plain."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

SHAPES = {
    # name: (H, W, C, K, R, S, stride, pad, dil, partial)
    "l1_1x1_64_256": (64, 64, 64, 256, 1, 1, 1, 0, 1, False),
    "l1_3x3_64": (64, 64, 64, 64, 3, 3, 1, 1, 1, False),
    "l1_1x1_256_64": (64, 64, 256, 64, 1, 1, 1, 0, 1, False),
    "l2_1x1_128_512": (32, 32, 128, 512, 1, 1, 1, 0, 1, False),
    "l3_3x3_256": (16, 16, 256, 256, 3, 3, 1, 1, 1, False),
    "l4_3x3_512": (16, 16, 512, 512, 3, 3, 1, 2, 2, False),
    "l4_1x1_512_2048": (16, 16, 512, 2048, 1, 1, 1, 0, 1, False),
    "regressor": (16, 16, 2048, 272, 3, 3, 1, 1, 1, False),
    "stem_rgb": (256, 256, 3, 64, 7, 7, 2, 3, 1, False),
    "pc_3x3_64": (64, 64, 64, 64, 3, 3, 1, 1, 1, True),
    "pc_1x1_64_256": (64, 64, 64, 256, 1, 1, 1, 0, 1, True),
    "pc_1x1_256_64": (64, 64, 256, 64, 1, 1, 1, 0, 1, True),
    "pc_3x3_128": (32, 32, 128, 128, 3, 3, 1, 1, 1, True),
    "pc_1x1_128_512": (32, 32, 128, 512, 1, 1, 1, 0, 1, True),
    "pc_stem": (256, 256, 1, 64, 7, 7, 2, 3, 1, True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default=",".join(SHAPES))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=1)
    a = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    b2 = ge.load_package()
    dev = torch.device("cuda:0")
    shapes = {SHAPES[n]: 1 for n in a.shapes.split(",")}
    rows = bench.time_conv_kernels(b2, shapes, a.batch, "bf16", dev, reps=a.reps)
    for r in rows:
        print("%-6s %-50s %.3f ms  %7.1f TF/s %7.0f GB/s tc=%d" % (r["op"], r["shape"], r["ms"], r["tflops"], r["gbs"], r["tc"]))


if __name__ == "__main__":
    main()
