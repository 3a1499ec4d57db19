#!/usr/bin/env python
"""Launches fprop / dgrad / wgrad of every PartialConv shape of the bench workload ONCE (after one warm-up launch each)
between cudaProfilerStart/Stop, printing the launch order -- the target of

    ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\
dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_pconv_ncu.csv \
        python tools/pconv_ncu_probe.py

`tools/pconv_ncu_summary.py` joins the CSV with the printed order into profiles/r02_pconv_tensor_pipe.json and
profiles/traffic.json (the keys bench.py's `roofline` looks up)."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="partial_fusionnet")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_pconv_order.json"))
    a = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    b2 = ge.load_package()
    L = b2._lib
    dev = torch.device("cuda:0")
    args = argparse.Namespace(workload=a.workload, model="resnet50", side=256, joints=17, batch=a.batch)
    wl = bench.WORKLOADS[a.workload]
    cfg = b2.train_args(model="resnet50", num_joints=17, side_in=256, depth_only=not wl["fused"], do_fusion=wl["fused"],
                        half_acc=True)
    net = getattr(getattr(b2, wl["kind"]), "resnet50")(cfg, False).to(dev).train()
    shapes = {k: v for k, v in bench.conv_layer_table(b2, net, args, dev).items() if k[-1]}      # PartialConv layers only
    order, calls = [], []
    for key in shapes:
        H, W, Cin, K, R, S, stride, pad, dil, partial = key
        flags = L.CONV_PARTIAL | L.CONV_X_PREMASKED | L.CONV_DY_PRESCALED
        d = b2.ops.make_desc((a.batch, H, W, Cin), K, R, S, stride, pad, dil, L.BF16, flags)
        x = torch.randn(a.batch, H, W, Cin, device=dev).bfloat16()
        w = (torch.randn(K, R, S, Cin, device=dev) * 0.05).bfloat16()
        y = torch.empty(a.batch, d.Ho, d.Wo, K, device=dev, dtype=torch.bfloat16)
        dy = torch.randn(a.batch, d.Ho, d.Wo, K, device=dev).bfloat16()
        dx = torch.empty_like(x)
        dw = torch.zeros(K, R, S, Cin, device=dev)
        mask = (torch.rand(a.batch, H, W, device=dev) > 0.25).float()
        mo = torch.empty(a.batch, d.Ho, d.Wo, device=dev)
        ratio = torch.ones(a.batch, d.Ho, d.Wo, device=dev)
        sums = torch.empty(L.BN_PARTS * 2 * K, dtype=torch.float32, device=dev)
        wsb = max(L.lib().b2_conv_workspace_bytes(C.byref(d), op) for op in (0, 1, 2))
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
        st = L.stream()
        shape = "N%d %dx%dx%d->%d k%d s%d p%d d%d partial" % (a.batch, H, W, Cin, K, R, stride, pad, dil)
        keep = (d, x, w, y, dy, dx, dw, mask, mo, ratio, sums, ws)
        fns = {
            "fprop": lambda d=d, x=x, mask=mask, w=w, y=y, mo=mo, ratio=ratio, sums=sums, ws=ws: L.call(
                "b2_pconv_fprop", C.byref(d), L.ptr(x), L.ptr(mask), L.ptr(w), None, L.ptr(y), L.ptr(mo), L.ptr(ratio),
                L.ptr(sums), L.ptr(ws), ws.numel(), st),
            "dgrad": lambda d=d, dy=dy, w=w, mask=mask, dx=dx, ws=ws: L.call(
                "b2_pconv_dgrad", C.byref(d), L.ptr(dy), None, L.ptr(w), L.ptr(mask), L.ptr(dx), L.ptr(ws), ws.numel(), st),
            "wgrad": lambda d=d, x=x, mask=mask, dy=dy, dw=dw, ws=ws: L.call(
                "b2_pconv_wgrad", C.byref(d), L.ptr(x), L.ptr(mask), L.ptr(dy), None, L.ptr(dw), L.ptr(ws), ws.numel(), st),
        }
        for op, fn in fns.items():
            if op == "dgrad" and Cin <= 4:
                continue
            fn()                                  # warm-up outside the profiled range
            calls.append((op, shape, fn, keep))
    torch.cuda.synchronize()
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    torch.cuda.cudart().cudaProfilerStart()
    for op, shape, fn, _ in calls:
        flush.max()                               # evict the L2 between the measured launches (shows up as a reduce kernel)
        fn()
        order.append("conv_%s %s" % (op, shape))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    json.dump(order, open(a.out, "w"), indent=1)
    print("\n".join(order))


if __name__ == "__main__":
    main()
