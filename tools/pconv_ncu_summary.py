#!/usr/bin/env python
"""Joins the ncu CSV of tools/pconv_ncu_probe.py with its launch order: per PartialConv op the tensor-pipe activity of its
tcgen05 kernel, the summed duration and DRAM bytes of all its kernels.  Writes profiles/r02_pconv_tensor_pipe.json,
profiles/traffic.json and a markdown table."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csv_path, order_path = sys.argv[1], sys.argv[2]
order = json.load(open(order_path))
rows, hdr = [], None
for r in csv.reader(open(csv_path)):
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        rows.append(dict(zip(hdr, r)))
launches = {}
for d in rows:
    launches.setdefault(int(d["ID"]), {"name": d["Kernel Name"]})[d["Metric Name"]] = (float(d["Metric Value"].replace(",", "")), d["Metric Unit"])
seq = [launches[k] for k in sorted(launches)]
# split at the L2-eviction reduce kernels: one group of kernels per measured op
groups, cur = [], None
for l in seq:
    if "reduce_kernel" in l["name"] or "ReduceOp" in l["name"]:
        if cur is not None:
            groups.append(cur)
        cur = []
    elif cur is not None:
        cur.append(l)
if cur:
    groups.append(cur)
groups = [g for g in groups if g]
assert len(groups) == len(order), (len(groups), len(order))


def val(l, key, scale=1.0):
    v, u = l.get(key, (0.0, ""))
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "nsecond": 1e-3}.get(u, 1.0)
    return v * mult * scale


pipe, traffic, table = {}, {}, []
for name, g in zip(order, groups):
    main = max(g, key=lambda l: val(l, "gpu__time_duration.sum"))
    tc = [l for l in g if "conv_tc_kernel" in l["name"] or "wgrad_tc_kernel" in l["name"]]
    main = tc[0] if tc else main
    us = sum(val(l, "gpu__time_duration.sum") for l in g)
    dram = sum(val(l, "dram__bytes_read.sum") + val(l, "dram__bytes_write.sum") for l in g)
    p = main.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", (None, ""))[0]
    pipe[name] = p
    traffic[name] = {"dram_bytes": dram, "kernels": len(g), "ncu_us": us}
    table.append((name, len(g), us, val(main, "gpu__time_duration.sum"), p, dram / 1e6))
json.dump(pipe, open(os.path.join(ROOT, "profiles", "r02_pconv_tensor_pipe.json"), "w"), indent=1)
tj = os.path.join(ROOT, "profiles", "traffic.json")
old = {}
try:
    old = json.load(open(tj))
except (OSError, ValueError):
    pass
old.update(traffic)
json.dump(old, open(tj, "w"), indent=1)
print("| op | kernels | total us (ncu) | tcgen05 kernel us | tensor pipe active % | DRAM MB (read + write) |\n|---|---:|---:|---:|---:|---:|")
for t in table:
    print("| `%s` | %d | %.1f | %.1f | %s | %.1f |" % (t[0], t[1], t[2], t[3], "%.1f" % t[4] if t[4] is not None else "-", t[5]))
