#!/usr/bin/env python
"""Benchmark of the depth-stream hot path: training samples/s of the two-stream PartialConv ResNet-50
(BASELINE.json configs[2] shape at north_star's 256x256 / J=17: partial_fusionnet with the fixed stems, batch 64 per
GPU, bf16, synthetic RGB + depth + validity holes) -- forward, head, loss, backward, clip-norm, Adam -- on N B200s of
one node.  `--workload fusionnet` selects configs[1] (plain two-stream net, no PartialConv layers).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the reference algorithm on the host CPU cores

Prints ONE JSON line (rank 0).  `value` times K steps with the batch resident in HBM; `e2e`
times K steps through Trainer.train_step with pinned HOST batches (H2D inside, loss read back
every step); `roofline` times the dominant convolution kernel alone with CUDA events;
`cpu_baseline` is the reference's own modules (imported unmodified from baseline/_ref) stepped on the host cores
on a bounded sample; `gpu_eager_baseline` the same imported modules on this B200 in eager PyTorch (cuDNN).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "training samples/sec", "samples/s"
WORKLOADS = {
    "fusionnet": dict(kind="fusionnet", fused=True, desc="fusionnet ResNet-50 two-stream RGB+depth"),
    "partial_fusionnet": dict(kind="partial_fusionnet", fused=True,
                              desc="partial_fusionnet ResNet-50 (PartialConv depth stream, fixed stems)"),
    "partial_depthnet": dict(kind="partial_depthnet", fused=False, desc="partial_depthnet ResNet-50 (PartialConv)"),
    "depthnet": dict(kind="depthnet", fused=False, desc="depthnet ResNet-50 depth-only"),
}
# forward GFLOP per sample at 256x256, J=17; training = 3x
FWD_GFLOP = {"fusionnet": 24.39, "partial_fusionnet": 24.39, "partial_depthnet": 18.57, "depthnet": 18.57}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="partial_fusionnet", choices=sorted(WORKLOADS))
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--side", type=int, default=256)
    ap.add_argument("--joints", type=int, default=17)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--cpu-batch", type=int, default=8, help="batch of the CPU baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the eager-PyTorch reference on the GPU")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--layers", action="store_true", help="also write the per-layer kernel table to profiles/")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------- reference arms
def _ref_cfg(args):
    from types import SimpleNamespace
    wl = WORKLOADS[args.workload]
    return SimpleNamespace(stride=16, depth=16, num_joints=args.joints, side_in=args.side, depth_only=not wl["fused"],
                           early_dist=False, skip_relu=False, extra_channel=False, joint_space=False, pretrain=False)


def _host_batch(args, batch, seed=1):
    """The same synthetic batch generator the product arm uses (b2pose.synthetic_batch), on the host."""
    import __graft_entry__ as ge
    b2 = ge.load_package()
    return b2.synthetic_batch(batch, args.side, args.joints, None, seed=seed, invalid_frac=0.25)


def cpu_step_rate(args, steps, warmup, batch):
    """The reference training step on the host cores, all threads.  The reference's own modules
    (partial_fusionnet.py / utils.py ..., imported unmodified from baseline/_ref) when they are staged --
    kind "reference" -- else the oracle port (oracle/pose_oracle.py StepOracle) -- kind "port".
    Returns (samples/s, cores, description, ms per step, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import torch
    import ref_step
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = WORKLOADS[args.workload]
    if ref_step.available():
        ref = ref_step.import_reference()
        torch.manual_seed(0)
        orc = ref_step.RefStep(ref, wl["kind"], args.model, _ref_cfg(args), torch.device("cpu"), args.joints - 1)
        data = _host_batch(args, batch)
        kind = "reference"
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pose_oracle as po
        cfg = po.net_config(side_in=args.side, num_joints=args.joints)
        sd = po.init_state(wl["kind"], args.model, cfg, seed=0)
        orc = po.StepOracle(sd, wl["kind"], args.model, cfg, key_index=args.joints - 1)
        data = po.synth_batch(batch, args.side, args.joints, seed=1, invalid_frac=0.25)
        kind = "port"
    for _ in range(warmup):
        orc.step(data)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.step(data)
    dt = time.perf_counter() - t0
    sample = "%d steps of batch %d (fp32, %s %s %dx%d J=%d, %s) after %d warm-up" % (
        steps, batch, wl["kind"], args.model, args.side, args.side, args.joints,
        "reference modules imported from baseline/_ref" if kind == "reference" else "oracle port", warmup)
    return batch * steps / dt, cores, sample, dt / steps * 1e3, kind


def gpu_eager_rates(args, dev, steps=5, warmup=2):
    """BASELINE.md section 4's "bar to beat": the reference's own modules in eager PyTorch (cuDNN / ATen) on this
    B200, same workload and per-GPU batch, forward + head + loss + backward + clip + Adam per step, batch resident
    in HBM, CUDA events.  Variants: fp32 as shipped (NCHW), bf16 autocast NCHW, bf16 autocast channels_last."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import torch
    import ref_step
    if not ref_step.available():
        return {"unavailable": "baseline/_ref is not staged (run python baseline/stage_reference.py where /root/reference exists)"}
    ref = ref_step.import_reference()
    wl = WORKLOADS[args.workload]
    data = tuple(t.to(dev) for t in _host_batch(args, args.batch))
    out = {"unit": UNIT, "batch": args.batch, "steps": steps, "warmup": warmup,
           "what": "reference modules imported unmodified from baseline/_ref (partial_fusionnet with the documented "
                   "2-line stem fix), eager PyTorch %s on this GPU, step = depth_train.py:384-456" % torch.__version__}
    variants = (("fp32_nchw", None, False), ("bf16_autocast_nchw", torch.bfloat16, False),
                ("bf16_autocast_channels_last", torch.bfloat16, True))
    for name, ac, cl in variants:
        try:
            torch.manual_seed(0)
            st = ref_step.RefStep(ref, wl["kind"], args.model, _ref_cfg(args), dev, args.joints - 1, autocast=ac,
                                  channels_last=cl)
            for _ in range(warmup):
                st.step(data)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                loss = st.step(data)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": args.batch / ms * 1e3, "ms_per_step": ms, "loss": float(loss)}
        except Exception as e:      # noqa: BLE001  (a variant the stock code cannot run is reported, not fatal)
            out[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        finally:
            st = None
            torch.cuda.empty_cache()
    return out


def workload_config(args, world):
    wl = WORKLOADS[args.workload]
    return {"workload": "%s, %s, batch %d/GPU, %dx%d, J=%d, D=16, stride 16, fwd+head+loss+bwd+clip+Adam"
                        % (wl["desc"], args.model, args.batch, args.side, args.side, args.joints),
            "parallelism": "dp%d" % world}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(1, min(args.warmup, 2))
    # bounded: a step of batch 8 takes ~2-4 s on the box's cores; cap total work at a few minutes
    steps = min(steps, 12)
    value, cores, sample, ms, kind = cpu_step_rate(args, steps, warm, args.cpu_batch)
    how = ("the reference's own modules imported unmodified from baseline/_ref" if kind == "reference"
           else "oracle port of the reference step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args, args.gpus), sample="CPU (%s, fp32, all host threads): each step = one "
                       "batch of %d of the same workload" % (how, args.cpu_batch)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------- dominant-kernel roofline
def conv_layer_table(b2, net, args, dev):
    """Unique convolution shapes of the net (fprop geometry) with multiplicity, via forward hooks
    replaced by a dry shape walk: run one tiny-batch forward and record every conv_bn / ConvFn call."""
    import torch
    shapes = {}
    ops = b2.ops
    orig = ops.make_desc

    def rec(xshape, K, R, S, stride, pad, dil, dtype, flags):
        d = orig(xshape, K, R, S, stride, pad, dil, dtype, flags)
        key = (d.H, d.W, d.C, d.K, d.R, d.S, d.stride, d.pad, d.dil, bool(flags & 1))
        shapes[key] = shapes.get(key, 0) + 1
        return d
    ops.make_desc = rec
    try:
        with torch.no_grad():
            color = torch.zeros(1, 3, args.side, args.side, device=dev)
            depth = torch.ones(1, 1, args.side, args.side, device=dev)
            net(color, depth) if getattr(net, "fused", False) else net(depth if net.conv1.in_channels == 1 else color)
    finally:
        ops.make_desc = orig
    return shapes


def time_conv_kernels(b2, shapes, batch, dtype, dev, reps=5):
    """Times fprop / dgrad / wgrad of every unique conv shape alone (CUDA events on the launch
    stream, L2 flushed between reps by a 256 MB memset).  Returns rows sorted by total time."""
    import ctypes as C
    import torch
    L = b2._lib
    rows = []
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)    # 256 MB > 126 MB L2; READ to evict (clean lines)
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    for key, count in shapes.items():
        H, W, Cin, K, R, S, stride, pad, dil, partial = key
        flags = (L.CONV_PARTIAL | L.CONV_X_PREMASKED | L.CONV_DY_PRESCALED) if partial else 0
        d = b2.ops.make_desc((batch, H, W, Cin), K, R, S, stride, pad, dil, L.BF16 if dtype == "bf16" else L.F32, flags)
        x = torch.randn(batch, H, W, Cin, device=dev).to(tdt)
        w = (torch.randn(K, R, S, Cin, device=dev) * 0.05).to(tdt)
        y = torch.empty(batch, d.Ho, d.Wo, K, device=dev, dtype=tdt)
        dy = torch.randn(batch, d.Ho, d.Wo, K, device=dev).to(tdt)
        dx = torch.empty_like(x)
        dw = torch.zeros(K, R, S, Cin, device=dev)
        mask = torch.ones(batch, H, W, device=dev) if partial else None
        mo = torch.empty(batch, d.Ho, d.Wo, device=dev) if partial else None
        ratio = torch.ones(batch, d.Ho, d.Wo, device=dev) if partial else None
        sums = torch.empty(L.BN_PARTS * 2 * K, dtype=torch.float32, device=dev)
        wsb = max(L.lib().b2_conv_workspace_bytes(C.byref(d), op) for op in (0, 1, 2))
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
        st = L.stream()
        calls = {
            "fprop": lambda: L.call("b2_pconv_fprop", C.byref(d), L.ptr(x), L.ptr(mask), L.ptr(w), None, L.ptr(y),
                                    L.ptr(mo), L.ptr(ratio), L.ptr(sums), L.ptr(ws), ws.numel(), st),
            "dgrad": lambda: L.call("b2_pconv_dgrad", C.byref(d), L.ptr(dy), None, L.ptr(w), L.ptr(mask), L.ptr(dx),
                                    L.ptr(ws), ws.numel(), st),
            "wgrad": lambda: L.call("b2_pconv_wgrad", C.byref(d), L.ptr(x), L.ptr(mask), L.ptr(dy), None, L.ptr(dw),
                                    L.ptr(ws), ws.numel(), st),
        }
        flops = 2.0 * batch * d.Ho * d.Wo * K * Cin * R * S
        esz = 2 if dtype == "bf16" else 4
        bytes_io = {"fprop": (x.numel() + y.numel() + w.numel()) * esz,
                    "dgrad": (dy.numel() + dx.numel() + w.numel()) * esz,
                    "wgrad": (x.numel() + dy.numel()) * esz + dw.numel() * 4}
        for op, fn in calls.items():
            if op == "dgrad" and Cin <= 4:
                continue                       # network inputs need no gradient
            fn()
            ts = []
            for _ in range(reps):
                flush.max()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = statistics.median(ts)
            rows.append(dict(op=op, shape="N%d %dx%dx%d->%d k%d s%d p%d d%d%s" % (batch, H, W, Cin, K, R, stride, pad,
                                                                                 dil, " partial" if partial else ""),
                             count=count, ms=ms, total_ms=ms * count, tflops=flops / ms / 1e9,
                             gbs=bytes_io[op] / ms / 1e6, flops=flops, bytes=bytes_io[op],
                             tc=bool(L.lib().b2_conv_uses_tensor_cores(C.byref(d), {"fprop": 0, "dgrad": 1, "wgrad": 2}[op]))))
    rows.sort(key=lambda r: -r["total_ms"])
    return rows


# ----------------------------------------------------------------------------- own arm
def run_b200(args):
    import torch
    import __graft_entry__ as ge
    b2 = ge.load_package()
    from b2pose import parallel as P
    rank, local, world = P.init_from_env("nccl")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback); use --impl reference for CPU"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world != args.gpus and rank == 0:
        print("note: WORLD_SIZE=%d but --gpus %d; using WORLD_SIZE" % (world, args.gpus), file=sys.stderr)
    import torch.distributed as dist
    wl = WORKLOADS[args.workload]
    cfg = b2.train_args(model=args.model, num_joints=args.joints, side_in=args.side, stride=16, depth=16,
                        depth_only=not wl["fused"], do_fusion=wl["fused"], half_acc=args.dtype == "bf16",
                        batch_size=args.batch)
    torch.manual_seed(0)
    net = getattr(getattr(b2, wl["kind"]), args.model)(cfg, False).to(dev).train()
    trainer = b2.Trainer(cfg, net, dict(key_index=args.joints - 1), use_graph=not args.no_graph)
    trainer.adapt_learn_rate(2)
    host = b2.synthetic_batch(args.batch, args.side, args.joints, None, seed=1 + rank, invalid_frac=0.25, pin=True)
    resident = tuple(t.to(dev) for t in host)
    h2d = sum(t.numel() * t.element_size() for t in host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # graph capture (3 eager steps + capture) is set-up, then the W warm-up steps the contract asks for
    for _ in range(4 if trainer.use_graph else 1):
        trainer.train_step(resident)
    L = b2._lib
    for _ in range(max(args.warmup, 3)):
        out = trainer.train_step(resident)
    barrier()

    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    losses = []

    def timed(batch, read_back, steps=None):
        steps = args.steps if steps is None else steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        last = None
        pending = None
        if read_back:
            trainer.prefetch(batch)                 # H2D of step 0 (inside the timed region)
        for i in range(steps):
            last = trainer.train_step(batch)
            if read_back:
                if i + 1 < steps:
                    trainer.prefetch(batch)         # H2D of step i+1 overlaps the compute of step i
                # D2H of this step's loss into pinned memory; it is consumed one step later so the
                # host never stalls the launch of the next step
                loss_host[i % 2].copy_(last["loss"], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if pending is not None:
                    pending[0].synchronize()
                    losses.append(float(loss_host[pending[1]]))
                pending = (ev, i % 2)
        if pending is not None:
            pending[0].synchronize()
            losses.append(float(loss_host[pending[1]]))
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), 0.0)
        if read_back:
            ms = max(ms, wall)                  # host-side stalls count in the end-to-end number
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, last

    # every rank runs the same number of steps (collectives must line up); rank 0 also samples clocks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)          # let nvidia-smi start; the first rows are idle-clock samples
    for _ in range(3):
        trainer.train_step(resident)
    barrier()
    if rank == 0:
        sampler.rows.clear()
    ms_dev, out = timed(resident, False)
    # keep the GPUs under the same load until a few clock samples exist (count derived from the
    # all-reduced time, so it is identical on every rank)
    extra = 0 if ms_dev >= 1500.0 else min(200, int(1500.0 / max(ms_dev / args.steps, 0.1)))
    for _ in range(extra):
        trainer.train_step(resident)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    del losses[:]
    timed(host, True, steps=max(args.warmup, 3))      # warm-up of the end-to-end path itself (staging buffers, copy stream)
    del losses[:]
    ms_e2e, out2 = timed(host, True)
    loss = losses[-1] if losses else float(out2["loss"])
    if not (loss == loss):
        raise RuntimeError("bench: loss is NaN")

    # kernels launched per step (entry-point calls recorded while the step ran / was captured)
    per_step = trainer.launches_per_step
    total = args.batch * world * args.steps
    line = {
        "metric": METRIC, "value": total / ms_dev * 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": dict(workload_config(args, world), cuda_graph=trainer.use_graph,
                       l2="per-step working set (activations, several GB) far exceeds the 126 MB L2; no flush needed"),
        "e2e": {"value": total / ms_e2e * 1e3, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(per_step) * args.steps,
        "loss": loss,
        "model_tflops": 3 * FWD_GFLOP.get(args.workload, 0) * total / ms_dev if args.side == 256 else None,
    }
    if world > 1:
        dist.barrier()
    if rank == 0:
        line["clocks"] = clocks
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        shapes = conv_layer_table(b2, net, args, dev)
        rows = time_conv_kernels(b2, shapes, args.batch, args.dtype, dev)
        bf16_peak, hbm_peak = peaks.get("bf16_tflops", 1590.0), peaks.get("hbm_gbs", 6650.0)

        def roof_of(r):
            """The bound that applies to one launch (min of the two roofs) and the fraction of it that was reached."""
            t_tensor, t_hbm = r["flops"] / (bf16_peak * 1e12), r["bytes"] / (hbm_peak * 1e9)
            if t_tensor >= t_hbm:
                return dict(bound="tensor", achieved=r["tflops"], peak=bf16_peak, unit="TFLOP/s", frac=r["tflops"] / bf16_peak)
            return dict(bound="hbm", achieved=r["gbs"], peak=hbm_peak, unit="GB/s", frac=r["gbs"] / hbm_peak)

        def load_json(name):
            try:
                return json.load(open(os.path.join(ROOT, "profiles", name)))
            except (OSError, ValueError):
                return {}
        pipe = load_json("r02_pconv_tensor_pipe.json")      # ncu sm__pipe_tensor_cycles_active per shape (committed capture)
        partial_rows = [r for r in rows if " partial" in r["shape"]]
        # the dominant kernel: the PartialConv shape (north_star's operator) with the largest share of the step; nets
        # without PartialConv layers fall back to the largest convolution
        top = max(partial_rows, key=lambda r: r["total_ms"]) if partial_rows else rows[0]
        roof = roof_of(top)
        kname = "conv_%s %s" % (top["op"], top["shape"])
        traffic = load_json("traffic.json").get(kname, {}).get("dram_bytes")   # per launch, from `ncu --set full`
        conv_ms = max(sum(r["total_ms"] for r in rows), 1e-9)
        roof.update(traffic=traffic, kernel=kname, launch_ms=top["ms"], algorithmic_bytes=top["bytes"],
                    algorithmic_flops=top["flops"], tensor_cores=top["tc"],
                    tensor_pipe_active_pct=pipe.get(kname),
                    peak_source="MEASURED_PEAKS.json (burst: kernel timed alone)" if peaks else "fallback",
                    timing="CUDA events on the launch stream, median of 5, 256 MB read between launches evicts the L2",
                    share_of_conv_time=top["total_ms"] / conv_ms)
        # every PartialConv shape of the net (the metric's "partial-conv tensor-pipe util %" half)
        roof["pconv"] = [dict(kernel="conv_%s %s" % (r["op"], r["shape"]), count=r["count"], launch_ms=r["ms"],
                              tflops=r["tflops"], gbs=r["gbs"],
                              tensor_pipe_active_pct=pipe.get("conv_%s %s" % (r["op"], r["shape"])), **roof_of(r))
                         for r in sorted(partial_rows, key=lambda r: -r["total_ms"])]
        if partial_rows:
            pf = sum(r["flops"] * r["count"] for r in partial_rows)
            pt = sum(r["total_ms"] for r in partial_rows)
            roof["pconv_family"] = dict(tflops=pf / pt / 1e9, frac_of_tensor_peak=pf / pt / 1e9 / bf16_peak,
                                        launches_ms=pt, share_of_conv_time=pt / conv_ms)
        cf = sum(r["flops"] * r["count"] for r in rows)
        roof["conv_family"] = dict(tflops=cf / conv_ms / 1e9, frac_of_tensor_peak=cf / conv_ms / 1e9 / bf16_peak,
                                   launches_ms=conv_ms, note="FLOP-weighted over every convolution launch of one step "
                                   "(fprop + dgrad + wgrad), each timed alone")
        if line.get("model_tflops"):
            sus = peaks.get("bf16_tflops_sustained", bf16_peak)
            roof["step"] = dict(model_tflops=line["model_tflops"], frac_of_sustained_tensor_peak=line["model_tflops"] / sus,
                                peak=sus, peak_source="MEASURED_PEAKS.json (sustained: inside a long step)")
        line["roofline"] = roof
        if args.layers:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "conv_layers.json"), "w") as f:
                json.dump(rows, f, indent=1)
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample, _, kind = cpu_step_rate(args, args.cpu_steps, 1, args.cpu_batch)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        if world == 1 and not args.no_gpu_baseline:
            del trainer, net
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager_rates(args, dev)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
