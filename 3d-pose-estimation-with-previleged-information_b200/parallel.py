"""Data-parallel plumbing: one process per GPU, NCCL over NVLink/NVSwitch (gloo on CPU for the
host-logic tests).  Replaces the reference's single-process ``nn.DataParallel`` wrapping
(depth_main.py:72): parameters are broadcast once (not every step), the batch is sharded by rank,
and the only per-step exchange is the gradient all-reduce over the flat gradient buffer, cut into
~25 MB buckets issued back to back (NVSwitch: collective cost is launch/latency bound, not link
bound, so a few large buckets are right).  The sum is turned into the mean inside the fused
Adam kernel (inv_scale = 1/world), so no separate divide pass touches the gradients.
"""
import datetime
import os

import torch
import torch.distributed as dist


COMM_SMS = int(os.environ.get("B2POSE_COMM_SMS", "8"))     # SMs left to the overlapped collectives


def init_from_env(backend=None, timeout_s=180):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            # overlapped mode only (B2POSE_DDP_OVERLAP=1): the gradient all-reduce runs beside the shallow half of
            # backward, and every SM a collective holds is one that the persistent convolution grids have to leave
            # free (b2_set_sm_reserve) -- so the collectives are held to a handful of CTAs.  An EXPOSED all-reduce
            # wants every CTA NCCL would take by itself (measured at 2 GPUs: 0.8 ms with 8 CTAs, 0.3 ms unrestricted)
            if COMM_SMS > 0 and os.environ.get("B2POSE_DDP_OVERLAP", "0") != "0":
                os.environ.setdefault("NCCL_MAX_CTAS", str(COMM_SMS))
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local),
                                    timeout=datetime.timedelta(seconds=timeout_s))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    timeout=datetime.timedelta(seconds=timeout_s))
    return rank, local, world


def shard_range(n_items, rank, world):
    """[lo, hi) of the items rank owns when n_items are dealt out as evenly as possible."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bucket_bounds(n_elems, bucket_elems, align=64):
    """Cut [0, n_elems) into contiguous buckets of about bucket_elems (multiples of align)."""
    bucket_elems = max(align, (int(bucket_elems) + align - 1) // align * align)
    out, lo = [], 0
    while lo < n_elems:
        hi = min(n_elems, lo + bucket_elems)
        out.append((lo, hi))
        lo = hi
    return out


def broadcast_flat(flat, group, src=0):
    dist.broadcast(flat.w, src, group=group)
    if getattr(flat, "w16", None) is not None:
        dist.broadcast(flat.w16, src, group=group)


def merge_ranges(ranges):
    """Sorted union of half-open ranges; touching ranges are merged."""
    out = []
    for lo, hi in sorted(ranges):
        if out and lo <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], hi))
        else:
            out.append((lo, hi))
    return out


def split_ranges(n_elems, deep):
    """(deep, shallow): the merged `deep` ranges and their complement in [0, n_elems)."""
    deep = merge_ranges(deep)
    shallow, pos = [], 0
    for lo, hi in deep:
        if lo > pos:
            shallow.append((pos, lo))
        pos = hi
    if pos < n_elems:
        shallow.append((pos, n_elems))
    return deep, shallow


class GradBuckets:
    """Bucketed SUM all-reduce of a flat gradient buffer (reverse order = the order backward fills it).

    ``deep`` (optional list of element ranges) names the part of the buffer that is complete after the FIRST stage of
    a two-stage backward pass: `start_deep()` issues its buckets on the communication stream while the second stage
    still computes, `finish()` issues the rest and makes the current stream wait for everything."""

    def __init__(self, flat, group, bucket_mb=25.0, deep=None, compress=False):
        self.flat, self.group = flat, group
        # compress: exchange the gradients as bf16 (half the bytes on the wire); the casts either side are two
        # streaming passes over the flat buffer (~30 us each for 30 M parameters)
        self.g16 = torch.empty(flat.n, dtype=torch.bfloat16, device=flat.g.device) if (compress and flat.g.is_cuda) else None
        per = int(bucket_mb * (1 << 20)) // (2 if self.g16 is not None else flat.g.element_size())
        self.bounds = list(reversed(bucket_bounds(flat.n, per)))
        self.deep_bounds, self.shallow_bounds, self._works = [], self.bounds, []
        if deep:
            d, sh = split_ranges(flat.n, deep)
            cut = lambda rs: [(lo + a, lo + b) for lo, hi in reversed(rs) for a, b in reversed(bucket_bounds(hi - lo, per))]
            self.deep_bounds, self.shallow_bounds = cut(d), cut(sh)
        self.comm_stream = torch.cuda.Stream(device=flat.g.device) if flat.g.is_cuda else None

    def _issue(self, bounds):
        buf = self.flat.g if self.g16 is None else self.g16
        return [dist.all_reduce(buf[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                for lo, hi in bounds]

    def _pack(self, bounds):
        if self.g16 is not None:
            from . import _lib as L
            for lo, hi in merge_ranges(bounds):
                L.call("b2_cast_f32_to_bf16", self.flat.g[lo:hi].data_ptr(), self.g16[lo:hi].data_ptr(), hi - lo, L.stream())

    def _unpack(self, bounds):
        if self.g16 is not None:
            from . import _lib as L
            for lo, hi in merge_ranges(bounds):
                L.call("b2_cast_bf16_to_f32", self.g16[lo:hi].data_ptr(), self.flat.g[lo:hi].data_ptr(), hi - lo, L.stream())

    def allreduce(self):
        self._pack(self.bounds)
        for w in self._issue(self.bounds):
            w.wait()
        self._unpack(self.bounds)

    def start_deep(self):
        """After the first backward stage: all-reduce the deep ranges beside whatever the current stream runs next."""
        self._pack(self.deep_bounds)
        if self.comm_stream is None:
            self._works = self._issue(self.deep_bounds)
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self._works = self._issue(self.deep_bounds)

    def finish(self):
        """After the second stage: all-reduce the rest, then wait for every outstanding bucket."""
        self._pack(self.shallow_bounds)
        works = self._works + self._issue(self.shallow_bounds)
        self._works = []
        for w in works:
            w.wait()
        self._unpack(self.bounds)
