"""Training-step loop of the depth stream: drop-in for ``depth_train.Trainer`` (``vanilla_train``
depth_train.py:376-462, ``fusion_train`` :286-373, ``adapt_learn_rate`` :621-638, ``*_infer``
:650-679) and ``train.Trainer.cam_train`` (train.py:145-192).

B200-first differences from the reference loop (results identical within tolerance):
  * all parameters live in ONE flat fp32 buffer (gradients, Adam moments and the bf16 shadow
    filters likewise), so clip-norm + Adam is two kernels instead of ~480 tensor ops and the
    data-parallel exchange is a handful of bucketed NCCL all-reduces over contiguous memory;
  * ``half_acc`` selects bf16 tensor-core compute with fp32 masters (no loss scaling needed); the
    reference's inf-skip survives as a finite check inside the fused Adam kernel, without the
    per-parameter host syncs of depth_train.py:435;
  * the whole step (zero-grad, forward, head, loss, backward, clip, Adam) is captured in a CUDA
    graph and replayed; the learning rate and Adam bias corrections are read from a small device
    buffer so the schedule moves without re-capture;
  * one process per GPU; every rank takes the masked mean over ITS valid joints and the gradients are averaged over
    ranks.  DataParallel's loss is the mean over the valid joints of the gathered batch, so the two agree exactly when
    every rank holds the same number of valid joints and differ by the ratio cnt_rank / mean(cnt) otherwise (documented
    deviation: no extra collective for the count); BN statistics stay per rank like DataParallel's per-replica BN;
  * the Adam bias-correction step advances on steps the fused kernel skips for non-finite gradients, whereas the
    reference does not call optimizer.step() there (:435-438); with bf16 + fp32 masters this only matters after a
    genuine overflow.
"""
import math
from types import SimpleNamespace

import torch
from torch import nn

from . import _lib as L
from . import ops, utils
from .layers import BatchNorm2d, _B2ConvBase

_ALIGN = 64      # elements; keeps every parameter slice 256-byte (fp32) / 128-byte (bf16) aligned


def train_args(**kw):
    """Namespace with the reference's option names and defaults (opts.py:1-78)."""
    base = dict(model="resnet50", half_acc=False, depth_only=True, do_fusion=False, do_teach=False,
                partial_conv=False, pretrain=False, early_dist=False, skip_relu=False, extra_channel=False,
                joint_space=False, warmup=1, n_epochs=20, batch_size=64, side_in=257, stride=16, num_joints=19,
                depth=16, warmup_factor=0.2, learn_rate=5e-5, learn_decay=0.2, grad_norm=5.0, grad_scaling=32.0,
                weight_decay=4e-5, depth_range=1000.0, loss_div=10.0, criterion="SmoothL1", semi_teach=False,
                sigmoid=False, bin_dist=False, do_freeze=False, alpha_init=0.1, alpha_dest=0.1, alpha_span=10)
    base.update(kw)
    return SimpleNamespace(**base)


def synthetic_batch(n, side, num_joints, device, seed=1, invalid_frac=0.25, key_index=None, pin=False):
    """Synthetic (color, depth, true_cam, true_val) of the dataset tuple layout
    (depth_datasets.py:199-237): color ~ N(0,1), depth = U(0.05,1) with rectangular holes of
    invalid (zero) pixels, joints ~ N(0, 300 mm), ~90 % valid with the root always valid."""
    g = torch.Generator().manual_seed(seed)
    color = torch.randn(n, 3, side, side, generator=g)
    depth = torch.rand(n, 1, side, side, generator=g) * 0.95 + 0.05
    lo, hi = max(2, side // 16), max(3, (side * 3) // 8)
    for i in range(n):
        holes = torch.ones(side, side)
        guard = 0
        while float(1 - holes.mean()) < invalid_frac and guard < 1000:
            h, w = (int(v) for v in torch.randint(lo, hi + 1, (2,), generator=g))
            top = int(torch.randint(0, max(1, side - h + 1), (1,), generator=g))
            left = int(torch.randint(0, max(1, side - w + 1), (1,), generator=g))
            holes[top:top + h, left:left + w] = 0
            guard += 1
        depth[i, 0] *= holes
    true_cam = torch.randn(n, num_joints, 3, generator=g) * 300.0
    true_val = torch.rand(n, num_joints, generator=g) < 0.9
    true_val[:, (num_joints - 1) if key_index is None else key_index] = True
    out = (color, depth, true_cam, true_val)
    if pin:
        return tuple(t.pin_memory() for t in out)
    return tuple(t.to(device) for t in out) if device is not None else out


class _FlatState:
    """All trainable parameters (and grads / Adam moments / bf16 shadows) as views of flat buffers."""

    def __init__(self, model, want_shadow):
        params = [p for p in model.parameters() if p.requires_grad]
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.n = total
        self.w = torch.zeros(total, dtype=torch.float32, device=dev)
        self.g = torch.zeros_like(self.w)
        self.m = torch.zeros_like(self.w)
        self.v = torch.zeros_like(self.w)
        self.w16 = torch.zeros(total, dtype=torch.bfloat16, device=dev) if want_shadow else None
        self.params, self.offsets = params, offs
        owner, bias_owner = {}, {}
        for mod in model.modules():
            if isinstance(mod, _B2ConvBase):
                owner[id(mod.weight)] = mod
                if mod.bias is not None:
                    bias_owner[id(mod.bias)] = mod
        for p, o in zip(params, offs):
            n = p.numel()
            if p.dim() == 4:                       # filters: keep KRSC memory order
                K, Cc, R, S = p.shape
                src = p.data.permute(0, 2, 3, 1).contiguous().view(-1).float()
                self.w[o:o + n].copy_(src)
                p.data = self.w[o:o + n].view(K, R, S, Cc).permute(0, 3, 1, 2)
                p.grad = self.g[o:o + n].view(K, R, S, Cc).permute(0, 3, 1, 2)
                if id(p) in owner:
                    owner[id(p)]._grad_sink = p.grad
                    if want_shadow:
                        owner[id(p)]._shadow = self.w16[o:o + n].view(K, R, S, Cc)
            else:
                self.w[o:o + n].copy_(p.data.view(-1).float())
                p.data = self.w[o:o + n].view(p.shape)
                p.grad = self.g[o:o + n].view(p.shape)
                if id(p) in bias_owner:
                    bias_owner[id(p)]._bias_sink = p.grad
        if want_shadow:
            self.refresh_shadow()

    def refresh_shadow(self):
        if self.w16 is not None:
            L.call("b2_cast_f32_to_bf16", L.ptr(self.w), L.ptr(self.w16), self.n, L.stream())


class Trainer:
    def __init__(self, args, model, data_info, use_graph=True, process_group=None, bucket_mb=128.0, overlap=None):
        self.model = model
        self.data_info = data_info
        self.key_index = data_info["key_index"] if isinstance(data_info, dict) else data_info.key_index
        g = lambda name, default: getattr(args, name, default)
        self.half_acc = bool(g("half_acc", False))
        self.depth_only = bool(g("depth_only", True))
        self.do_fusion = bool(g("do_fusion", False))
        if g("semi_teach", False):
            raise NotImplementedError("semi_teach needs the reference's private data loaders; pass the unlabelled "
                                      "batch to train_step(batch, semi_batch=...) instead")
        # distillation ("privileged information") step: depth_train.py:52-56,97-99
        self.do_teach = bool(g("do_teach", False))
        self.sigmoid, self.bin_dist = bool(g("sigmoid", False)), bool(g("bin_dist", False))
        self.do_freeze = bool(g("do_freeze", False))
        self.alpha_init, self.alpha_dest = g("alpha_init", 0.1), g("alpha_dest", 0.1)
        self.alpha_span = g("alpha_span", 10)
        self.alpha = float(self.alpha_init)
        self.teacher = None
        self.depth, self.num_joints = args.depth, args.num_joints
        self.side_in, self.stride = args.side_in, args.stride
        self.depth_range = g("depth_range", 1000.0)
        self.warmup, self.learn_rate = g("warmup", 1), g("learn_rate", 5e-5)
        self.learn_decay, self.num_epochs = g("learn_decay", 0.2), g("n_epochs", 20)
        self.warmup_factor = g("warmup_factor", 0.2)
        self.grad_norm, self.loss_div = g("grad_norm", 5.0), g("loss_div", 10.0)
        self.weight_decay = g("weight_decay", 4e-5)
        self.criterion = g("criterion", "SmoothL1")
        self.thresh = g("thresh", None)                 # dict(solid=, close=, rough=) mm: depth_train.py:62 / train.py:47-51
        if self.thresh is None and g("thresh_rough", None) is not None:
            self.thresh = dict(solid=g("thresh_solid", None), close=g("thresh_close", None), rough=g("thresh_rough", None))
        if self.criterion not in ops.CRITERIA:
            raise ValueError("criterion must be one of %s" % sorted(ops.CRITERIA))
        self.legacy = getattr(model, "kind", "") == "resnet"           # train.py:174 has no loss_div
        if self.half_acc:
            self.model = self.model.half()      # bf16 compute, fp32 masters stay in the flat buffer
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("Trainer needs the model on a CUDA device; there is no CPU fallback")
        self.device = dev
        self.list_names = [n for n, _ in model.named_parameters()]
        self.flat = _FlatState(model, want_shadow=self.half_acc)
        self.list_params = self.flat.params
        # weights loaded from outside (model.load_state_dict) land in the fp32 flat buffer through the re-pointed
        # parameters; the bf16 shadow filters the kernels read must follow
        model.register_load_state_dict_post_hook(lambda module, incompatible: self.flat.refresh_shadow())
        self.bns = [m for m in model.modules() if isinstance(m, BatchNorm2d)]
        for m in self.bns:
            m.defer_count = True
            m._grad_sinks = (m.weight.grad, m.bias.grad)
        self.lr = self.learn_rate
        self.step_count = 0
        self.betas, self.eps = (0.9, 0.999), 1e-8
        self.hyper = torch.zeros(4, dtype=torch.float32, device=dev)
        self._hyper_ring = [(torch.zeros(4, dtype=torch.float32).pin_memory(), None) for _ in range(8)]
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.use_graph = use_graph
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._stage = {}
        self._prefetched = None
        self._stage_free = None
        self.launches_per_step = 0
        self._graphs = {}
        self._static = {}
        self._static_semi = {}
        self._extras = {}
        self.pg = process_group
        self.world = 1
        self.overlap = False
        self.buckets = None
        self._force_two = __import__("os").environ.get("B2POSE_FORCE_TWO_STAGE", "0") != "0"
        self.comm_sms = 0
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            import torch.distributed as dist
            self.pg = process_group if process_group is not None else dist.group.WORLD
            self.world = dist.get_world_size(self.pg)
            from .parallel import broadcast_flat, GradBuckets, COMM_SMS
            broadcast_flat(self.flat, self.pg)
            # two-stage backward: layer3 / layer4 / regressor gradients (~90 % of the bytes) are complete when the
            # backward pass reaches the input of layer3 and are all-reduced while the shallow half still computes
            # (measured at 2 GPUs, profiles/r02_ddp_overlap.md: the overlapped exchange LOSES to one exposed all-reduce
            #  -- 15.83-15.99 ms against 15.61 ms per step -- because the NCCL kernels take SMs away from the persistent
            #  one-CTA-per-SM convolution grids; it stays available behind B2POSE_DDP_OVERLAP=1)
            env = __import__("os").environ
            self.overlap = overlap if overlap is not None else env.get("B2POSE_DDP_OVERLAP", "0") != "0"
            # bf16 compute: exchange the gradients as bf16 too (they come out of bf16 activations); fp32 mode keeps fp32
            compress = self.half_acc and env.get("B2POSE_DDP_BF16", "1") != "0"
            deep = [(o, o + (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN)
                    for name, p, o in zip(self.list_names, self.flat.params, self.flat.offsets)
                    if name.startswith(("layer3.", "layer4.", "regressor.", "cam_regressor.", "mat_regressor."))]
            bucket_mb = float(env.get("B2POSE_BUCKET_MB", bucket_mb))
            self.buckets = GradBuckets(self.flat, self.pg, bucket_mb, deep=deep if self.overlap else None,
                                       compress=compress)
            self.comm_sms = COMM_SMS
            for m in model.buffers():
                dist.broadcast(m, 0, group=self.pg)

    # ------------------------------------------------------------------ schedule
    def adapt_learn_rate(self, epoch):
        e = epoch - 1
        if e < self.warmup:
            lr = self.learn_rate * self.warmup_factor
        elif e < 15:
            lr = self.learn_rate
        elif e < 20:
            lr = self.learn_rate * self.learn_decay
        elif e < 25:
            lr = self.learn_rate * self.learn_decay ** 2
        else:
            lr = self.learn_rate * self.learn_decay ** 3
        self.lr = lr
        return lr

    def to(self, image, device):
        return image.to(device, non_blocking=True)      # the bf16 cast happens on device (ops.to_nhwc)

    # ------------------------------------------------------------------ distillation (depth_train.py:107-129,156-283)
    def set_teacher(self, teacher):
        """depth_train.py:107-108.  The teacher only ever runs under no_grad (:194-195)."""
        if next(teacher.parameters()).device != self.device:
            raise RuntimeError("the teacher must live on the trainer's device (%s)" % self.device)
        self.teacher = teacher.half() if self.half_acc else teacher
        for p in self.teacher.parameters():
            p.requires_grad_(False)
        self._graphs.clear()                              # captured steps do not contain the teacher

    def get_dist_weight(self, epoch):
        """depth_train.py:641-647."""
        import numpy as np
        alphas = np.linspace(self.alpha_init, self.alpha_dest, self.alpha_span)
        return float(alphas[epoch - 1]) if epoch - 1 < self.alpha_span else float(self.alpha_dest)

    def freeze_batchnorm(self):
        """depth_train.py:156-158."""
        self.teacher.eval()
        self.model.freeze_batchnorm()

    def teach_infer(self, color_image, depth_image):
        """depth_train.py:682-691."""
        if getattr(self.teacher, "fused", False) or self.do_fusion:
            return self.teacher(color_image, depth_image)
        return self.teacher(depth_image if self.depth_only else color_image)

    def distill(self, batch, teach_last, last_feat, atten_map):
        """depth_train.py:115-129 as one fused reduction (+ one backward pass)."""
        return utils.mimic_loss(teach_last, last_feat, atten_map, sigmoid=self.sigmoid, bin_dist=self.bin_dist)

    def _distill_forward(self, color, depth, atten):
        with torch.no_grad():
            _, teach_last = self.teach_infer(color, depth)
        cam_feat, last_feat = self.vanilla_infer(color, 0, True)       # the student always sees the colour image (:197)
        return cam_feat, self.distill(color.size(0), teach_last, last_feat, atten)

    # ------------------------------------------------------------------ forward pieces
    def vanilla_infer(self, in_image, i_batch=0, ret_last=False):
        out = self.model(in_image)
        cam_feat, last_feat = (out, None) if self.legacy else out
        if self.legacy and isinstance(cam_feat, tuple):
            cam_feat = cam_feat[0]
        return (cam_feat, last_feat) if ret_last else cam_feat

    def fusion_infer(self, color_image, depth_image, i_batch=0, ret_last=False):
        cam_feat, last_feat = self.model(color_image, depth_image)
        return (cam_feat, last_feat) if ret_last else cam_feat

    def _forward_loss(self, color, depth, true_cam, true_val, atten=None, semi=None):
        self._extras = {}
        if atten is not None:
            if self.teacher is None:
                raise RuntimeError("a 5-tuple batch (with an attention map) needs set_teacher() first")
            cam_feat, dist_loss = self._distill_forward(color, depth, atten)
            coords = utils.heatmap_coords(cam_feat, self.depth, self.num_joints, self.depth_range)
            cam_loss, spec = utils.pose_loss(coords, true_cam, true_val, self.key_index, self.loss_div, self.criterion)
            alpha = self.hyper[3]                      # device scalar: the schedule moves without re-capture
            loss = dist_loss * alpha + cam_loss        # depth_train.py:219
            self._extras = dict(cam_loss=cam_loss.detach(), dist_loss=dist_loss.detach())
            if semi is not None:                       # semi_train, depth_train.py:132-153,221-229
                _, semi_loss = self._distill_forward(semi[0], semi[1], semi[-1])
                loss = loss + semi_loss * alpha
                self._extras["semi_dist_loss"] = semi_loss.detach()
            return loss, spec
        if self.do_fusion or getattr(self.model, "fused", False):
            cam_feat = self.fusion_infer(color, depth)
        else:
            kind = getattr(self.model, "kind", "depthnet")
            use_depth = kind == "partial_depthnet" or (kind == "depthnet" and self.depth_only)
            cam_feat = self.vanilla_infer(depth if use_depth else color)
        coords = utils.heatmap_coords(cam_feat, self.depth, self.num_joints, self.depth_range)
        loss, spec = utils.pose_loss(coords, true_cam, true_val, self.key_index,
                                     1.0 if self.legacy else self.loss_div, self.criterion)
        return loss, spec

    # ------------------------------------------------------------------ one optimisation step
    def _fwd_bwd(self, batch, semi=None):
        self.flat.g.zero_()
        ops.bn_arena_begin(self.device)          # zeroed per-channel accumulators of the totals BatchNorm path
        ops.wgrad_overlap_begin(self.device)     # weight gradients run on a side stream until the join below
        try:
            if semi is not None:
                loss, spec = self._forward_loss(*batch, semi=semi)
            else:
                loss, spec = self._forward_loss(*batch)
            ops.wgrad_overlap_sync(self.device)   # filter transforms prepared on the side stream during forward
            loss.backward()
        finally:
            ops.wgrad_overlap_end(self.device)
            ops.bn_arena_end(self.device)
        return loss.detach(), spec

    # ---- data parallel: backward in two stages around the input of layer3 (see nets.ResNet.forward) ----
    def _two_stage(self, semi):
        if semi is not None or self.teacher is not None:
            return False
        # (B2POSE_FORCE_TWO_STAGE=1: run the staged backward in a single process too -- tools/two_stage_check.py)
        return (self.overlap and self.world > 1) or self._force_two

    def _fwd_bwd_deep(self, batch):
        """Stage 1: zero-grad, forward, head, loss, backward down to the input of layer3.  On return every gradient of
        layer3 / layer4 / the regressor is in the flat buffer and the weight-gradient stream has been joined."""
        self.flat.g.zero_()
        ops.bn_arena_begin(self.device)
        ops.wgrad_overlap_begin(self.device)
        self.model._mark_boundary = True
        try:
            loss, spec = self._forward_loss(*batch)
            f, cut = self.model._boundary
            ops.wgrad_overlap_sync(self.device)
            loss.backward()                      # stops at the detached leaf `cut` (nets.ResNet.forward)
            ops.wgrad_overlap_join(self.device)
        except BaseException:
            ops.wgrad_overlap_end(self.device)
            ops.bn_arena_end(self.device)
            raise
        finally:
            self.model._mark_boundary = False
            self.model._boundary = None
        self._stage2 = (f, cut.grad)
        return loss.detach(), spec

    def _bwd_shallow(self):
        """Stage 2: the rest of the backward pass (fusion, layer1/2/5/6, stems)."""
        f, df = self._stage2
        self._stage2 = None
        try:
            torch.autograd.backward([f], [df])
        finally:
            ops.wgrad_overlap_end(self.device)
            ops.bn_arena_end(self.device)

    def _update(self):
        f = self.flat
        self.sumsq.zero_()
        L.call("b2_grad_sumsq", L.ptr(f.g), f.n, L.ptr(self.sumsq), L.stream())
        L.call("b2_adam_step", L.ptr(f.w), L.ptr(f.g), L.ptr(f.m), L.ptr(f.v), L.ptr(f.w16), f.n,
               float(self.lr), self.betas[0], self.betas[1], self.eps, float(self.weight_decay), 1,
               L.ptr(self.sumsq), float(self.grad_norm), 1.0 / self.world, L.ptr(self.hyper), L.stream())

    def _set_hyper(self):
        self.step_count += 1
        t = self.step_count
        slot = t % len(self._hyper_ring)
        host, ev = self._hyper_ring[slot]
        if ev is not None:
            ev.synchronize()                         # the async copy that last read this pinned slot is done
        host[0] = self.lr
        host[1] = 1.0 - self.betas[0] ** t
        host[2] = math.sqrt(1.0 - self.betas[1] ** t)
        host[3] = self.alpha
        self.hyper.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._hyper_ring[slot] = (host, ev)

    def _mode_key(self):
        """Everything a captured step bakes in besides the batch shapes: replaying a graph captured under other
        BatchNorm modes or hyper-parameters would silently run the stale configuration."""
        return (tuple(m.training for m in self.bns), self.model.training, self.criterion, float(self.weight_decay),
                float(self.grad_norm), self.betas, float(self.eps), float(self.loss_div), float(self.depth_range),
                self.teacher is not None, self.sigmoid, self.bin_dist, self.world)

    def _buffers(self, table, batch):
        key = tuple(tuple(t.shape) for t in batch)
        st = table.get(key)
        if st is None:
            st = tuple(torch.empty(t.shape, dtype=(torch.uint8 if t.dtype == torch.bool else t.dtype),
                                   device=self.device) for t in batch)
            table[key] = st
        return key, st

    def prefetch(self, batch):
        """Start the host->device copy of the NEXT batch on a side stream so that it overlaps the
        current step (the DataLoader's pinned batches; depth_train.py:386-389 does this copy inline).
        `train_step(batch)` with the same tensors then only does a device-to-device hand-over."""
        fresh = tuple(tuple(t.shape) for t in batch) not in self._stage
        _, stage = self._buffers(self._stage, batch)
        cs = self._copy_stream
        if fresh:
            # the staging buffers may recycle memory whose last users are still queued on the compute
            # stream; order the first side-stream write after them
            cs.wait_stream(torch.cuda.current_stream())
        if self._stage_free is not None:
            cs.wait_event(self._stage_free)          # the previous hand-over has read the staging buffers
        with torch.cuda.stream(cs):
            for s, t in zip(stage, batch):
                s.copy_(t.view(torch.uint8) if t.dtype == torch.bool else t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._prefetched = (tuple(id(t) for t in batch), ev, stage)

    def _static_batch(self, batch):
        key, st = self._buffers(self._static, batch)
        pf = self._prefetched
        if pf is not None and pf[0] == tuple(id(t) for t in batch):
            cur = torch.cuda.current_stream()
            cur.wait_event(pf[1])
            for s, g in zip(st, pf[2]):
                s.copy_(g, non_blocking=True)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(cur)
            self._prefetched = None
            return key, st
        for s, t in zip(st, batch):
            s.copy_(t.view(torch.uint8) if t.dtype == torch.bool else t, non_blocking=True)
        return key, st

    def train_step(self, batch, semi_batch=None):
        """One fwd + bwd + clip + Adam step on ``batch`` = (color, depth, true_cam, true_val)
        (host or device tensors).  Returns dict(loss=0-dim tensor, spec_cam=[N,J,3], grad_sumsq).

        With a teacher set (``set_teacher``) and a 5-tuple batch (..., atten_map) this is the distillation
        step of depth_train.py:179-283: loss = dist_loss * alpha + cam_loss; the dict then also carries
        ``cam_loss`` and ``dist_loss`` (``loss`` is the total).  ``semi_batch`` = an unlabelled 5-tuple whose
        mimic term is added (semi_train, :132-153)."""
        self._set_hyper()
        key, st = self._static_batch(batch)
        st_semi = None
        if semi_batch is not None:
            if len(batch) != 5 or len(semi_batch) != 5:
                raise ValueError("semi_batch needs the distillation 5-tuples (color, depth, true_cam, true_val, atten_map)")
            self._static, keep = self._static_semi, self._static          # separate static buffers
            try:
                key2, st_semi = self._static_batch(semi_batch)
            finally:
                self._static_semi, self._static = self._static, keep
            key = (key, key2)
        dist_on = self.world > 1
        n0 = L.launches
        two = self._two_stage(st_semi)
        if not self.use_graph:
            loss, spec = self._eager_step(st, st_semi, two)
            extras = self._extras
            self.launches_per_step = L.launches - n0
        else:
            key = (key, self._mode_key())
            entry = self._graphs.get(key)
            if entry is None:
                entry = self._capture(key, st)
            if entry["stage"] < 3:               # eager warm-up iterations before capture
                loss, spec = self._eager_step(st, st_semi, two)
                extras = self._extras
                entry["stage"] += 1
                if entry["stage"] == 3:
                    self._do_capture(entry, st, st_semi, two)
            else:
                entry["fb"].replay()
                if entry.get("fb2") is not None:     # two-stage backward: deep buckets travel beside the shallow stage
                    if self.buckets is not None:
                        self.buckets.start_deep()
                    entry["fb2"].replay()
                    if self.buckets is not None:
                        self.buckets.finish()
                elif dist_on:
                    self.buckets.allreduce()
                entry["up"].replay()
                loss, spec, extras = entry["loss"], entry["spec"], entry["extras"]
        if self.bns and self.model.training:
            live = [m.num_batches_tracked for m in self.bns if m.training]      # none after freeze_batchnorm()
            if live:
                torch._foreach_add_(live, 1)
        return dict(loss=loss, spec_cam=spec, grad_sumsq=self.sumsq, **extras)

    def _eager_step(self, st, st_semi, two):
        if two:
            loss, spec = self._fwd_bwd_deep(st)
            if self.buckets is not None:
                self.buckets.start_deep()
            self._bwd_shallow()
            if self.buckets is not None:
                self.buckets.finish()
        else:
            loss, spec = self._fwd_bwd(st, st_semi)
            if self.world > 1:
                self.buckets.allreduce()
        self._update()
        return loss, spec

    def _capture(self, key, st):
        entry = dict(stage=0)
        self._graphs[key] = entry
        return entry

    def _do_capture(self, entry, st, st_semi=None, two=False):
        torch.cuda.synchronize()
        pool = torch.cuda.graph_pool_handle()
        n0 = L.launches
        fb, fb2 = torch.cuda.CUDAGraph(), None
        if two:
            with torch.cuda.graph(fb, pool=pool):
                loss, spec = self._fwd_bwd_deep(st)
            fb2 = torch.cuda.CUDAGraph()
            # the grids of the shallow stage leave the communication SMs free: its kernels run beside the NCCL
            # kernels of the deep buckets
            L.call("b2_set_sm_reserve", int(self.comm_sms))
            try:
                with torch.cuda.graph(fb2, pool=pool):
                    self._bwd_shallow()
            finally:
                L.call("b2_set_sm_reserve", 0)
        else:
            with torch.cuda.graph(fb, pool=pool):
                loss, spec = self._fwd_bwd(st, st_semi)
        up = torch.cuda.CUDAGraph()
        with torch.cuda.graph(up, pool=pool):
            self._update()
        entry.update(fb=fb, fb2=fb2, up=up, loss=loss, spec=spec, extras=self._extras)
        self.launches_per_step = L.launches - n0     # libb2pose kernels recorded into the two graphs
        # the capture itself did not execute: the grads in the flat buffer are from the last eager
        # warm-up step and have been consumed already, nothing to redo.

    # ------------------------------------------------------------------ epoch loops (reference API)
    def _epoch(self, epoch, data_loader, device):
        n_batches = len(data_loader)
        loss_avg, total = 0.0, 0
        it = iter(data_loader)
        nxt = next(it, None)
        i_batch = -1
        while nxt is not None:
            batch, i_batch = tuple(nxt), i_batch + 1
            out = self.train_step(batch)
            nxt = next(it, None)
            if nxt is not None and not nxt[0].is_cuda:
                nxt = tuple(nxt)
                self.prefetch(nxt)                  # overlaps this step's compute
            n = batch[2].size(0)
            val = out["loss"].item()
            print("| train Epoch[%d] [%d/%d]  Loss %1.4f" % (epoch, i_batch, n_batches, val), flush=True)
            loss_avg += val * n
            total += n
        loss_avg /= max(total, 1)
        print("\n=> train Epoch[%d]  Cam Loss: %1.4f\n" % (epoch, loss_avg))
        return dict(cam_train_loss=loss_avg)

    def distill_train(self, epoch, data_loader, device=None, semi_loader=None):
        """depth_train.py:161-283: data_loader yields (color, depth, true_cam, true_val, atten_map)."""
        if self.teacher is None:
            raise RuntimeError("distill_train needs set_teacher() first")
        if self.do_freeze:
            self.freeze_batchnorm()
        self.alpha = self.get_dist_weight(epoch)
        print("\n=> alpha value: {:.2f}".format(self.alpha))
        n_batches = len(data_loader)
        cam_sum = dist_sum = 0.0
        cam_n = dist_n = 0
        semi_it = iter(semi_loader) if semi_loader is not None else None
        for i_batch, batch in enumerate(data_loader):
            semi = None
            if semi_it is not None:
                semi = next(semi_it, None)
                if semi is None:                         # depth_train.py:133-138: restart the semi worker
                    semi_it = iter(semi_loader)
                    semi = next(semi_it)
            out = self.train_step(tuple(batch), None if semi is None else tuple(semi))
            n = batch[2].size(0)
            cam, dist = out["cam_loss"].item(), out["dist_loss"].item()
            cam_sum, cam_n = cam_sum + cam * n, cam_n + n
            dist_sum, dist_n = dist_sum + dist * n, dist_n + n
            msg = "[=] train Epoch[{0}] Batch[{1}|{2}] ".format(epoch, i_batch, n_batches)
            msg += " Cam Loss {:.4f} ".format(cam) + " Dist Loss {:.4f} ".format(dist)
            if semi is not None:
                sn = semi[2].size(0)
                sd = out["semi_dist_loss"].item()
                dist_sum, dist_n = dist_sum + sd * sn, dist_n + sn
                msg += " Semi Loss {:.4f}".format(sd)
            print(msg, flush=True)
        cam_sum /= max(cam_n, 1)
        dist_sum /= max(dist_n, 1)
        print("\n=> train Epoch[%d]  Cam Loss: %1.4f  Dist Loss: %1.4f\n\n" % (epoch, cam_sum, dist_sum))
        return dict(dist_train_loss=dist_sum, cam_train_loss=cam_sum)

    def vanilla_train(self, epoch, data_loader, device=None):
        return self._epoch(epoch, data_loader, device)

    def fusion_train(self, epoch, data_loader, device=None):
        return self._epoch(epoch, data_loader, device)

    cam_train = vanilla_train        # train.py:145-192 (legacy RGB loop; same head + loss)

    def train(self, epoch, data_loader):
        self.model.train()
        self.adapt_learn_rate(epoch)
        if self.do_teach:
            return self.distill_train(epoch, data_loader, self.device)
        if self.do_fusion:
            return self.fusion_train(epoch, data_loader, self.device)
        return self.vanilla_train(epoch, data_loader, self.device)

    # ------------------------------------------------------------------ evaluation loops (depth_train.py:477-618)
    @torch.no_grad()
    def _test_epoch(self, epoch, test_loader, device=None):
        """fusion_test / vanilla_test: eval forward + head + loss per batch, back-rotation and the
        analyze / parse_epoch metrics accumulated on the device (one read-back per epoch).
        test_loader yields (color, depth, true_cam, true_val, back_rotate)."""
        thresh = getattr(self, "thresh", None)
        if thresh is None:
            raise RuntimeError("set trainer.thresh = dict(solid=, close=, rough=) (metadata['thresholds'], "
                               "depth_train.py:62) before testing")
        mirror = self.data_info.get("mirror") if isinstance(self.data_info, dict) else getattr(self.data_info, "mirror", None)
        acc = utils.MetricAccumulator(mirror, thresh, self.device)
        n_batches = len(test_loader)
        loss_avg, total = 0.0, 0
        for i_batch, (color, depth, true_cam, true_val, back_rotate) in enumerate(test_loader):
            color, depth = color.to(self.device), depth.to(self.device)
            true_cam, true_val = true_cam.to(self.device), true_val.to(self.device)
            loss, spec = self._forward_loss(color, depth, true_cam,
                                            true_val.view(torch.uint8) if true_val.dtype == torch.bool else true_val)
            acc.update(spec, true_cam, true_val, torch.as_tensor(back_rotate).to(self.device))
            n = true_cam.size(0)
            val = loss.item()
            loss_avg += val * n
            total += n
            print("| test Epoch[%d] [%d/%d]  Cam Loss %1.4f" % (epoch, i_batch, n_batches, val), flush=True)
        loss_avg /= max(total, 1)
        record = dict(test_loss=loss_avg)
        res = acc.result()
        res.pop("batch_size")
        record.update(res)
        print("\n=> test Epoch[%d]  Cam Loss: %1.4f\n" % (epoch, loss_avg))
        print("=>[SPEC] cam_mean: %1.3f  [pck]: %1.3f  [auc]: %1.3f\n"
              % (record["cam_mean"], record["score_pck"], record["score_auc"]))
        return record

    def fusion_test(self, epoch, test_loader, device=None):
        return self._test_epoch(epoch, test_loader, device)

    def vanilla_test(self, epoch, test_loader, device=None):
        return self._test_epoch(epoch, test_loader, device)

    cam_test = vanilla_test

    def test(self, epoch, test_loader):
        """depth_train.py:610-618."""
        self.model.eval()
        return self._test_epoch(epoch, test_loader, self.device)

    @torch.no_grad()
    def predict(self, batch):
        """Eval-mode forward + head: returns (spec_cam [N,J,3], loss) like the *_test loops' core."""
        was = self.model.training
        self.model.eval()
        try:
            batch = tuple(t.to(self.device) for t in batch)
            loss, spec = self._forward_loss(batch[0], batch[1], batch[2],
                                            batch[3].view(torch.uint8) if batch[3].dtype == torch.bool else batch[3])
        finally:
            self.model.train(was)
        return spec, loss
