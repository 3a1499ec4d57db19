"""torch.autograd bindings of the libb2pose kernels.

Internal activations are NHWC-contiguous tensors ``[N, H, W, C]`` (fp32 or bf16), veils / masks
are ``[N, H, W]`` fp32 in {0, 1}, filters are the reference's ``[K, C, R, S]`` parameters held in
channels_last memory (= KRSC).  Every function here launches CUDA kernels through the C ABI; none
has a PyTorch or CPU fallback.
"""
import ctypes as C

import torch
from torch.autograd import Function

from . import _lib as L


# --------------------------------------------------------------------------- helpers
def conv_out_size(h, w, r, s, stride, pad, dil):
    return ((h + 2 * pad - dil * (r - 1) - 1) // stride + 1,
            (w + 2 * pad - dil * (s - 1) - 1) // stride + 1)


def make_desc(xshape, K, R, S, stride, pad, dil, dtype, flags):
    N, H, W, Cin = xshape
    Ho, Wo = conv_out_size(H, W, R, S, stride, pad, dil)
    if Ho <= 0 or Wo <= 0:
        raise ValueError("convolution output would be empty for input %dx%d" % (H, W))
    return L.ConvDesc(N, H, W, Cin, K, R, S, stride, pad, dil, Ho, Wo, dtype, flags)


_workspaces = {}
_retired_workspaces = []      # replaced scratch buffers: captured CUDA graphs keep raw pointers into them


def bn_partials(C, device):
    """Uninitialised partial-sum buffer float[BN_PARTS][2C] (the producers write every slot)."""
    return torch.empty(L.BN_PARTS * 2 * C, dtype=torch.float32, device=device)


# ---- zeroed per-channel accumulators of the totals BatchNorm path --------------------------------
# Producers ADD into float[2C] vectors; they are cut from one arena per device that the training step
# zeroes once (`bn_arena_begin`), so a step costs one memset instead of two fills per layer.  Outside
# a managed step (or when the arena is too small) a fresh zero tensor is returned instead.
class _Arena:
    __slots__ = ("buf", "off", "demand", "retired")

    def __init__(self):
        self.buf, self.off, self.demand, self.retired = None, 0, 0, []


_arenas = {}
_totals_ok = {}
_BN_TOTALS = __import__("os").environ.get("B2POSE_BN_TOTALS", "1") != "0"


def bn_totals_supported(C_, dtype):
    key = (C_, dtype)
    v = _totals_ok.get(key)
    if v is None:
        v = _BN_TOTALS and dtype == torch.bfloat16 and bool(L.lib().b2_bn_totals_supported(C_, L.BF16))
        _totals_ok[key] = v
    return v


def bn_arena_begin(device):
    """Start a training step: zero the arena (growing it to last step's demand first; never while a
    CUDA graph is being captured, and retired buffers stay alive because captured graphs point at them)."""
    a = _arenas.setdefault(device.index, _Arena())
    need = max(a.demand, 1 << 16)
    if (a.buf is None or a.buf.numel() < need) and not torch.cuda.is_current_stream_capturing():
        if a.buf is not None:
            a.retired.append(a.buf)
        a.buf = torch.empty(need + need // 4, dtype=torch.float32, device=device)
    a.off = a.demand = 0
    if a.buf is not None:
        a.buf.zero_()


def bn_arena_end(device):
    a = _arenas.get(device.index)
    if a is not None:
        a.off = -1                    # slices are handed out only inside a managed step


def bn_totals(C_, device):
    """A zeroed float[2C] accumulator."""
    n = (2 * C_ + 63) // 64 * 64
    a = _arenas.get(device.index)
    if a is not None and a.off >= 0 and a.buf is not None:
        a.demand += n
        if a.off + n <= a.buf.numel():
            out = a.buf[a.off:a.off + 2 * C_]
            a.off += n
            return out
    return torch.zeros(2 * C_, dtype=torch.float32, device=device)


# ---- second stream for the depth branch of the two-stream (fusion) nets ---------------------------
# RGB trunk (conv1, layer1, layer2) and depth trunk (conv2, layer5, layer6) are independent up to the fusion
# convolution; the depth trunk is enqueued on its own stream so its kernels fill the launch gaps and tails of
# the other trunk's persistent kernels (autograd replays each node's backward on the stream of its forward,
# so the backward pass overlaps the same way).  B2POSE_TWO_STREAMS=0 disables it.
_branch_streams = {}
TWO_STREAMS = __import__("os").environ.get("B2POSE_TWO_STREAMS", "1") != "0"


_stream_slots = {}          # (device index, cuda_stream handle) -> private workspace slot of an auxiliary stream


def _aux_stream(device, name, slot):
    key = (device.index, name)
    st = _branch_streams.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _branch_streams[key] = st
        _stream_slots[(device.index, st.cuda_stream)] = slot
    return st


def branch_stream(device):
    return _aux_stream(device, "trunk", 2)


def shortcut_stream(device):
    """Stream for the down-sampling shortcut (1x1 conv + BN) of a block's first unit: it runs beside the block's
    main path.  One per trunk: the shortcut of a block on the trunk stream gets its own."""
    on_trunk = _stream_slots.get((device.index, torch.cuda.current_stream(device).cuda_stream)) == 2
    return _aux_stream(device, "shortcut_b" if on_trunk else "shortcut_a", 4 if on_trunk else 3)


def _aux_slot(device):
    return _stream_slots.get((device.index, torch.cuda.current_stream(device).cuda_stream), 0)


def workspace(nbytes, device, slot=0):
    """Grow-only scratch buffer per device, shared by all calls (the kernels that use it run on one
    stream, in order).  It is deliberately NOT keyed by stream: under CUDA-graph capture the current
    stream is the capture stream, and a buffer first allocated there would live in that graph's private
    memory pool -- a second graph (another Trainer in the same process) reusing it after the first
    graph was destroyed faulted.  The Trainer's eager warm-up steps grow it to its final size before
    any capture.  A buffer that has to grow AFTER a capture (an evaluation batch between training steps, a
    larger batch shape) is retired, not freed: replays of the captured graphs still write through its
    address (same policy as `_Arena.retired`)."""
    if nbytes == 0:
        return None, 0
    if slot == 0:
        slot = _aux_slot(device)     # concurrent trunk / shortcut streams have their own scratch
    ws = _workspaces.get((device.index, slot))
    if ws is None or ws.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("libb2pose workspace would have to grow during CUDA-graph capture; run the "
                               "same shapes eagerly once before capturing")
        if ws is not None:
            _retired_workspaces.append(ws)
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[(device.index, slot)] = ws
    return ws, ws.numel()


def filter_krsc(weight, dtype, shadow=None):
    """The [K,C,R,S] parameter as a KRSC-contiguous tensor of ``dtype`` (no copy when the parameter
    already lives in channels_last memory and has the compute dtype)."""
    wk = weight.detach().permute(0, 2, 3, 1)
    if dtype == weight.dtype:
        return wk if wk.is_contiguous() else wk.contiguous()
    if shadow is not None:
        return shadow
    if weight.dtype != torch.float32 or dtype != torch.bfloat16:
        raise TypeError("unsupported filter cast %s -> %s" % (weight.dtype, dtype))
    wk = wk if wk.is_contiguous() else wk.contiguous()
    out = torch.empty(wk.shape, dtype=torch.bfloat16, device=wk.device)
    L.call("b2_cast_f32_to_bf16", L.ptr(wk), L.ptr(out), wk.numel(), L.stream())
    return out


def to_nhwc(x, dtype):
    """NCHW fp32/bf16 network input -> NHWC tensor of the compute dtype (no autograd)."""
    L.require_cuda(x)
    N, Cc, H, W = x.shape
    x = x.detach()
    if Cc == 1 or x.permute(0, 2, 3, 1).is_contiguous():
        y = x.permute(0, 2, 3, 1)
        y = y if y.is_contiguous() else y.contiguous()
        if y.dtype == dtype:
            return y
        if y.dtype == torch.float32 and dtype == torch.bfloat16:
            out = torch.empty(y.shape, dtype=dtype, device=y.device)
            L.call("b2_cast_f32_to_bf16", L.ptr(y), L.ptr(out), y.numel(), L.stream())
            return out
        return y.to(dtype)
    x = x.float().contiguous()
    out = torch.empty((N, H, W, Cc), dtype=dtype, device=x.device)
    L.call("b2_nchw_to_nhwc", L.ptr(x), L.ptr(out), N, Cc, H, W, L.dt(out), L.stream())
    return out


def veil_from_depth(depth_nhwc):
    """veil = (depth != 0).float()  -- partial_depthnet.py:215."""
    N, H, W, _ = depth_nhwc.shape
    veil = torch.empty((N, H, W), dtype=torch.float32, device=depth_nhwc.device)
    L.call("b2_veil_from_depth", L.ptr(depth_nhwc), L.ptr(veil), veil.numel(), L.dt(depth_nhwc), L.stream())
    return veil


def _conv_fprop(desc, x, mask, wk, bias, want_ratio, want_stats, totals=False, keep_ws=None):
    dev = x.device
    y = torch.empty((desc.N, desc.Ho, desc.Wo, desc.K), dtype=x.dtype, device=dev)
    partial = bool(desc.flags & L.CONV_PARTIAL)
    mask_out = torch.empty((desc.N, desc.Ho, desc.Wo), dtype=torch.float32, device=dev) if partial else None
    ratio = torch.empty_like(mask_out) if (partial and want_ratio) else None
    sums = None
    if want_stats:
        sums = bn_totals(desc.K, dev) if totals else bn_partials(desc.K, dev)
    if keep_ws is not None:        # stem layers in a training step: a private buffer whose im2col matrix wgrad reuses
        ws, wsn = keep_ws, keep_ws.numel()
    else:
        ws, wsn = workspace(L.lib().b2_conv_workspace_bytes(C.byref(desc), 0), dev)
    if want_stats and totals:
        desc.flags |= L.CONV_BN_TOTALS
    try:
        L.call("b2_pconv_fprop", C.byref(desc), L.ptr(x), L.ptr(mask), L.ptr(wk), L.ptr(bias), L.ptr(y),
               L.ptr(mask_out), L.ptr(ratio), L.ptr(sums), L.ptr(ws), wsn, L.stream())
    finally:
        desc.flags &= ~L.CONV_BN_TOTALS
    return y, mask_out, ratio, sums


def _conv_dgrad(desc, dy, ratio, wk, mask, addend=None):
    """dx = dgrad(...) (+ addend).  When the tensor-core kernel can reduce-add into an existing
    tensor, `addend` (the gradient of a residual branch) is accumulated in place and returned;
    otherwise the sum is formed by a separate add."""
    dev = dy.device
    if addend is not None:
        desc.flags |= L.CONV_DX_ACCUMULATE
        fused = bool(L.lib().b2_conv_uses_tensor_cores(C.byref(desc), 1)) and addend.is_contiguous() \
            and addend.dtype == dy.dtype
        if fused:
            try:
                ws, wsn = workspace(L.lib().b2_conv_workspace_bytes(C.byref(desc), 1), dev)
                L.call("b2_pconv_dgrad", C.byref(desc), L.ptr(dy), L.ptr(ratio), L.ptr(wk), L.ptr(mask),
                       L.ptr(addend), L.ptr(ws), wsn, L.stream())
            finally:
                desc.flags &= ~L.CONV_DX_ACCUMULATE
            return addend
        desc.flags &= ~L.CONV_DX_ACCUMULATE
    dx = torch.empty((desc.N, desc.H, desc.W, desc.C), dtype=dy.dtype, device=dev)
    ws, wsn = workspace(L.lib().b2_conv_workspace_bytes(C.byref(desc), 1), dev)
    L.call("b2_pconv_dgrad", C.byref(desc), L.ptr(dy), L.ptr(ratio), L.ptr(wk), L.ptr(mask), L.ptr(dx),
           L.ptr(ws), wsn, L.stream())
    return dx if addend is None else dx + addend


# ---- weight gradients on a side stream ------------------------------------------------------------
# dW of a layer is needed only by the optimizer, so inside a managed training step (`wgrad_overlap_begin`
# ... `wgrad_overlap_end`, used by the Trainer) the wgrad kernels are enqueued on a second stream and run
# concurrently with the BatchNorm-backward / dgrad chain of the following layers: they fill the tails of
# the persistent kernels and overlap the HBM-bound BatchNorm streams.  Every tensor a side-stream kernel
# reads is kept alive until the streams join (the caching allocator would otherwise hand its memory to
# the main stream while the side stream still reads it); the side stream has its own scratch workspace.
class _Overlap:
    __slots__ = ("stream", "keep", "active", "main", "prepared")

    def __init__(self):
        self.stream, self.keep, self.active, self.main, self.prepared = None, [], False, None, False


_overlaps = {}
_WGRAD_OVERLAP = __import__("os").environ.get("B2POSE_WGRAD_OVERLAP", "1") != "0"


def wgrad_overlap_begin(device):
    if not _WGRAD_OVERLAP:
        return
    o = _overlaps.setdefault(device.index, _Overlap())
    if o.stream is None:
        o.stream = torch.cuda.Stream(device=device)
    o.keep, o.active, o.main, o.prepared = [], True, torch.cuda.current_stream(device), False


def wgrad_overlap_sync(device):
    """Forward / backward boundary: the filter transforms enqueued on the side stream during the forward pass
    (`_prepare_dgrad_filter`) must be complete before the first dgrad."""
    o = _overlaps.get(device.index)
    if o is not None and o.active and o.prepared:
        torch.cuda.current_stream(device).wait_stream(o.stream)
        o.prepared = False


def _prepare_dgrad_filter(desc, wk):
    """The tensor-core dgrad reads the filter flipped / transposed.  Inside a managed step the transform runs on the
    (otherwise idle) weight-gradient stream during the FORWARD pass, so the 86 small kernels leave the backward
    critical path.  Returns the prepared buffer or None (not a managed step / no tensor-core dgrad)."""
    o = _overlaps.get(wk.device.index)
    if o is None or not o.active:
        return None
    nbytes = L.lib().b2_pconv_dgrad_filter_bytes(C.byref(desc))
    if nbytes == 0:
        return None
    wt = torch.empty(nbytes, dtype=torch.uint8, device=wk.device)
    o.stream.wait_stream(torch.cuda.current_stream(wk.device))      # wk (and the block behind wt) are ready
    with torch.cuda.stream(o.stream):
        L.call("b2_pconv_dgrad_filter", C.byref(desc), L.ptr(wk), L.ptr(wt), nbytes, L.stream())
    o.prepared = True
    return wt


def wgrad_overlap_join(device):
    """Mid-step join (two-stage backward): the main stream waits for the side-stream wgrads issued so far; the
    overlap stays active for the second stage."""
    o = _overlaps.get(device.index)
    if o is None or not o.active:
        return
    o.main.wait_stream(o.stream)
    o.keep = []


def wgrad_overlap_end(device):
    """Join: the main stream waits for every side-stream wgrad; the kept tensors are released."""
    o = _overlaps.get(device.index)
    if o is None or not o.active:
        return
    o.main.wait_stream(o.stream)
    o.keep, o.active, o.main = [], False, None


def _conv_wgrad(desc, x, mask, dy, ratio, sink=None, col_ws=None):
    o = _overlaps.get(dy.device.index)
    if o is None or not o.active or sink is None:
        return _conv_wgrad_now(desc, x, mask, dy, ratio, sink, 0, col_ws)
    o.stream.wait_stream(torch.cuda.current_stream(dy.device))      # dy (and x) are ready
    o.keep.append((x, mask, dy, ratio, col_ws))
    with torch.cuda.stream(o.stream):
        _conv_wgrad_now(desc, x, mask, dy, ratio, sink, 1, col_ws)
    return None


def _conv_wgrad_now(desc, x, mask, dy, ratio, sink=None, ws_slot=0, col_ws=None):
    """dw (fp32, KRSC) accumulated into ``sink`` (a [K,C,R,S] channels_last view of the flat gradient
    buffer; returns None) or into a fresh zero tensor (returned as logical [K,C,R,S])."""
    dev = dy.device
    if sink is not None:
        dw = sink.permute(0, 2, 3, 1)
        if not dw.is_contiguous() or dw.dtype != torch.float32:
            raise RuntimeError("gradient sink must be an fp32 channels_last view")
    else:
        dw = torch.zeros((desc.K, desc.R, desc.S, desc.C), dtype=torch.float32, device=dev)
    if col_ws is not None:         # the fprop's private workspace: its im2col matrix is still in it
        desc.flags |= L.CONV_WS_HAS_COL
        ws, wsn = col_ws, col_ws.numel()
    else:
        ws, wsn = workspace(L.lib().b2_conv_workspace_bytes(C.byref(desc), 2), dev, ws_slot)
    try:
        L.call("b2_pconv_wgrad", C.byref(desc), L.ptr(x), L.ptr(mask), L.ptr(dy), L.ptr(ratio), L.ptr(dw),
               L.ptr(ws), wsn, L.stream())
    finally:
        desc.flags &= ~L.CONV_WS_HAS_COL
    return None if sink is not None else dw.permute(0, 3, 1, 2)


# --------------------------------------------------------------------------- convolution
class ConvFn(Function):
    """(Partial) convolution.  forward(x, mask, weight, bias, shadow, cfg) -> (y, mask_out).

    cfg = (stride, pad, dil, partial, premasked, force_ffma).  Backward follows the autograd of
    partial_conv.py:46-53: dRaw = dOut*ratio, dW = wgrad(x*m, dRaw), dX = dgrad(W, dRaw)*m,
    db = sum(dOut * mask_out).
    """

    @staticmethod
    def forward(ctx, x, mask, weight, bias, shadow, cfg, sinks=None):
        # sinks = (dw sink, db sink or None): views of the Trainer's flat gradient buffer that the kernels accumulate
        # into directly (no zero-filled temporaries, no AccumulateGrad add) -- the regressor's 3x3 convolution
        stride, pad, dil, partial, premasked, force_ffma = cfg
        ctx.sinks = sinks
        L.require_cuda(x, mask, weight, bias)
        ctx.set_materialize_grads(False)             # no zero-filled gradient for the mask output
        x = x.contiguous()
        K, _, R, S = weight.shape
        flags = (L.CONV_PARTIAL if partial else 0) | (L.CONV_X_PREMASKED if premasked else 0) | \
                (L.CONV_FORCE_FFMA if force_ffma else 0)
        desc = make_desc(x.shape, K, R, S, stride, pad, dil, L.dt(x), flags)
        wk = filter_krsc(weight, x.dtype, shadow)
        b32 = None if bias is None else bias.detach().float().contiguous()
        if partial:
            mask = mask.contiguous()
        ctx.self_masked = False
        if partial and not premasked and not force_ffma and x.dtype == torch.bfloat16 and R * S > 1:
            # The tensor-core kernels read x through TMA and cannot multiply by the mask on the way in: form x*mask
            # once here (one streaming pass) and take the pre-masked path -- the CUDA-core kernel this replaces ran
            # 20-40x slower than cuDNN on the isolated-layer sweep.  dgrad then scales its output rows by the mask.
            desc.flags |= L.CONV_X_PREMASKED
            if L.lib().b2_conv_uses_tensor_cores(C.byref(desc), 0):
                xm = torch.empty_like(x)
                L.call("b2_scale_rows", L.ptr(x), L.ptr(mask), L.ptr(xm), x.shape[0] * x.shape[1] * x.shape[2],
                       x.shape[3], L.dt(x), L.stream())
                x, ctx.self_masked = xm, True
            else:
                desc.flags &= ~L.CONV_X_PREMASKED
        y, mask_out, ratio, _ = _conv_fprop(desc, x, mask if partial else None, wk, b32, True, False)
        ctx.desc, ctx.has_bias, ctx.wdtype = desc, bias is not None, weight.dtype
        ctx.save_for_backward(x, mask if partial else None, wk, ratio, mask_out)
        if partial:
            ctx.mark_non_differentiable(mask_out)
        return y, mask_out

    @staticmethod
    def backward(ctx, dy, _dmask):
        if dy is None:
            return (None,) * 7
        x, mask, wk, ratio, mask_out = ctx.saved_tensors
        sinks = ctx.sinks
        desc = ctx.desc
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if ctx.self_masked:                  # x was masked here, not by the producer: dx = dgrad(...) * mask
                desc.flags &= ~L.CONV_X_PREMASKED
            try:
                dx = _conv_dgrad(desc, dy, ratio, wk, mask)
            finally:
                if ctx.self_masked:
                    desc.flags |= L.CONV_X_PREMASKED
        if ctx.needs_input_grad[2]:
            if sinks is not None:
                _conv_wgrad(desc, x, mask, dy, ratio, sink=sinks[0])          # accumulates into the flat buffer
            else:
                dw = _conv_wgrad(desc, x, mask, dy, ratio).to(ctx.wdtype)
        if ctx.has_bias and ctx.needs_input_grad[3]:
            to_sink = sinks is not None and sinks[1] is not None
            acc = sinks[1] if to_sink else torch.zeros(desc.K, dtype=torch.float32, device=dy.device)
            rows = desc.N * desc.Ho * desc.Wo
            L.call("b2_col_sum", L.ptr(dy), L.ptr(mask_out), L.ptr(acc), rows, desc.K, L.dt(dy), L.stream())
            db = None if to_sink else acc
        return dx, None, dw, db, None, None, None


class ConvBNFn(Function):
    """conv -> BatchNorm (+ residual) (+ ReLU) (* veil) as one autograd node.

    forward(x, mask, weight, shadow, gamma, beta, running_mean, running_var, residual, cfg)
      -> (z, mask_out)
    cfg = (stride, pad, dil, partial, premasked, relu, mask_output, training, momentum, eps, force_ffma)

    The conv epilogue (or a stats kernel) produces per-channel sum / sum-of-squares, one apply
    kernel normalises; backward is reduce + apply (which also folds the PartialConv ratio into
    dRaw) + dgrad + wgrad.  Mirrors conv->bn->relu of partial_depthnet.py:143-157.
    """

    @staticmethod
    def forward(ctx, x, mask, weight, shadow, gamma, beta, running_mean, running_var, residual, cfg, sinks=None,
                dx_holder=None, res_holder=None, park_holder=None, x2=None):
        # x2: the convolution runs over the channel concatenation [x, x2] (the fusion unit, fusionnet.py:137) without
        # materialising it -- the tensor-core kernels read / write the two halves through two tensor maps
        # (B2_CONV_X_CONCAT); where they cannot (fp32), the concatenation is formed here
        # dx_holder / res_holder: a dict shared by the first and the last conv+BN node of a residual
        # block with identity shortcut.  The last node parks the shortcut's gradient there instead of
        # returning it; the first node (whose backward always runs later) folds it into its dx with a
        # TMA reduce-add, so autograd never launches a separate add over the block input's gradient.
        # park_holder: the same dict handed to the block's down-sampling shortcut node (1x1 conv + BN on the block
        # input, created right AFTER the first node so that autograd runs its backward before the first node's): it parks
        # its own dx there, with an event of its stream, instead of returning it.
        stride, pad, dil, partial, premasked, relu, mask_output, training, momentum, eps, force_ffma = cfg
        L.require_cuda(x, mask, weight, gamma)
        # (without this, autograd hands backward a freshly zero-filled tensor for the unused mask gradient: one ATen
        #  fill kernel per PartialConv layer and step)
        ctx.set_materialize_grads(False)
        x = x.contiguous()
        K, _, R, S = weight.shape
        flags = (L.CONV_PARTIAL if partial else 0) | (L.CONV_X_PREMASKED if premasked else 0) | \
                (L.CONV_FORCE_FFMA if force_ffma else 0)
        ctx.concat = ctx.cat_split = None
        if x2 is not None:
            L.require_cuda(x2)
            x2 = x2.contiguous()
            full = tuple(x.shape[:3]) + (x.shape[3] + x2.shape[3],)
            probe = make_desc(full, K, R, S, stride, pad, dil, L.dt(x), flags | L.CONV_X_CONCAT)
            if x.shape == x2.shape and all(L.lib().b2_conv_uses_tensor_cores(C.byref(probe), op) for op in (0, 1, 2)):
                flags |= L.CONV_X_CONCAT
                ctx.concat = True
            else:
                ctx.cat_split = x.shape[3]
                x, x2 = torch.cat([x, x2], dim=3), None
        desc = make_desc(x.shape if x2 is None else full, K, R, S, stride, pad, dil, L.dt(x), flags)
        wk = filter_krsc(weight, x.dtype, shadow)
        if partial:
            mask = mask.contiguous()
        ctx.self_masked = False
        if partial and not premasked and not force_ffma and x.dtype == torch.bfloat16 and R * S > 1:
            # un-premasked 3x3 PartialConv (first conv of a BasicBlock): form x*mask once and stay on the tensor cores
            desc.flags |= L.CONV_X_PREMASKED
            if L.lib().b2_conv_uses_tensor_cores(C.byref(desc), 0):
                xm = torch.empty_like(x)
                L.call("b2_scale_rows", L.ptr(x), L.ptr(mask), L.ptr(xm), x.shape[0] * x.shape[1] * x.shape[2],
                       x.shape[3], L.dt(x), L.stream())
                x, ctx.self_masked = xm, True
            else:
                desc.flags &= ~L.CONV_X_PREMASKED
        fast = bn_totals_supported(K, x.dtype)       # totals path: no finalize kernels
        ctx.col_ws = None
        if x.shape[3] <= 4 and ctx.needs_input_grad[2] and L.lib().b2_conv_uses_tensor_cores(C.byref(desc), 2):
            # stem (im2col + GEMM): keep the workspace so that wgrad reuses the im2col matrix instead of rebuilding it
            nb = max(L.lib().b2_conv_workspace_bytes(C.byref(desc), 0), L.lib().b2_conv_workspace_bytes(C.byref(desc), 2))
            ctx.col_ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=x.device)
        y, mask_out, ratio, sums = _conv_fprop(desc, x if x2 is None else L.TensorPair(x, x2),
                                               mask if partial else None, wk, None, True, training, fast, ctx.col_ws)
        rows = desc.N * desc.Ho * desc.Wo
        z = torch.empty_like(y)
        mean = torch.empty(K, dtype=torch.float32, device=x.device)
        invstd = torch.empty_like(mean)
        if residual is not None:
            residual = residual.contiguous()
        row_mask = mask_out if (mask_output and partial) else None
        gate = None
        if fast:
            if relu and residual is not None:
                # ReLU gate of a residual layer as a bitmask (1/16 of z's bytes) for the two backward passes
                gate = torch.empty((rows * K // 8 + 15) // 16 * 16, dtype=torch.uint8, device=x.device)
            L.call("b2_bn_apply_totals", L.ptr(y), L.ptr(sums), rows, L.ptr(running_mean), L.ptr(running_var),
                   float(momentum), float(eps), int(training), L.ptr(gamma), L.ptr(beta), L.ptr(residual),
                   L.ptr(row_mask), int(relu), L.ptr(z), L.ptr(mean), L.ptr(invstd), L.ptr(gate), K, L.dt(y),
                   L.stream())
        else:
            L.call("b2_bn_finalize", L.ptr(sums), rows, K, L.ptr(running_mean), L.ptr(running_var), float(momentum),
                   float(eps), int(training), L.ptr(mean), L.ptr(invstd), L.stream())
            L.call("b2_bn_apply", L.ptr(y), L.ptr(mean), L.ptr(invstd), L.ptr(gamma), L.ptr(beta), L.ptr(residual),
                   L.ptr(row_mask), int(relu), L.ptr(z), rows, K, L.dt(y), L.stream())
        ctx.fast = fast
        ctx.wt = _prepare_dgrad_filter(desc, wk) if ctx.needs_input_grad[0] else None
        ctx.desc, ctx.relu, ctx.training, ctx.wdtype = desc, relu, training, weight.dtype
        ctx.has_res = residual is not None
        ctx.sinks = sinks
        ctx.dx_holder, ctx.res_holder, ctx.park_holder = dx_holder, res_holder, park_holder
        # z is only needed for the ReLU gate of residual layers; otherwise the gate is recomputed from y
        ctx.save_for_backward(x, mask if partial else None, wk, ratio, y,
                              z if (relu and residual is not None and gate is None) else None,
                              mean, invstd, gamma.detach(), beta.detach(), row_mask, gate, x2)
        if partial:
            ctx.mark_non_differentiable(mask_out)
        return z, mask_out

    @staticmethod
    def backward(ctx, dz, _dmask):
        if dz is None:                                   # nothing downstream used z
            return (None,) * 15
        x, mask, wk, ratio, y, z, mean, invstd, gamma, beta, row_mask, gate, x2 = ctx.saved_tensors
        desc = ctx.desc
        dz = dz.contiguous()
        dev = dz.device
        K, rows = desc.K, desc.N * desc.Ho * desc.Wo
        sinks = ctx.sinks
        dy = torch.empty_like(y)
        dres = torch.empty_like(y) if (ctx.has_res and ctx.needs_input_grad[8]) else None
        if sinks is not None:
            dgamma, dbeta = sinks[1], sinks[2]            # accumulate straight into the flat gradient buffer
        else:
            dgamma = torch.zeros(K, dtype=torch.float32, device=dev)
            dbeta = torch.zeros(K, dtype=torch.float32, device=dev)
        if ctx.fast:
            gsum = bn_totals(K, dev)
            L.call("b2_bn_bwd_reduce_totals", L.ptr(dz), L.ptr(z), L.ptr(y), L.ptr(mean), L.ptr(invstd), L.ptr(gamma),
                   L.ptr(beta), L.ptr(row_mask), int(ctx.relu), L.ptr(gsum), L.ptr(gate), rows, K, L.dt(dz), L.stream())
            L.call("b2_bn_bwd_apply_totals", L.ptr(dz), L.ptr(z), L.ptr(y), L.ptr(mean), L.ptr(invstd), L.ptr(gamma),
                   L.ptr(beta), L.ptr(gsum), L.ptr(row_mask), L.ptr(ratio), int(ctx.relu), int(ctx.training),
                   L.ptr(dy), L.ptr(dres), L.ptr(dgamma), L.ptr(dbeta), L.ptr(gate), rows, K, L.dt(dz), L.stream())
        else:
            parts = bn_partials(K, dev)
            L.call("b2_bn_bwd_reduce", L.ptr(dz), L.ptr(z), L.ptr(y), L.ptr(mean), L.ptr(invstd), L.ptr(gamma),
                   L.ptr(beta), L.ptr(row_mask), int(ctx.relu), L.ptr(parts), rows, K, L.dt(dz), L.stream())
            gsum = torch.empty(2 * K, dtype=torch.float32, device=dev)
            L.call("b2_bn_bwd_finalize", L.ptr(parts), K, L.ptr(gsum), L.ptr(dgamma), L.ptr(dbeta), L.stream())
            L.call("b2_bn_bwd_apply", L.ptr(dz), L.ptr(z), L.ptr(y), L.ptr(mean), L.ptr(invstd), L.ptr(gamma),
                   L.ptr(beta), L.ptr(gsum), L.ptr(row_mask), L.ptr(ratio), int(ctx.relu), int(ctx.training), L.ptr(dy),
                   L.ptr(dres), rows, K, L.dt(dz), L.stream())
        # dy now holds dRaw = dOut * ratio -> tell the conv kernels not to scale again
        desc.flags |= L.CONV_DY_PRESCALED
        dx = dw = dx2 = None
        if ctx.needs_input_grad[2]:        # first: on the side stream it then runs beside this layer's dgrad
            dw = _conv_wgrad(desc, x if x2 is None else L.TensorPair(x, x2), mask, dy, None,
                             sinks[0] if sinks is not None else None, ctx.col_ws)
            if dw is not None:
                dw = dw.to(ctx.wdtype)
        if ctx.concat:
            # gradient of the two halves of the (never materialised) concatenation, written by one dgrad launch
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[14]:
                dx, dx2 = torch.empty_like(x), torch.empty_like(x2)
                ws, wsn = workspace(L.lib().b2_conv_workspace_bytes(C.byref(desc), 1), dev)
                L.call("b2_pconv_dgrad", C.byref(desc), L.ptr(dy), None, L.ptr(wk), None, L.ptr(L.TensorPair(dx, dx2)),
                       L.ptr(ws), wsn, L.stream())
        elif ctx.needs_input_grad[0]:
            addend = None
            if ctx.dx_holder is not None:
                addend = ctx.dx_holder.pop("dres", None)
                ev = ctx.dx_holder.pop("ev", None)
                ctx.dx_holder["open"] = False            # a shortcut node that runs later returns its dx to autograd
                if ev is not None:                       # parked on the shortcut's stream
                    cur = torch.cuda.current_stream(dev)
                    cur.wait_event(ev)
                    addend.record_stream(cur)
            if ctx.self_masked:                  # x was masked in this node: dgrad scales its rows by the mask
                desc.flags &= ~L.CONV_X_PREMASKED
            wt = ctx.wt
            if wt is not None:                   # filter transform done during the forward pass, off this path
                desc.flags |= L.CONV_W_PREPARED
            try:
                dx = _conv_dgrad(desc, dy, None, wk if wt is None else wt, mask, addend)
            finally:
                desc.flags &= ~L.CONV_W_PREPARED
                if ctx.self_masked:
                    desc.flags |= L.CONV_X_PREMASKED
        desc.flags &= ~L.CONV_DY_PRESCALED
        if sinks is not None:
            dgamma = dbeta = None
        if dres is not None and ctx.res_holder is not None:
            ctx.res_holder["dres"] = dres          # delivered through the block's first node
            dres = None
        if dx is not None and ctx.park_holder is not None and ctx.park_holder.get("open", False) \
                and "dres" not in ctx.park_holder:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            ctx.park_holder["dres"], ctx.park_holder["ev"] = dx, ev
            dx = None
        if ctx.cat_split is not None and dx is not None:          # the concatenation was formed in forward (fp32 path)
            dx, dx2 = dx[..., :ctx.cat_split], dx[..., ctx.cat_split:]
        return dx, None, dw, None, dgamma, dbeta, None, None, dres, None, None, None, None, None, dx2


class MaxPoolFn(Function):
    """MaxPool2d(3, 2, 1) on x and (optionally) the veil in one launch."""

    @staticmethod
    def forward(ctx, x, veil):
        L.require_cuda(x, veil)
        ctx.set_materialize_grads(False)             # no zero-filled gradient for the pooled veil
        x = x.contiguous()
        N, H, W, Cc = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, Ho, Wo, Cc), dtype=x.dtype, device=x.device)
        arg = torch.empty((N, Ho, Wo, Cc), dtype=torch.uint8, device=x.device)
        vout = None
        if veil is not None:
            veil = veil.contiguous()
            vout = torch.empty((N, Ho, Wo), dtype=torch.float32, device=x.device)
        L.call("b2_maxpool3x3s2_fwd", L.ptr(x), L.ptr(veil), L.ptr(y), L.ptr(arg), L.ptr(vout), N, H, W, Cc,
               L.dt(x), L.stream())
        ctx.shape = (N, H, W, Cc)
        ctx.save_for_backward(arg)
        if vout is not None:
            ctx.mark_non_differentiable(vout)
        return y, vout

    @staticmethod
    def backward(ctx, dy, _dv):
        if dy is None:
            return None, None
        (arg,) = ctx.saved_tensors
        N, H, W, Cc = ctx.shape
        dy = dy.contiguous()
        dx = torch.empty((N, H, W, Cc), dtype=dy.dtype, device=dy.device)
        L.call("b2_maxpool3x3s2_bwd", L.ptr(dy), L.ptr(arg), L.ptr(dx), N, H, W, Cc, L.dt(dy), L.stream())
        return dx, None


# --------------------------------------------------------------------------- head
def _logit_layout(feat):
    """(tensor, layout): 0 = NHWC memory, 1 = NCHW memory, for a logical [N, D*J, H, W] tensor."""
    if feat.dim() != 4:
        raise ValueError("expected a [N, D*J, H, W] tensor")
    if feat.permute(0, 2, 3, 1).is_contiguous() and not (feat.is_contiguous() and feat.shape[1] != 1):
        return feat, 0
    if feat.is_contiguous():
        return feat, 1
    return feat.contiguous(), 1


class HeadFn(Function):
    """Fused to_heatmap + decode: logits [N, D*J, H, W] -> coords [N, J, 3] (x, y, z) * depth_range."""

    @staticmethod
    def forward(ctx, feat, depth, num_joints, depth_range):
        L.require_cuda(feat)
        feat, layout = _logit_layout(feat)
        N, CH, H, W = feat.shape
        if CH != depth * num_joints:
            raise ValueError("feature has %d channels, expected depth*num_joints = %d" % (CH, depth * num_joints))
        dev = feat.device
        coords = torch.empty((N, num_joints, 3), dtype=torch.float32, device=dev)
        vmax = torch.empty((N, num_joints), dtype=torch.float32, device=dev)
        vsum = torch.empty_like(vmax)
        L.call("b2_head_fwd", L.ptr(feat), N, num_joints, depth, H, W, layout, L.dt(feat), float(depth_range),
               L.ptr(coords), L.ptr(vmax), L.ptr(vsum), L.stream())
        ctx.cfg = (N, num_joints, depth, H, W, layout, float(depth_range))
        ctx.save_for_backward(feat, coords, vmax, vsum)
        return coords

    @staticmethod
    def backward(ctx, dcoords):
        feat, coords, vmax, vsum = ctx.saved_tensors
        N, J, D, H, W, layout, rng = ctx.cfg
        dcoords = dcoords.float().contiguous()
        dfeat = torch.empty_like(feat)
        L.call("b2_head_bwd", L.ptr(feat), L.ptr(dcoords), L.ptr(coords), L.ptr(vmax), L.ptr(vsum), N, J, D, H, W,
               layout, L.dt(feat), rng, L.ptr(dfeat), L.stream())
        return dfeat, None, None, None


class ToHeatmapFn(Function):
    """utils.to_heatmap (utils.py:154-175): [N, D*J, H, W] -> softmax heat-map [N, J, H, W, D] fp32."""

    @staticmethod
    def forward(ctx, feat, depth, num_joints, height, width):
        L.require_cuda(feat)
        feat = feat.reshape(-1, depth * num_joints, height, width)
        feat, layout = _logit_layout(feat)
        N = feat.shape[0]
        dev = feat.device
        heat = torch.empty((N, num_joints, height, width, depth), dtype=torch.float32, device=dev)
        scratch = torch.empty((N, num_joints, 3), dtype=torch.float32, device=dev)
        L.call("b2_heatmap_softmax", L.ptr(feat), N, num_joints, depth, height, width, layout, L.dt(feat),
               L.ptr(heat), L.ptr(scratch), L.stream())
        ctx.cfg = (N, num_joints, depth, height, width, layout, feat.dtype)
        ctx.feat_like = feat
        ctx.save_for_backward(heat)
        return heat

    @staticmethod
    def backward(ctx, dheat):
        (heat,) = ctx.saved_tensors
        N, J, D, H, W, layout, dtype = ctx.cfg
        dheat = dheat.float().contiguous()
        dfeat = torch.empty_like(ctx.feat_like)
        L.call("b2_heatmap_softmax_bwd", L.ptr(heat), L.ptr(dheat), N, J, D, H, W, layout, L.dt(dfeat),
               L.ptr(dfeat), L.stream())
        return dfeat, None, None, None, None


class DecodeFn(Function):
    """utils.decode (utils.py:178-194): heat-map [N, J, H, W, D] -> [N, J, 3] * depth_range."""

    @staticmethod
    def forward(ctx, heat, depth_range):
        L.require_cuda(heat)
        heat = heat.float().contiguous()
        N, J, H, W, D = heat.shape
        coords = torch.empty((N, J, 3), dtype=torch.float32, device=heat.device)
        L.call("b2_heatmap_decode", L.ptr(heat), N, J, D, H, W, float(depth_range), L.ptr(coords), L.stream())
        ctx.cfg = (N, J, D, H, W, float(depth_range))
        return coords

    @staticmethod
    def backward(ctx, dcoords):
        N, J, D, H, W, rng = ctx.cfg
        dcoords = dcoords.float().contiguous()
        dheat = torch.empty((N, J, H, W, D), dtype=torch.float32, device=dcoords.device)
        L.call("b2_heatmap_decode_bwd", L.ptr(dcoords), N, J, D, H, W, rng, L.ptr(dheat), L.stream())
        return dheat, None


CRITERIA = {"SmoothL1": 0, "L1": 1, "MSE": 2}


class PoseLossFn(Function):
    """Root-relative shift + masked mean loss (depth_train.py:397-405) -> (loss, spec_cam)."""

    @staticmethod
    def forward(ctx, coords, true_cam, valid, key_index, loss_div, criterion):
        L.require_cuda(coords, true_cam, valid)
        coords = coords.float().contiguous()
        true_cam = true_cam.float().contiguous()
        valid = valid.to(torch.uint8).contiguous()
        N, J, _ = coords.shape
        dev = coords.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        spec = torch.empty_like(coords)
        dcoords = torch.empty_like(coords)
        L.call("b2_pose_loss", L.ptr(coords), L.ptr(true_cam), L.ptr(valid), N, J, int(key_index), float(loss_div),
               CRITERIA[criterion], L.ptr(loss), L.ptr(spec), L.ptr(dcoords), L.stream())
        ctx.save_for_backward(dcoords)
        ctx.mark_non_differentiable(spec)
        return loss, spec

    @staticmethod
    def backward(ctx, dloss, _dspec):
        (dcoords,) = ctx.saved_tensors
        return dcoords * dloss, None, None, None, None, None


MIMIC_MODES = {"l2": 0, "sigmoid": 1, "bce": 2}


class MimicLossFn(Function):
    """Feature-mimic loss of the distillation step (Trainer.distill, depth_train.py:115-129).

    forward(teach_last, last_feat, atten_map, mode) -> 0-dim fp32 loss; gradient flows to last_feat only
    (the teacher runs under no_grad, depth_train.py:194-195)."""

    @staticmethod
    def forward(ctx, teach, student, atten, mode):
        L.require_cuda(teach, student, atten)
        if teach.shape != student.shape or teach.dim() != 4:
            raise ValueError("teacher / student features must be [N, C, H, W] of one shape, got %s and %s"
                             % (tuple(teach.shape), tuple(student.shape)))
        N, Cc, H, W = student.shape
        if atten.numel() != N * H * W:
            raise ValueError("attention map must hold N*H*W = %d values, got %s" % (N * H * W, tuple(atten.shape)))
        student, layout = _logit_layout(student)
        teach = teach.detach().to(student.dtype)
        if layout == 0:
            if not teach.permute(0, 2, 3, 1).is_contiguous():
                teach = teach.contiguous(memory_format=torch.channels_last)
        else:
            teach = teach.contiguous()
        atten = atten.detach().float().contiguous()
        dev = student.device
        partials = torch.empty(N * L.MIMIC_PARTS, dtype=torch.float32, device=dev)
        scale = torch.empty(N, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        L.call("b2_mimic_loss_fwd", L.ptr(teach), L.ptr(student), L.ptr(atten), N, Cc, H * W, layout, L.dt(student),
               int(mode), L.ptr(partials), L.ptr(scale), L.ptr(loss), L.stream())
        ctx.cfg = (N, Cc, H * W, layout, int(mode))
        ctx.save_for_backward(teach, student, atten, scale)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        teach, student, atten, scale = ctx.saved_tensors
        N, Cc, HW, layout, mode = ctx.cfg
        dloss = dloss.float().contiguous()
        ds = torch.empty_like(student)
        L.call("b2_mimic_loss_bwd", L.ptr(teach), L.ptr(student), L.ptr(atten), L.ptr(scale), L.ptr(dloss), N, Cc, HW,
               layout, L.dt(student), mode, L.ptr(ds), L.stream())
        return None, ds, None, None


def attention_map(image_coords, side_in, side_out):
    """utils.get_attention on device: image_coords [N, J, 2] (x, y) -> [N, 1, side_out, side_out] fp32."""
    L.require_cuda(image_coords)
    c = image_coords.detach().float().contiguous()
    N, J, two = c.shape
    if two != 2:
        raise ValueError("image_coords must be [N, J, 2]")
    out = torch.empty((N, 1, side_out, side_out), dtype=torch.float32, device=c.device)
    L.call("b2_attention_map", L.ptr(c), N, J, int(side_in), int(side_out), L.ptr(out), L.stream())
    return out


def unproject_depth(img, intrinsic):
    """utils.to_depth on device: img [..., H, W] fp32 CUDA tensor, intrinsic 3x3 (host array-like)."""
    import numpy as np
    L.require_cuda(img)
    img = img.float().contiguous()
    H, W = img.shape[-2:]
    K = np.asarray(intrinsic, np.float32)
    kinv = np.linalg.inv(K[:2, :2]).astype(np.float32).reshape(-1)
    kin = (C.c_float * 4)(*[float(v) for v in kinv])
    cc = (C.c_float * 2)(float(K[0, 2]), float(K[1, 2]))
    out = torch.empty_like(img)
    L.call("b2_unproject_depth", L.ptr(img), L.ptr(out), img.numel() // (H * W), H, W, kin, cc, L.stream())
    return out
