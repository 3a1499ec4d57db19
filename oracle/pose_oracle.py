"""CPU oracle for the depth-stream hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU (torch fp32 ATen ops + a numpy loop version for
small cases), what the reference computes on the hot path named in
BASELINE.json.  It is the checker: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product package never imports anything from ``oracle/``.

Parity status: the reference ships no tests or golden vectors,
so parity is *unpinned by the reference's own tests*.  The oracle is pinned
instead against outputs of the reference code itself, executed in the build
container by ``oracle/make_golden.py`` (which imports /root/reference) and
committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` re-checks
this file against those fixtures on every run.

Design: everything is a pure function of a *state dict* that uses the
reference's parameter names (``conv1.weight``, ``layer2.0.downsample.1.bias``,
``fusion.conv.weight``, ``regressor.bias`` ...).  There are no nn.Modules here;
the networks are driven by a small table (``stage_table``) so the file does not
mirror the reference's class structure.

Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# a2 / a3  partial convolution   (partial_conv.py:32-58)
# ----------------------------------------------------------------------------


def partial_conv(x, mask, weight, bias=None, stride=1, padding=0, dilation=1):
    """Mask-renormalised convolution, single-channel mask (partial_conv.py:32-58).

    x      [N,C,H,W]  (any float dtype), mask [N,1,H,W] (fp32 0/1)
    weight [K,C,R,S], bias [K] or None
    returns (out [N,K,H',W'] in x.dtype, mask_out [N,1,H',W'] in mask.dtype)

    Gradients flow to x, weight and bias through ordinary autograd; the mask
    algebra is constant (partial_conv.py:35 ``torch.no_grad``).
    """
    R, S = weight.shape[2], weight.shape[3]
    window = float(R * S)                                   # partial_conv.py:28 (single-channel ones)
    with torch.no_grad():
        box = torch.ones(1, 1, R, S, dtype=mask.dtype)
        cnt = F.conv2d(mask, box, None, stride, padding, dilation)      # :39
        ratio = window / (cnt + 1e-6)                                   # :41
        mask_out = cnt.clamp(0, 1)                                      # :43
        ratio = (ratio * mask_out).to(x.dtype)                          # :44
    raw = F.conv2d(x * mask.to(x.dtype), weight, bias, stride, padding, dilation)   # :46
    if bias is None:
        out = raw * ratio                                               # :53
    else:
        b = bias.view(1, -1, 1, 1)
        out = ((raw - b) * ratio + b) * mask_out                        # :49-51
    return out, mask_out


def partial_conv_loops(x, mask, weight, bias=None, stride=1, padding=0, dilation=1):
    """Same op with explicit numpy loops (no library convolution) for tiny cases.

    Independent of ATen's conv so that the torch-based oracle above is itself
    cross-checked.  float32 throughout, accumulation in float64 then rounded.
    """
    x = np.asarray(x, np.float32)
    mask = np.asarray(mask, np.float32)
    w = np.asarray(weight, np.float32)
    N, C, H, W = x.shape
    K, _, R, S = w.shape
    Ho = (H + 2 * padding - dilation * (R - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (S - 1) - 1) // stride + 1
    out = np.zeros((N, K, Ho, Wo), np.float32)
    mout = np.zeros((N, 1, Ho, Wo), np.float32)
    for n in range(N):
        for oh in range(Ho):
            for ow in range(Wo):
                cnt = np.float32(0)
                acc = np.zeros(K, np.float64)
                for r in range(R):
                    ih = oh * stride - padding + r * dilation
                    if ih < 0 or ih >= H:
                        continue
                    for s in range(S):
                        iw = ow * stride - padding + s * dilation
                        if iw < 0 or iw >= W:
                            continue
                        m = mask[n, 0, ih, iw]
                        cnt += m
                        if m != 0:
                            acc += w[:, :, r, s].astype(np.float64) @ (x[n, :, ih, iw].astype(np.float64) * m)
                ratio = np.float32(R * S) / (cnt + np.float32(1e-6))
                mo = np.float32(min(max(cnt, 0.0), 1.0))
                ratio = np.float32(ratio * mo)
                raw = acc.astype(np.float32)
                if bias is None:
                    out[n, :, oh, ow] = raw * ratio
                else:
                    b = np.asarray(bias, np.float32)
                    out[n, :, oh, ow] = (raw * ratio + b) * mo       # (raw+b-b)*ratio+b
                mout[n, 0, oh, ow] = mo
    return out, mout


def renorm_ratio(window: int, count: int) -> float:
    """fp32 value of ``window/(count+1e-6) * clamp(count,0,1)`` (partial_conv.py:41-44)."""
    c = np.float32(count)
    r = np.float32(window) / (c + np.float32(1e-6))
    return float(np.float32(r * np.float32(min(max(c, 0.0), 1.0))))


# ----------------------------------------------------------------------------
# a9 / a10  volumetric heat-map head   (utils.py:154-194)
# ----------------------------------------------------------------------------


def to_heatmap(feat, depth, num_joints, height, width):
    """Softmax over the H*W*D voxels of every (sample, joint) (utils.py:154-175).

    feat [N, depth*num_joints, H, W] with channel = d*num_joints + j
    returns [N, num_joints, H, W, depth]
    """
    vol = feat.reshape(-1, depth, num_joints, height, width).permute(0, 2, 3, 4, 1)
    vol = vol.reshape(vol.shape[0], num_joints, -1)
    vol = (vol - vol.amax(dim=2, keepdim=True)).exp()
    vol = vol / vol.sum(dim=2, keepdim=True)
    return vol.reshape(-1, num_joints, height, width, depth)


def decode(heatmap, depth_range):
    """Soft-argmax: marginal expectations on linspace(0,2,n) grids (utils.py:178-194).

    returns [N, J, 3] ordered (x = W axis, y = H axis, z = D axis), times depth_range.
    """
    out = []
    for keep_axis in (3, 2, 4):                       # x<-W, y<-H, z<-D
        others = tuple(a for a in (2, 3, 4) if a != keep_axis)
        marg = heatmap.sum(dim=others)
        grid = torch.linspace(0.0, 2.0, marg.shape[-1], device=marg.device).view(1, 1, -1)
        out.append((grid * marg).sum(dim=2))
    return torch.stack(out, dim=2) * depth_range


def pose_loss(cam_feat, true_cam, true_val, *, depth, num_joints, side_out, depth_range,
              key_index, loss_div=10.0, criterion="SmoothL1"):
    """Head + root-relative shift + masked loss (depth_train.py:395-405, train.py:166-174).

    returns (loss, spec_cam [N,J,3])
    """
    heat = to_heatmap(cam_feat.float(), depth, num_joints, side_out, side_out)
    rel = decode(heat, depth_range)
    rel = rel - rel[:, key_index:key_index + 1]
    spec = rel + true_cam[:, key_index:key_index + 1]
    sel = true_val.reshape(-1).bool()
    a = spec.reshape(-1, 3)[sel] / loss_div
    b = true_cam.reshape(-1, 3)[sel] / loss_div
    if criterion == "SmoothL1":
        loss = F.smooth_l1_loss(a, b, reduction="mean")
    elif criterion == "L1":
        loss = F.l1_loss(a, b, reduction="mean")
    elif criterion == "MSE":
        loss = F.mse_loss(a, b, reduction="mean")
    else:
        raise ValueError(criterion)
    return loss, spec


def mpjpe(spec_cam, true_cam, valid):
    """``cam_mean`` of utils.analyze (utils.py:253-262): mean joint distance over valid joints."""
    d = np.linalg.norm(np.asarray(spec_cam) - np.asarray(true_cam), axis=-1).reshape(-1)
    return float(d[np.asarray(valid).reshape(-1).astype(bool)].mean())


# ----------------------------------------------------------------------------
# a13  ray-length -> z-depth   (utils.py:68-75, cameralib.py:188-200 no-distortion branch)
# ----------------------------------------------------------------------------


def to_depth(image, intrinsic):
    """image [H,W] float, intrinsic 3x3.  Returns image / sqrt(xn^2 + yn^2 + 1 + 1).

    The reference's ``image_to_camera`` already appends the homogeneous 1, and
    ``to_depth`` then adds another 1 under the root (utils.py:75) -- kept as is.
    Point coordinates (cameralib.py:190) and the intrinsic matrix (cameralib.py:93) are both
    float32, so the whole expression evaluates in float32 for a float32 image.
    """
    image = np.asarray(image)
    K = np.asarray(intrinsic, np.float32)
    H, W = image.shape
    u, v = np.meshgrid(range(W), range(H))
    pts = np.stack([u, v], axis=-1).reshape(-1, 2).astype(np.float32)
    nrm = (pts - K[:2, 2]) @ np.linalg.inv(K[:2, :2]).T
    hom = np.concatenate([nrm, np.ones((nrm.shape[0], 1), nrm.dtype)], axis=1).reshape(H, W, 3)
    return image / np.sqrt(np.sum(hom ** 2, axis=-1) + 1)


# ----------------------------------------------------------------------------
# a4-a8  networks, table driven
# ----------------------------------------------------------------------------

KINDS = ("depthnet", "partial_depthnet", "fusionnet", "partial_fusionnet", "resnet")
DEPTHS = {"resnet18": ("basic", (2, 2, 2, 2)), "resnet50": ("bottleneck", (3, 4, 6, 3))}


def net_config(**kw):
    """Namespace with the fields model constructors read (opts.py; SURVEY §5 'Config')."""
    base = dict(stride=16, depth=16, num_joints=17, side_in=256, depth_only=True,
                early_dist=False, skip_relu=False, extra_channel=False, joint_space=False,
                pretrain=False)
    base.update(kw)
    return SimpleNamespace(**base)


def stage_strides(net_stride):
    """Per-stage stride / dilation (partial_depthnet.py:169-175, same in every net file)."""
    lg = math.log2(net_stride)
    s2 = int(min(max(lg, 2), 3) - 1)
    s3 = int(min(max(lg, 3), 4) - 2)
    s4 = int(min(max(lg, 4), 5) - 3)
    d2 = 3 - s2
    d3 = d2 * (3 - s3)
    d4 = d3 * (3 - s4)
    return (1, s2, s3, s4), (1, d2, d3, d4)


def stage_table(model, cfg):
    """[(layer_name, planes, n_blocks, stride, dilation)] for layer1..layer4."""
    _, counts = DEPTHS[model]
    strides, dils = stage_strides(cfg.stride)
    return [("layer%d" % (i + 1), 64 << i, counts[i], strides[i], dils[i]) for i in range(4)]


def _block_convs(block, inplanes, planes, stride, dilation):
    """[(suffix, cin, cout, k, stride, pad, dil)] for the residual branch of one block."""
    if block == "bottleneck":      # partial_depthnet.py:86-113
        return [("1", inplanes, planes, 1, 1, 0, 1),
                ("2", planes, planes, 3, stride, dilation, dilation),
                ("3", planes, planes * 4, 1, 1, 0, 1)]
    return [("1", inplanes, planes, 3, stride, dilation, dilation),     # partial_depthnet.py:20-38
            ("2", planes, planes, 3, 1, 1, 1)]


def param_shapes(kind, model, cfg):
    """Ordered {name: shape} of every parameter/buffer, reference naming."""
    block, _ = DEPTHS[model]
    exp = 4 if block == "bottleneck" else 1
    shapes = {}

    def bn(prefix, c):
        shapes[prefix + ".weight"] = (c,)
        shapes[prefix + ".bias"] = (c,)
        shapes[prefix + ".running_mean"] = (c,)
        shapes[prefix + ".running_var"] = (c,)
        shapes[prefix + ".num_batches_tracked"] = ()

    def stages(names):
        inpl = 64
        for (lname, planes, nblk, stride, dil), outname in zip(stage_table(model, cfg), names):
            if outname is None:
                inpl = planes * exp
                continue
            for b in range(nblk):
                pre = "%s.%d" % (outname, b)
                s, d = (stride, dil) if b == 0 else (1, 1)
                for suf, ci, co, k, _, _, _ in _block_convs(block, inpl if b == 0 else planes * exp, planes, s, d):
                    shapes["%s.conv%s.weight" % (pre, suf)] = (co, ci, k, k)
                    bn("%s.bn%s" % (pre, suf), co)
                if b == 0 and (stride != 1 or inpl != planes * exp):
                    shapes[pre + ".downsample.0.weight"] = (planes * exp, inpl, 1, 1)
                    bn(pre + ".downsample.1", planes * exp)
            inpl = planes * exp

    fusion = kind in ("fusionnet", "partial_fusionnet")
    if fusion:
        shapes["conv1.weight"] = (64, 3, 7, 7)
        shapes["conv2.weight"] = (64, 1, 7, 7)
        bn("bn1", 64)
        bn("bn2", 64)
        stages(["layer1", "layer2", None, None])
        shapes["fusion.conv.weight"] = (128 * exp, 256 * exp, 1, 1)
        bn("fusion.bn", 128 * exp)
        stages([None, None, "layer3", "layer4"])
        stages(["layer5", "layer6", None, None])
    else:
        if kind == "resnet":
            cin = 4 if cfg.extra_channel else 3
        elif kind == "partial_depthnet":
            cin = 1
        else:
            cin = 1 if cfg.depth_only else 3
        shapes["conv1.weight"] = (64, cin, 7, 7)
        bn("bn1", 64)
        stages(["layer1", "layer2", "layer3", "layer4"])
    head = "cam_regressor" if kind == "resnet" else "regressor"
    shapes[head + ".weight"] = (cfg.depth * cfg.num_joints, 512 * exp, 3, 3)
    shapes[head + ".bias"] = (cfg.depth * cfg.num_joints,)
    if kind == "resnet" and cfg.joint_space:
        shapes["mat_regressor.weight"] = (cfg.num_joints, 512 * exp, 3, 3)
        shapes["mat_regressor.bias"] = (cfg.num_joints,)
    return shapes


def init_state(kind, model, cfg, seed=0, dtype=torch.float32):
    """Deterministic state dict (CPU generator, key order of ``param_shapes``).

    Conv weights ~ N(0, sqrt(2/(k*k*out))) -- the fan_out Kaiming rule every net
    file uses (partial_depthnet.py:187-189, fusionnet.py:187-190); BN affine =
    (1, 0) (:191-193); regressors ~ U(+-1/sqrt(fan_in)) like nn.Conv2d's default
    (created after the init loop, partial_depthnet.py:195).  This is *an* init
    with the reference's distribution, used identically by oracle, golden
    generator and product tests; it does not try to reproduce torch's global
    RNG stream.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shp in param_shapes(kind, model, cfg).items():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros((), dtype=torch.long)
        elif name.endswith("running_mean") or (name.endswith(".bias") and "regressor" not in name):
            sd[name] = torch.zeros(shp, dtype=dtype)
        elif name.endswith("running_var") or (len(shp) == 1 and name.endswith(".weight")):
            sd[name] = torch.ones(shp, dtype=dtype)
        elif "regressor" in name:
            fan_in = 512 * (4 if DEPTHS[model][0] == "bottleneck" else 1) * 9
            bound = 1.0 / math.sqrt(fan_in)
            sd[name] = (torch.rand(shp, generator=g, dtype=dtype) * 2 - 1) * bound
        else:
            co, _, k, _ = shp
            sd[name] = torch.randn(shp, generator=g, dtype=dtype) * math.sqrt(2.0 / (k * k * co))
    return sd


class _RoundBF16(torch.autograd.Function):
    """Round to bf16 in the forward AND the backward pass: the device path stores every activation (conv output y,
    block output z) and every activation gradient (dy, dx) as bf16; everything between two stores is fp32."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def round_bf16(x):
    """The `act_round` hook that restates the bf16 storage contract of the device path (see net_forward)."""
    return _RoundBF16.apply(x)


def _ident(x):
    return x


def _bn(sd, prefix, x, training, momentum=0.1, eps=1e-5):
    out = F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                       sd[prefix + ".weight"], sd[prefix + ".bias"], training, momentum, eps)
    if training and (prefix + ".num_batches_tracked") in sd:
        sd[prefix + ".num_batches_tracked"] += 1
    return out


def _keep(trace, **rec):
    """Append one record to a trace list, keeping the gradients of the recorded activations after backward()."""
    if trace is not None:
        for v in rec.values():
            if torch.is_tensor(v) and v.requires_grad and not v.is_leaf:
                v.retain_grad()
        trace.append(rec)


def _run_stage(sd, lname, block, planes, nblk, stride, dil, x, veil, training, skip_last_relu=False, rnd=_ident,
               trace=None):
    """One ResNet stage.  ``veil`` None -> plain convs; else partial convs threading the veil
    (partial_depthnet.py:62-75,140-157; residual branch plain on the *unmasked* input)."""
    exp = 4 if block == "bottleneck" else 1
    for b in range(nblk):
        pre = "%s.%d" % (lname, b)
        s, d = (stride, dil) if b == 0 else (1, 1)
        inpl = x.shape[1]
        res = x
        out = x
        x_in, veil_in = x, veil
        convs = _block_convs(block, inpl, planes, s, d)
        for i, (suf, _, _, k, cs, cp, cd) in enumerate(convs):
            w = sd["%s.conv%s.weight" % (pre, suf)]
            if veil is None:
                out = rnd(F.conv2d(out, w, None, cs, cp, cd))
            else:
                out, veil = partial_conv(out, veil, w, None, cs, cp, cd)
                out = rnd(out)
            out = _bn(sd, "%s.bn%s" % (pre, suf), out, training)
            if i + 1 < len(convs):
                out = rnd(F.relu(out))
        if (pre + ".downsample.0.weight") in sd:
            res = rnd(F.conv2d(res, sd[pre + ".downsample.0.weight"], None, s))
            res = rnd(_bn(sd, pre + ".downsample.1", res, training))
        out = out + res
        if not (skip_last_relu and b == nblk - 1):
            out = F.relu(out)
        x = rnd(out)
        _keep(trace, name=pre, x=x_in, veil=veil_in, out=x, veil_out=veil)
    return x, veil


def net_forward(sd, kind, model, cfg, x, y=None, training=True, act_round=None, trace=None):
    """Forward of any of the five nets.  Returns (z, last_feat) -- or, for kind
    'resnet', cam_feat / (cam_feat, mat_feat) like resnet.py:204-209.

    ``act_round`` (default: none, the reference's fp32 arithmetic) is applied wherever the device's bf16 mode stores
    an activation (``trace``: optional list that receives one record per stem / residual block / fusion / regressor
    with its input, output and -- after backward() -- their gradients, for block-level parity tests): after every convolution (the raw / renormalised output the BatchNorm statistics are taken from)
    and after every BatchNorm(+residual)(+ReLU) / pooling result.  ``act_round=round_bf16`` therefore restates the
    reference algorithm under the bf16-storage / fp32-accumulate contract (what torch.autocast does to the
    reference itself), which is what the bf16 parity tests compare against: training-mode BatchNorm at random
    initialisation amplifies any perturbation ~200x through ResNet-50 (tests/golden/bf16_sensitivity.npz), so an
    fp32 evaluation is not a meaningful reference for bf16 training-mode outputs -- the reference's own
    bf16-autocast forward is 0.6 relative away from its fp32 forward.

    depthnet.py:188-200, partial_depthnet.py:213-229, fusionnet.py:221-240,
    partial_fusionnet.py:250-274 (with the documented stem fix: RGB stem plain,
    depth stem partial -- SURVEY.md note 3).
    """
    block, _ = DEPTHS[model]
    tab = stage_table(model, cfg)
    partial = kind.startswith("partial_")
    skip = bool(getattr(cfg, "skip_relu", False)) and kind in ("depthnet", "fusionnet")
    pool = lambda t: F.max_pool2d(t, 3, 2, 1)
    rnd = _ident if act_round is None else act_round
    x = rnd(x)
    y = rnd(y) if y is not None else None

    def stem(inp, conv, bn, is_partial):
        if is_partial:
            veil = (inp != 0).float()
            t, veil = partial_conv(inp, veil, sd[conv + ".weight"], None, 2, 3, 1)
            t = pool(rnd(F.relu(_bn(sd, bn, rnd(t), training))))
            veil = pool(veil)
            _keep(trace, name=conv, x=inp, veil=None, out=t, veil_out=veil)
            return t, veil
        t = rnd(F.conv2d(inp, sd[conv + ".weight"], None, 2, 3))
        t = pool(rnd(F.relu(_bn(sd, bn, t, training))))
        _keep(trace, name=conv, x=inp, veil=None, out=t, veil_out=None)
        return t, None

    if kind in ("fusionnet", "partial_fusionnet"):
        a, _ = stem(x, "conv1", "bn1", False)
        b, veil = stem(y, "conv2", "bn2", partial)
        for (ln, planes, nblk, s, d), dn in zip(tab[:2], ("layer5", "layer6")):
            a, _ = _run_stage(sd, ln, block, planes, nblk, s, d, a, None, training, rnd=rnd, trace=trace)
            b, veil = _run_stage(sd, dn, block, planes, nblk, s, d, b, veil, training, rnd=rnd, trace=trace)
        f = rnd(F.conv2d(torch.cat([a, b], dim=1), sd["fusion.conv.weight"]))
        f = rnd(F.relu(_bn(sd, "fusion.bn", f, training)))
        _keep(trace, name="fusion", x=a, x2=b, veil=None, out=f, veil_out=None)
    else:
        f, veil = stem(x, "conv1", "bn1", partial)
        for (ln, planes, nblk, s, d) in tab[:2]:
            f, veil = _run_stage(sd, ln, block, planes, nblk, s, d, f, veil, training, rnd=rnd, trace=trace)

    ln, planes, nblk, s, d = tab[2]
    m, _ = _run_stage(sd, ln, block, planes, nblk, s, d, f, None, training, skip_last_relu=skip, rnd=rnd, trace=trace)
    ln, planes, nblk, s, d = tab[3]
    n, _ = _run_stage(sd, ln, block, planes, nblk, s, d, F.relu(m) if skip else m, None, training,
                      skip_last_relu=skip, rnd=rnd, trace=trace)
    top = F.relu(n) if skip else n
    if kind == "resnet":
        cam = F.conv2d(top, sd["cam_regressor.weight"], sd["cam_regressor.bias"], 1, 1)
        if "mat_regressor.weight" in sd:
            return cam, F.conv2d(top, sd["mat_regressor.weight"], sd["mat_regressor.bias"], 1, 1)
        return cam
    z = rnd(F.conv2d(top, sd["regressor.weight"], sd["regressor.bias"], 1, 1))
    _keep(trace, name="regressor", x=top, veil=None, out=z, veil_out=None)
    last = m if (getattr(cfg, "early_dist", False) and kind in ("depthnet", "fusionnet")) else n
    return z, last


# ----------------------------------------------------------------------------
# a11 / a12  one training step   (depth_train.py:384-456 non-half branch, train.py:153-186)
# ----------------------------------------------------------------------------


def trainable_names(sd):
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


class StepOracle:
    """Reference training step on CPU: forward, head, loss, backward, clip-norm, Adam.

    Restates depth_train.py:384-456 (non-half branch :451-456): Adam(lr, weight_decay as
    L2, default betas/eps) over all parameters (the two 'bn'/'non-bn' groups of
    depth_train.py:22-25 share every hyper-parameter, so one group is equivalent),
    clip_grad_norm_(5.0).
    """

    def __init__(self, sd, kind, model, cfg, *, learn_rate=5e-5, weight_decay=4e-5, grad_norm=5.0,
                 depth_range=1000.0, loss_div=10.0, key_index=16, criterion="SmoothL1", act_round=None):
        self.sd, self.kind, self.model, self.cfg = sd, kind, model, cfg
        self.act_round = act_round          # see net_forward: round_bf16 restates the bf16 storage contract
        self.trace = None                   # set to a list to record per-block activations (net_forward `trace`)
        self.names = trainable_names(sd)
        for k in self.names:
            sd[k].requires_grad_(True)
        self.opt = torch.optim.Adam([sd[k] for k in self.names], learn_rate, weight_decay=weight_decay)
        self.grad_norm, self.depth_range, self.loss_div = grad_norm, depth_range, loss_div
        self.key_index, self.criterion = key_index, criterion
        self.side_out = (cfg.side_in - 1) // cfg.stride + 1

    def forward_loss(self, batch):
        color, depth, true_cam, true_val = batch
        if self.kind in ("fusionnet", "partial_fusionnet"):
            z, _ = net_forward(self.sd, self.kind, self.model, self.cfg, color, depth, True, self.act_round, self.trace)
        elif self.kind == "resnet":
            z = net_forward(self.sd, self.kind, self.model, self.cfg, color, None, True, self.act_round)
            z = z[0] if isinstance(z, tuple) else z
        else:
            inp = depth if (self.kind == "partial_depthnet" or self.cfg.depth_only) else color
            z, _ = net_forward(self.sd, self.kind, self.model, self.cfg, inp, None, True, self.act_round, self.trace)
        ld = 1.0 if self.kind == "resnet" else self.loss_div        # train.py:174 has no loss_div
        loss, spec = pose_loss(z, true_cam, true_val, depth=self.cfg.depth, num_joints=self.cfg.num_joints,
                               side_out=self.side_out, depth_range=self.depth_range,
                               key_index=self.key_index, loss_div=ld, criterion=self.criterion)
        return loss, spec, z

    def step(self, batch):
        loss, spec, z = self.forward_loss(batch)
        self.opt.zero_grad()
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_([self.sd[k] for k in self.names], self.grad_norm)
        self.opt.step()
        return float(loss), float(gn), spec.detach(), z.detach()


# ----------------------------------------------------------------------------
# distillation ("privileged information") step
# ----------------------------------------------------------------------------


def distill_loss(teach_last, last_feat, atten_map, sigmoid=False, bin_dist=False):
    """Trainer.distill, depth_train.py:115-129 (restated; `batch` there is last_feat.size(0))."""
    batch = last_feat.size(0)
    if bin_dist:
        diff = F.binary_cross_entropy_with_logits(last_feat, torch.sigmoid(teach_last))     # :117 (a scalar mean)
        diff = torch.mul(diff, atten_map)                                                     # :119
        return torch.sum(diff.view(batch, -1), dim=-1).mean()                                 # :121
    diff = (torch.sigmoid(teach_last) - torch.sigmoid(last_feat)) if sigmoid else (teach_last - last_feat)   # :123
    diff = torch.mul(diff, atten_map)                                                         # :125
    return torch.linalg.norm(diff.view(batch, -1), dim=-1).mean()                             # :127


def get_attention(side_in, stride, image_coords, attention=True):
    """utils.get_attention, utils.py:14-42 (numpy, float64)."""
    side_out = (side_in - 1) // stride + 1
    if not attention:
        return np.ones((side_out, side_out))[None]
    cx, cy = np.meshgrid(np.arange(side_out), np.arange(side_out))
    cx, cy = cx[..., None], cy[..., None]
    dist_x = cx - image_coords[:, 0] / (side_in / side_out)
    dist_y = cy - image_coords[:, 1] / (side_in / side_out)
    radial = np.exp(-(dist_x ** 2 + dist_y ** 2) / 5.0).sum(axis=-1)
    return (radial / np.amax(radial))[None]


def dist_weight_at(epoch, alpha_init=0.1, alpha_dest=0.1, alpha_span=10):
    """Trainer.get_dist_weight, depth_train.py:641-647."""
    alphas = np.linspace(alpha_init, alpha_dest, alpha_span)
    return float(alphas[epoch - 1]) if epoch - 1 < alpha_span else float(alpha_dest)


def synth_distill_batch(n, side, num_joints, stride=16, seed=1, invalid_frac=0.25):
    """synth_batch + the attention map the data set attaches (depth_datasets.py builds it with
    utils.get_attention from the projected joints): joints ~ U(0, side) pixels."""
    color, depth, true_cam, true_val = synth_batch(n, side, num_joints, seed, invalid_frac)
    g = torch.Generator().manual_seed(seed + 1000)
    coords = torch.rand(n, num_joints, 2, generator=g) * side
    atten = np.stack([get_attention(side, stride, coords[i].numpy().astype(np.float64)) for i in range(n)])
    return color, depth, true_cam, true_val, torch.tensor(atten, dtype=torch.float32), coords


class DistillOracle(StepOracle):
    """distill_train core, depth_train.py:179-283 (non-half branch): a fixed teacher (train-mode BN under
    no_grad unless `freeze`), the student sees the colour image, loss = dist * alpha + cam."""

    def __init__(self, sd, kind, model, cfg, teacher_sd, teacher_kind, teacher_cfg, *, sigmoid=False, bin_dist=False,
                 freeze=False, **kw):
        super().__init__(sd, kind, model, cfg, **kw)
        self.tsd, self.tkind, self.tcfg = teacher_sd, teacher_kind, teacher_cfg
        self.sigmoid, self.bin_dist, self.freeze = sigmoid, bin_dist, freeze

    def step(self, batch, alpha):
        color, depth, true_cam, true_val, atten = batch[:5]
        with torch.no_grad():
            if self.tkind in ("fusionnet", "partial_fusionnet"):
                _, teach_last = net_forward(self.tsd, self.tkind, self.model, self.tcfg, color, depth, not self.freeze)
            else:
                _, teach_last = net_forward(self.tsd, self.tkind, self.model, self.tcfg,
                                            depth if self.tcfg.depth_only else color, None, not self.freeze)
        z, last = net_forward(self.sd, self.kind, self.model, self.cfg, color, None, not self.freeze)
        dist = distill_loss(teach_last, last, atten, self.sigmoid, self.bin_dist)
        cam, spec = pose_loss(z, true_cam, true_val, depth=self.cfg.depth, num_joints=self.cfg.num_joints,
                              side_out=self.side_out, depth_range=self.depth_range, key_index=self.key_index,
                              loss_div=self.loss_div, criterion=self.criterion)
        loss = dist * alpha + cam
        self.opt.zero_grad()
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_([self.sd[k] for k in self.names], self.grad_norm)
        self.opt.step()
        return float(cam), float(dist), float(gn), spec.detach(), last.detach()


# ----------------------------------------------------------------------------
# 2-D heat-map head (mat_utils.py:31-55) and camera projection (back_project.py:12-36)
# ----------------------------------------------------------------------------


def mat_to_heatmap(ausgabe, num_joints, height, width):
    """mat_utils.to_heatmap, mat_utils.py:31-41."""
    heatmap = ausgabe.view(-1, num_joints, height * width)
    heatmap = torch.exp(heatmap - torch.max(heatmap, dim=2, keepdim=True)[0])
    heatmap = heatmap / torch.sum(heatmap, dim=2, keepdim=True)
    return heatmap.view(-1, num_joints, height, width)


def mat_decode(heatmap, map_range):
    """mat_utils.decode, mat_utils.py:44-55."""
    heat_x, heat_y = torch.sum(heatmap, dim=2), torch.sum(heatmap, dim=3)
    grid_x = torch.linspace(0.0, 1.0, heat_x.size(-1)).view(1, 1, -1)
    grid_y = torch.linspace(0.0, 1.0, heat_y.size(-1)).view(1, 1, -1)
    return torch.stack((torch.sum(grid_x * heat_x, dim=-1), torch.sum(grid_y * heat_y, dim=-1)), dim=2) * map_range


def project_points(X, cam):
    """back_project.projectPoints, back_project.py:12-36 (X: 3 x N)."""
    K, R, t, Kd = (np.asarray(cam[k], np.float64) for k in ("K", "R", "t", "distCoef"))
    x = R @ np.asarray(X, np.float64) + t.reshape(3, 1)
    x[0:2, :] = x[0:2, :] / x[2, :]
    r = x[0, :] * x[0, :] + x[1, :] * x[1, :]
    rad = 1 + Kd[0] * r + Kd[1] * r * r + Kd[4] * r * r * r
    x0 = x[0, :] * rad + 2 * Kd[2] * x[0, :] * x[1, :] + Kd[3] * (r + 2 * x[0, :] * x[0, :])
    # as written in the reference (:29-30) row 0 is overwritten before row 1 is computed, so the tangential
    # cross term of row 1 sees the already DISTORTED x -- reproduced, not corrected
    x1 = x[1, :] * rad + 2 * Kd[3] * x0 * x[1, :] + Kd[2] * (r + 2 * x[1, :] * x[1, :])
    x[0, :] = K[0, 0] * x0 + K[0, 1] * x1 + K[0, 2]
    x[1, :] = K[1, 0] * x0 + K[1, 1] * x1 + K[1, 2]
    return x


# ----------------------------------------------------------------------------
# evaluation metrics (utils.py:197-262, depth_train.py:522-527)
# ----------------------------------------------------------------------------


def statistics(basic, flip, tangent, thresh):
    """utils.statistics, utils.py:197-221: sequential elimination into the error taxonomy."""
    dist = dict(basic=basic, flip=flip, tangent=tangent)

    def count_and_eliminate(condition):
        remains = np.nonzero(np.logical_not(condition))
        for k in dist:
            dist[k] = dist[k][remains]
        return np.count_nonzero(condition)

    count = float(dist["basic"].size)
    solid = count_and_eliminate(dist["basic"] <= thresh["solid"]) / count
    close = count_and_eliminate(dist["basic"] <= thresh["close"]) / count
    depth = count_and_eliminate(dist["tangent"] <= thresh["close"]) / count
    jitter = count_and_eliminate(dist["basic"] <= thresh["rough"]) / count
    switch = count_and_eliminate(dist["flip"] <= thresh["rough"]) / count
    return dict(solid=solid, close=close, depth=depth, jitter=jitter, switch=switch, fail=dist["basic"].size / count)


def analyze(spec_cam, true_cam, valid_mask, mirror, thresh, back_rotate=None):
    """utils.analyze, utils.py:234-262, after the back-rotation einsum of depth_train.py:522-523."""
    if back_rotate is not None:
        spec_cam = np.einsum("Bij,BCj->BCi", back_rotate, spec_cam)
        true_cam = np.einsum("Bij,BCj->BCi", back_rotate, true_cam)
    valid = valid_mask.flatten()
    dist = np.linalg.norm(spec_cam - true_cam, axis=-1).flatten()[valid]
    dist_flip = np.linalg.norm(spec_cam - true_cam[:, mirror], axis=-1).flatten()[valid]
    dist_tangent = np.linalg.norm(spec_cam[:, :, :2] - true_cam[:, :, :2], axis=-1).flatten()[valid]
    stats = statistics(dist, dist_flip, dist_tangent, thresh)
    stats.update(batch_size=dist.shape[0], score_pck=np.mean(dist / thresh["rough"] <= 1.0),
                 score_auc=np.mean(np.maximum(0, 1 - dist / thresh["rough"])), cam_mean=np.mean(dist))
    return stats


def parse_epoch(stats):
    """utils.parse_epoch, utils.py:224-231."""
    keys = ("solid", "close", "jitter", "depth", "switch", "fail", "score_pck", "score_auc", "cam_mean", "batch_size")
    values = np.array([[patch[key] for patch in stats] for key in keys])
    return dict(zip(keys[:-1], np.sum(values[-1] * values[:-1], axis=1) / np.sum(values[-1])))


# ----------------------------------------------------------------------------
# input pipeline (depth_datasets.py:39-56,153-217; cameralib.py:667-711)
# ----------------------------------------------------------------------------


def homography(old_K, old_R, new_K, new_R):
    """cameralib.reproject_image_fast, cameralib.py:672-674.  The matrices keep their own dtypes (a Camera
    holds float32 members, but square_pixels() / zoom() leave a float64 intrinsic matrix behind), so the
    products round exactly as the reference's do."""
    old_matrix = np.asarray(old_K) @ np.asarray(old_R)
    new_matrix = np.asarray(new_K) @ np.asarray(new_R)
    return (old_matrix @ np.linalg.inv(new_matrix)).astype(np.float32)


def remap_bilinear(image, hom, out_shape):
    """cv2.remap(image, x, y, INTER_LINEAR, BORDER_CONSTANT 0) with the maps of cameralib.py:690-693,
    restated from OpenCV's documented behaviour (4.x imgproc/imgwarp.cpp, `remap` + `remapBilinear`):
    coordinates rounded to 1/32 px (cvRound(v * INTER_TAB_SIZE)), 2-D weights = products of {1-f/32, f/32};
    uint8: 15-bit fixed-point weights, (sum + 2^14) >> 15; float: fp32 weighted sum; neighbours outside
    the frame read the border value.  image: [H, W] or [H, W, C]; returns [Ho, Wo(, C)]."""
    Ho, Wo = out_shape
    y, x = np.mgrid[:Ho, :Wo].astype(np.float32)
    coords = np.stack([x, y, np.ones_like(x)], axis=0).reshape(3, -1)
    coords = np.asarray(hom, np.float32) @ coords
    coords = coords[:2] / coords[2:]
    sx = np.rint(np.clip(coords[0] * np.float32(32), -1e9, 1e9)).astype(np.int64)
    sy = np.rint(np.clip(coords[1] * np.float32(32), -1e9, 1e9)).astype(np.int64)
    ix, iy, fx, fy = sx >> 5, sy >> 5, sx & 31, sy & 31
    img = image if image.ndim == 3 else image[..., None]
    H, W, Cc = img.shape

    def fetch(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(ok[:, None], v, 0)

    p00, p01, p10, p11 = fetch(iy, ix), fetch(iy, ix + 1), fetch(iy + 1, ix), fetch(iy + 1, ix + 1)
    if img.dtype == np.uint8:
        w00, w01 = (32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32
        w10, w11 = (32 - fx) * fy * 32, fx * fy * 32
        acc = (p00.astype(np.int64) * w00[:, None] + p01.astype(np.int64) * w01[:, None]
               + p10.astype(np.int64) * w10[:, None] + p11.astype(np.int64) * w11[:, None] + (1 << 14)) >> 15
        out = acc.astype(np.uint8)
    else:
        ax, ay = fx.astype(np.float32) / np.float32(32), fy.astype(np.float32) / np.float32(32)
        w00, w01 = (1 - ay) * (1 - ax), (1 - ay) * ax
        w10, w11 = ay * (1 - ax), ay * ax
        out = (p00.astype(np.float32) * w00[:, None] + p01.astype(np.float32) * w01[:, None]
               + p10.astype(np.float32) * w10[:, None] + p11.astype(np.float32) * w11[:, None]).astype(np.float32)
    out = out.reshape(Ho, Wo, Cc)
    return out if image.ndim == 3 else out[..., 0]


def enhance(image, nexponent, data_name="ntu"):
    """enhance_ntu / enhance_pku, depth_datasets.py:39-56 (np.float there is float64)."""
    image = image / (10.0 / 255.0)
    veil = ((0.1 if data_name == "ntu" else 0.5) <= image).astype(np.float64)
    dest = np.multiply(np.exp(-image), veil) if nexponent else (image / 3.0)
    return dest.astype(np.float32)[np.newaxis, :, :]


def normalize_rgb(image_u8, mean=(0.485, 0.456, 0.406), dev=(0.229, 0.224, 0.225)):
    """transforms.Compose([ToTensor(), Normalize(mean, dev)]), depth_datasets.py:92-94: HWC uint8 -> CHW fp32."""
    t = torch.from_numpy(np.ascontiguousarray(image_u8)).permute(2, 0, 1).float().div(255)
    m, s = torch.tensor(mean).view(3, 1, 1), torch.tensor(dev).view(3, 1, 1)
    return (t - m) / s


def learn_rate_at(epoch, *, learn_rate=5e-5, warmup=1, warmup_factor=0.2, learn_decay=0.2):
    """depth_train.py:621-638."""
    e = epoch - 1
    if e < warmup:
        return learn_rate * warmup_factor
    if e < 15:
        return learn_rate
    if e < 20:
        return learn_rate * learn_decay
    if e < 25:
        return learn_rate * learn_decay ** 2
    return learn_rate * learn_decay ** 3


def legacy_learn_rate_at(epoch, *, learn_rate, n_epochs, do_track=False):
    """train.py:380-392."""
    e = epoch - 1
    lr = learn_rate if e < n_epochs * 0.6 else (learn_rate * 0.2 if e < n_epochs * 0.9 else learn_rate * 0.04)
    return lr / 2 if (do_track and epoch != 1) else lr


# ----------------------------------------------------------------------------
# synthetic batches -- shared by tests, smoke and bench
# ----------------------------------------------------------------------------


def blob_mask(n, side, invalid_frac, gen):
    """Validity mask [n,1,side,side] with rectangular holes (16-96 px scaled to the
    image side) until ``invalid_frac`` of the pixels are invalid."""
    m = torch.ones(n, 1, side, side)
    lo, hi = max(2, side // 16), max(3, (side * 3) // 8)
    for i in range(n):
        guard = 0
        while (1 - m[i].mean()) < invalid_frac and guard < 10000:
            h = int(torch.randint(lo, hi + 1, (1,), generator=gen))
            w = int(torch.randint(lo, hi + 1, (1,), generator=gen))
            top = int(torch.randint(0, max(1, side - h + 1), (1,), generator=gen))
            left = int(torch.randint(0, max(1, side - w + 1), (1,), generator=gen))
            m[i, :, top:top + h, left:left + w] = 0
            guard += 1
    return m


def synth_batch(n, side, num_joints, seed=1, invalid_frac=0.25, key_index=None):
    """(color [n,3,S,S], depth [n,1,S,S], true_cam [n,J,3] mm, true_val [n,J] bool)."""
    g = torch.Generator().manual_seed(seed)
    color = torch.randn(n, 3, side, side, generator=g)
    depth = (torch.rand(n, 1, side, side, generator=g) * 0.95 + 0.05) * blob_mask(n, side, invalid_frac, g)
    true_cam = torch.randn(n, num_joints, 3, generator=g) * 300.0
    true_val = torch.rand(n, num_joints, generator=g) < 0.9
    true_val[:, (num_joints - 1) if key_index is None else key_index] = True
    return color, depth, true_cam, true_val
