"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE  --  test infrastructure only.

Run in the build container (needs /root/reference, which does not exist on the
GPU box):

    python oracle/make_golden.py

It imports the reference's own modules (partial_conv, partial_depthnet,
partial_fusionnet, depthnet, fusionnet, resnet, utils, cameralib) unmodified,
with empty stand-ins for the three third-party imports that are missing in this
image (imageio, transforms3d, pyyolo -- none is touched by the hot path), feeds
them seeded inputs and stores inputs + outputs as small fixtures.  Network
weights are NOT stored: they are regenerated from ``pose_oracle.init_state(seed)``
(deterministic CPU generator) and loaded into the reference model with
``load_state_dict``; fixtures hold the seeds and the reference's outputs.

The only deviation from "unmodified": ``partial_fusionnet`` crashes as shipped
because its two stem convs are swapped (partial_fusionnet.py:202-203 vs :251,:257;
SURVEY.md note 3).  We re-assign the two attributes after construction so the
RGB stem is plain and the depth stem partial, which is what ``manual_update``
(:293) and partial_depthnet.py:177 show was intended.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, HERE)
import pose_oracle as po  # noqa: E402


def import_reference():
    for name in ("imageio", "transforms3d", "pyyolo"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    import partial_conv, partial_depthnet, partial_fusionnet, depthnet, fusionnet, resnet, utils, cameralib  # noqa
    return dict(partial_conv=partial_conv, partial_depthnet=partial_depthnet,
                partial_fusionnet=partial_fusionnet, depthnet=depthnet, fusionnet=fusionnet,
                resnet=resnet, utils=utils, cameralib=cameralib)


def np_(t):
    return t.detach().cpu().float().numpy() if torch.is_tensor(t) else np.asarray(t)


# ------------------------------------------------------------------ known answers
def gen_known_answers(ref):
    PC = ref["partial_conv"].PartialConv
    out = {}
    conv = PC(1, 1, 3, padding=1, bias=True)
    with torch.no_grad():
        conv.weight.fill_(1.0)
        conv.bias.fill_(0.5)
    x = torch.arange(1, 17, dtype=torch.float32).view(1, 1, 4, 4).requires_grad_(True)
    m = torch.tensor([[0, 0, 0, 0], [0, 0, 1, 1], [1, 1, 1, 1], [1, 1, 1, 1]], dtype=torch.float32).view(1, 1, 4, 4)
    y, mo = conv(x, m)
    y.sum().backward()
    out.update(ka1_x=np_(x), ka1_mask=np_(m), ka1_out=np_(y), ka1_mask_out=np_(mo),
               ka1_dx=np_(x.grad), ka1_dw=np_(conv.weight.grad), ka1_db=np_(conv.bias.grad))
    # KA2: fp32 renormalisation ratios, computed by the reference's own expression
    pairs = [(9, 9), (1, 1), (49, 49), (9, 6), (9, 4), (9, 1), (49, 1), (49, 25), (9, 0), (49, 0), (1, 0)]
    vals = []
    for win, cnt in pairs:
        c = torch.tensor([float(cnt)])
        r = win / (c + 1e-6)
        vals.append(float(r * torch.clamp(c, 0, 1)))
    out.update(ka2_pairs=np.array(pairs, np.int32), ka2_ratio=np.array(vals, np.float32))
    # KA3: all-invalid window with bias -> exactly 0
    x3 = torch.randn(1, 1, 4, 4)
    y3, mo3 = conv(x3, torch.zeros(1, 1, 4, 4))
    out.update(ka3_out=np_(y3), ka3_mask_out=np_(mo3))
    return out


# ------------------------------------------------------------------ PartialConv layer cases
PCONV_CASES = [
    # name,        N, C,  K,  H,  W, k, s, p, d, bias
    ("k1",         2, 8,  16, 9,  9, 1, 1, 0, 1, False),
    ("k3",         2, 8,  16, 9,  11, 3, 1, 1, 1, False),
    ("k3s2",       2, 16, 8,  13, 13, 3, 2, 1, 1, False),
    ("k3d2",       2, 8,  8,  12, 12, 3, 1, 2, 2, False),
    ("k7s2_c1",    2, 1,  16, 17, 17, 7, 2, 3, 1, False),
    ("k3_bias",    2, 8,  8,  8,  8, 3, 1, 1, 1, True),
    ("k1s2",       2, 8,  8,  9,  9, 1, 2, 0, 1, False),
    ("k3_c64",     1, 64, 64, 10, 10, 3, 1, 1, 1, False),
    ("k1_c64_128", 1, 64, 128, 8, 8, 1, 1, 0, 1, False),
]


def gen_pconv_cases(ref):
    PC = ref["partial_conv"].PartialConv
    out = {}
    g = torch.Generator().manual_seed(1234)
    for name, N, C, K, H, W, k, s, p, d, has_bias in PCONV_CASES:
        conv = PC(C, K, kernel_size=k, stride=s, padding=p, dilation=d, bias=has_bias)
        w = torch.randn(K, C, k, k, generator=g) * (2.0 / (k * k * K)) ** 0.5
        with torch.no_grad():
            conv.weight.copy_(w)
            if has_bias:
                conv.bias.copy_(torch.randn(K, generator=g) * 0.1)
        x = torch.randn(N, C, H, W, generator=g).requires_grad_(True)
        m = po.blob_mask(N, max(H, W), 0.35, g)[:, :, :H, :W].contiguous()
        y, mo = conv(x, m)
        cot = torch.randn(y.shape, generator=g)
        (y * cot).sum().backward()
        out.update({f"{name}_x": np_(x), f"{name}_mask": np_(m), f"{name}_w": np_(conv.weight),
                    f"{name}_cot": np_(cot), f"{name}_out": np_(y), f"{name}_mask_out": np_(mo),
                    f"{name}_dx": np_(x.grad), f"{name}_dw": np_(conv.weight.grad)})
        if has_bias:
            out.update({f"{name}_b": np_(conv.bias), f"{name}_db": np_(conv.bias.grad)})
        # KA4 dtype behaviour: bf16 input, fp32 mask
        yb, mob = conv.bfloat16()(x.detach().bfloat16(), m)
        # no-bias path keeps the input dtype; the bias path multiplies by the fp32 mask_out last
        # (partial_conv.py:51) so type promotion makes its output fp32.  mask_out is always fp32.
        assert mob.dtype == torch.float32
        assert yb.dtype == (torch.float32 if has_bias else torch.bfloat16)
        out[f"{name}_out_bf16"] = np_(yb)
    out["names"] = np.array([c[0] for c in PCONV_CASES])
    out["specs"] = np.array([c[1:] for c in PCONV_CASES], np.int32)
    return out


# ------------------------------------------------------------------ head
HEAD_CASES = [("j17_16", 2, 17, 16, 16, 16), ("j19_17", 1, 19, 16, 17, 17), ("j25_17", 1, 25, 16, 17, 17),
              ("j5_d8_hw7x9", 2, 5, 8, 7, 9)]


def gen_head(ref):
    U = ref["utils"]
    out = {}
    g = torch.Generator().manual_seed(77)
    for name, N, J, D, H, W in HEAD_CASES:
        feat = (torch.randn(N, D * J, H, W, generator=g) * 3).requires_grad_(True)
        heat = U.to_heatmap(feat, D, J, H, W)
        coords = U.decode(heat, 1000.0)
        cot = torch.randn(coords.shape, generator=g)
        (coords * cot).sum().backward()
        out.update({f"{name}_feat": np_(feat), f"{name}_coords": np_(coords), f"{name}_cot": np_(cot),
                    f"{name}_dfeat": np_(feat.grad), f"{name}_heat_sum": np_(heat.sum(dim=(2, 3, 4)))})
        if name == "j5_d8_hw7x9":
            out[f"{name}_heat"] = np_(heat)
    # KA6
    uni = U.decode(U.to_heatmap(torch.zeros(1, 16 * 3, 5, 6), 16, 3, 5, 6), 1000.0)
    hot = torch.full((1, 16 * 2, 5, 6), -1e4)
    hot[0, 7 * 2 + 1, 3, 4] = 50.0      # joint 1 at (d=7,h=3,w=4)
    hot[0, 2 * 2 + 0, 0, 5] = 50.0      # joint 0 at (d=2,h=0,w=5)
    one = U.decode(U.to_heatmap(hot, 16, 2, 5, 6), 1000.0)
    out.update(ka6_uniform=np_(uni), ka6_onehot=np_(one))
    out["names"] = np.array([c[0] for c in HEAD_CASES])
    out["specs"] = np.array([c[1:] for c in HEAD_CASES], np.int32)
    return out


# ------------------------------------------------------------------ to_depth
def gen_to_depth(ref):
    U, C = ref["utils"], ref["cameralib"]
    out = {}
    rng = np.random.RandomState(5)
    for name, H, W, fx, fy, cx, cy in [("sq64", 64, 64, 365.0, 365.0, 32.0, 32.0),
                                       ("r48x80", 48, 80, 366.1, 364.2, 39.3, 25.7)]:
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
        img = (rng.rand(H, W).astype(np.float32) * 4000).astype(np.float32)
        img[rng.rand(H, W) < 0.2] = 0
        cam = C.Camera(intrinsic_matrix=K)
        res = U.to_depth(img, cam)
        out.update({f"{name}_img": img, f"{name}_K": K, f"{name}_out": np.asarray(res, np.float64)})
    out["names"] = np.array(["sq64", "r48x80"])
    return out


# ------------------------------------------------------------------ networks + training step
NET_CASES = [
    # tag,                 kind,                model,      side, N, J,  extra cfg
    ("pdepth18_s65",       "partial_depthnet",  "resnet18", 65,  2, 17, {}),
    ("pdepth50_s64",       "partial_depthnet",  "resnet50", 64,  2, 17, {}),
    ("pfusion50_s64",      "partial_fusionnet", "resnet50", 64,  2, 17, {}),
    ("pfusion18_s49_j25",  "partial_fusionnet", "resnet18", 49,  2, 25, {}),
    ("fusion50_s64",       "fusionnet",         "resnet50", 64,  2, 17, {}),
    ("fusion18_skip",      "fusionnet",         "resnet18", 64,  2, 17, dict(skip_relu=True, early_dist=True)),
    ("depth50_rgb_s64",    "depthnet",          "resnet50", 64,  2, 19, dict(depth_only=False)),
    ("depth18_d_s33",      "depthnet",          "resnet18", 33,  3, 17, {}),
    ("legacy50_s64",       "resnet",            "resnet50", 64,  2, 19, {}),
    ("pdepth50_stride8",   "partial_depthnet",  "resnet50", 64,  2, 17, dict(stride=8)),
]


def build_reference_net(ref, kind, model, cfg):
    mod = ref[kind]
    if kind == "resnet":
        net = getattr(mod, model)(cfg)
    else:
        net = getattr(mod, model)(cfg, False)
    if kind == "partial_fusionnet":       # documented stem fix (see module docstring)
        PC = ref["partial_conv"].PartialConv
        net.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        net.conv2 = PC(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
    return net


def reference_step(ref, net, kind, cfg, batch, opt, key_index, loss_div, depth_range=1000.0, grad_norm=5.0):
    """depth_train.py:384-456 (non-half branch) / train.py:153-186 around the imported model."""
    U = ref["utils"]
    color, depth, true_cam, true_val = batch
    side_out = (cfg.side_in - 1) // cfg.stride + 1
    if kind in ("fusionnet", "partial_fusionnet"):
        cam_feat, last = net(color, depth)
    elif kind == "resnet":
        cam_feat, last = net(color), None
    else:
        cam_feat, last = net(depth if (kind == "partial_depthnet" or cfg.depth_only) else color)
    heat = U.to_heatmap(cam_feat, cfg.depth, cfg.num_joints, side_out, side_out)
    rel = U.decode(heat, depth_range)
    rel = rel - rel[:, key_index:key_index + 1]
    spec = rel + true_cam[:, key_index:key_index + 1]
    crit = nn.SmoothL1Loss(reduction="mean")
    sel = true_val.view(-1)
    loss = crit(spec.view(-1, 3)[sel] / loss_div, true_cam.view(-1, 3)[sel] / loss_div)
    opt.zero_grad()
    loss.backward()
    gn = nn.utils.clip_grad_norm_(list(net.parameters()), grad_norm)
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
    opt.step()
    return loss, gn, spec, cam_feat, last, grads


def gen_nets(ref):
    out = {}
    tags = []
    for tag, kind, model, side, N, J, extra in NET_CASES:
        cfg = po.net_config(side_in=side, num_joints=J, **extra)
        torch.manual_seed(0)
        net = build_reference_net(ref, kind, model, cfg)
        sd = po.init_state(kind, model, cfg, seed=11)
        ref_keys = list(net.state_dict().keys())
        assert ref_keys == list(sd.keys()), (tag, set(ref_keys) ^ set(sd.keys()))
        for k, v in net.state_dict().items():
            assert tuple(v.shape) == tuple(sd[k].shape), (tag, k)
        net.load_state_dict(sd)
        net.train()
        key_index = J - 1
        loss_div = 1.0 if kind == "resnet" else 10.0
        opt = torch.optim.Adam([p for _, p in net.named_parameters()], 5e-5, weight_decay=4e-5)
        batch = po.synth_batch(N, side, J, seed=3, invalid_frac=0.25)
        # forward only in eval mode first (running stats at init) -> pins the inference path
        net.eval()
        with torch.no_grad():
            if kind in ("fusionnet", "partial_fusionnet"):
                z_eval = net(batch[0], batch[1])[0]
            elif kind == "resnet":
                z_eval = net(batch[0])
            else:
                z_eval = net(batch[1] if (kind == "partial_depthnet" or cfg.depth_only) else batch[0])[0]
        net.train()
        losses, gns = [], []
        for it in range(2):
            loss, gn, spec, z, last, grads = reference_step(ref, net, kind, cfg, batch, opt, key_index, loss_div)
            losses.append(float(loss))
            gns.append(float(gn))
            if it == 0:
                out[f"{tag}_z"] = np_(z)
                out[f"{tag}_spec"] = np_(spec)
                if last is not None:
                    out[f"{tag}_last_mean"] = np.array([float(last.mean()), float(last.abs().mean())], np.float64)
                    out[f"{tag}_last_slice"] = np_(last[:, :8])
                probe = ["conv1.weight", "layer1.0.conv1.weight", "layer2.0.downsample.0.weight",
                         "layer4.0.conv2.weight", "bn1.weight", "bn1.bias", "layer3.1.bn2.weight"]
                probe += [k for k in ("conv2.weight", "fusion.conv.weight", "layer6.1.conv2.weight",
                                      "layer5.0.bn1.bias", "regressor.bias", "cam_regressor.bias") if k in grads]
                for k in probe:
                    gk = grads[k]
                    out[f"{tag}_gnorm_{k}"] = np.array(float(gk.norm()), np.float64)
                    out[f"{tag}_gslice_{k}"] = np_(gk.reshape(-1)[:64])
        new_sd = net.state_dict()
        out[f"{tag}_loss"] = np.array(losses, np.float64)
        out[f"{tag}_gradnorm"] = np.array(gns, np.float64)
        out[f"{tag}_z_eval"] = np_(z_eval)
        out[f"{tag}_bn1_running_mean"] = np_(new_sd["bn1.running_mean"])
        out[f"{tag}_bn1_running_var"] = np_(new_sd["bn1.running_var"])
        out[f"{tag}_conv1_after"] = np_(new_sd["conv1.weight"].reshape(-1)[:64])
        out[f"{tag}_nparams"] = np.array(sum(p.numel() for p in net.parameters()), np.int64)
        tags.append(tag)
        print(tag, "loss", losses, "gn", gns, flush=True)
    out["tags"] = np.array(tags)
    return out


# ------------------------------------------------------------------ distillation step
DISTILL_CASES = [
    # tag, teacher kind, teacher extra, student extra, distill kwargs
    ("dist_pf18_l2",      "partial_fusionnet", {}, dict(depth_only=False), dict()),
    ("dist_f18skip_sig",  "fusionnet", dict(skip_relu=True, early_dist=True),
     dict(depth_only=False, skip_relu=True, early_dist=True), dict(sigmoid=True)),
    ("dist_pf18_bce_frz", "partial_fusionnet", {}, dict(depth_only=False), dict(bin_dist=True, freeze=True)),
]


def gen_distill(ref):
    import depth_train                      # noqa: imports cleanly; only Trainer.__init__ needs the private paths
    DT = depth_train.Trainer
    U = ref["utils"]
    out = {}
    # -- the loss alone (Trainer.distill called unbound on a namespace carrying the two switches)
    g = torch.Generator().manual_seed(5)
    names = []
    for name, (N, C, H, W) in (("small", (3, 16, 5, 5)), ("odd", (2, 6, 3, 4))):
        t = torch.randn(N, C, H, W, generator=g) * 2
        a = torch.rand(N, 1, H, W, generator=g)
        for mode, kw in (("l2", dict(sigmoid=False, bin_dist=False)), ("sigmoid", dict(sigmoid=True, bin_dist=False)),
                         ("bce", dict(sigmoid=False, bin_dist=True))):
            sfeat = (torch.randn(N, C, H, W, generator=g) * 2).requires_grad_(True)
            loss = DT.distill(types.SimpleNamespace(**kw), N, t, sfeat, a)
            loss.backward()
            key = f"loss_{name}_{mode}"
            out.update({key + "_t": np_(t), key + "_s": np_(sfeat), key + "_a": np_(a),
                        key + "_loss": np.array(float(loss), np.float64), key + "_ds": np_(sfeat.grad)})
            assert abs(float(po.distill_loss(t, sfeat.detach(), a, **kw)) - float(loss)) < 1e-6 * max(1, abs(float(loss)))
            names.append(key)
    out["loss_names"] = np.array(names)
    # -- attention maps
    for name, (side, stride, J) in (("att_257", (257, 16, 17)), ("att_64", (64, 16, 5)), ("att_48s8", (48, 8, 3))):
        coords = (torch.rand(J, 2, generator=g) * side).numpy().astype(np.float64)
        out[name + "_coords"] = coords
        out[name + "_cfg"] = np.array([side, stride], np.int32)
        out[name + "_map"] = U.get_attention(side, stride, coords, True)
        out[name + "_ones"] = U.get_attention(side, stride, coords, False)
    # -- schedule
    sched = types.SimpleNamespace(alpha_init=0.5, alpha_dest=0.1, alpha_span=5)
    out["alpha_sched"] = np.array([DT.get_dist_weight(sched, e) for e in range(1, 9)], np.float64)
    # -- whole steps: distill_train core (depth_train.py:179-283, non-half branch) around the imported nets
    tags = []
    for tag, tkind, textra, sextra, dkw in DISTILL_CASES:
        side, N, J, model = 64, 2, 17, "resnet18"
        tcfg = po.net_config(side_in=side, num_joints=J, **textra)
        scfg = po.net_config(side_in=side, num_joints=J, **sextra)
        teacher = build_reference_net(ref, tkind, model, tcfg)
        teacher.load_state_dict(po.init_state(tkind, model, tcfg, seed=21))
        student = build_reference_net(ref, "depthnet", model, scfg)
        student.load_state_dict(po.init_state("depthnet", model, scfg, seed=11))
        teacher.train()
        student.train()
        freeze = dkw.get("freeze", False)
        if freeze:                                            # Trainer.freeze_batchnorm, depth_train.py:156-158
            teacher.eval()
            student.freeze_batchnorm()
        ns = types.SimpleNamespace(sigmoid=dkw.get("sigmoid", False), bin_dist=dkw.get("bin_dist", False))
        opt = torch.optim.Adam(list(student.parameters()), 5e-5, weight_decay=4e-5)
        batch = po.synth_distill_batch(N, side, J, stride=16, seed=3)
        color, depth, true_cam, true_val, atten, _ = batch
        alpha, key_index, side_out = 0.3, J - 1, (side - 1) // 16 + 1
        cams, dists, gns = [], [], []
        for it in range(2):
            with torch.no_grad():
                _, teach_last = teacher(color, depth)
            cam_feat, last_feat = student(color)
            dist_loss = DT.distill(ns, N, teach_last, last_feat, atten)
            heat = U.to_heatmap(cam_feat, scfg.depth, J, side_out, side_out)
            rel = U.decode(heat, 1000.0)
            rel = rel - rel[:, key_index:key_index + 1]
            spec = rel + true_cam[:, key_index:key_index + 1]
            sel = true_val.view(-1)
            cam_loss = nn.SmoothL1Loss(reduction="mean")(spec.view(-1, 3)[sel] / 10.0, true_cam.view(-1, 3)[sel] / 10.0)
            loss = dist_loss * alpha + cam_loss
            opt.zero_grad()
            loss.backward()
            gn = nn.utils.clip_grad_norm_(list(student.parameters()), 5.0)
            if it == 0:
                out[f"{tag}_spec"] = np_(spec)
                out[f"{tag}_last_slice"] = np_(last_feat[:, :8])
                out[f"{tag}_teach_slice"] = np_(teach_last[:, :8])
                out[f"{tag}_g_l4"] = np_(dict(student.named_parameters())["layer4.1.conv2.weight"].grad.reshape(-1)[:64])
                out[f"{tag}_gn_l1"] = np.array(float(dict(student.named_parameters())["layer1.0.conv1.weight"].grad.norm()))
            opt.step()
            cams.append(float(cam_loss)); dists.append(float(dist_loss)); gns.append(float(gn))
        out[f"{tag}_cam"] = np.array(cams, np.float64)
        out[f"{tag}_dist"] = np.array(dists, np.float64)
        out[f"{tag}_gn"] = np.array(gns, np.float64)
        out[f"{tag}_teacher_bn1_rm"] = np_(teacher.state_dict()["bn1.running_mean"])
        tags.append(tag)
        print(tag, "cam", cams, "dist", dists, "gn", gns, flush=True)
    out["tags"] = np.array(tags)
    return out


# ------------------------------------------------------------------ 2-D head + projection
def gen_head2d(ref):
    import mat_utils                                   # imports cleanly (torch, numpy, cv2)
    (projectPoints,) = _reference_functions(os.path.join(REF, "back_project.py"), ["projectPoints"])
    g = torch.Generator().manual_seed(9)
    out = {}
    for name, (N, J, H, W) in (("sq", (2, 19, 17, 17)), ("rect", (3, 5, 9, 13))):
        feat = (torch.randn(N, J, H, W, generator=g) * 3).requires_grad_(True)
        heat = mat_utils.to_heatmap(feat, J, H, W)
        coords = mat_utils.decode(heat, 257.0)
        cot = torch.randn(N, J, 2, generator=g)
        (coords * cot).sum().backward()
        out.update({f"{name}_feat": np_(feat), f"{name}_heat": np_(heat), f"{name}_coords": np_(coords),
                    f"{name}_cot": np_(cot), f"{name}_dfeat": np_(feat.grad)})
    rng = np.random.RandomState(2)
    q, _ = np.linalg.qr(rng.randn(3, 3))
    cam = dict(K=np.matrix([[1400.0, 0.5, 950.0], [0, 1395.0, 540.0], [0, 0, 1]]), R=np.matrix(q),
               t=np.matrix(rng.randn(3, 1) * 10 + np.array([[0], [0], [300.0]])),
               distCoef=np.array([-0.28, 0.11, 1e-3, -5e-4, -0.02]))
    X = np.matrix(rng.randn(3, 40) * 60.0)
    out.update(proj_X=np.asarray(X), proj_K=np.asarray(cam["K"]), proj_R=np.asarray(cam["R"]), proj_t=np.asarray(cam["t"]),
               proj_Kd=cam["distCoef"], proj_out=np.asarray(projectPoints(X, cam)))
    return out


# ------------------------------------------------------------------ evaluation metrics
def gen_metrics(ref):
    U = ref["utils"]
    rng = np.random.RandomState(3)
    out = {}
    thresh = dict(solid=50.0, close=100.0, rough=150.0)
    mirror = np.array([0, 2, 1, 4, 3, 5, 7, 6, 8, 10, 9, 12, 11, 14, 13, 16, 15])
    batches = []
    for b, N in enumerate((5, 8, 3)):
        true = (rng.randn(N, 17, 3) * 300).astype(np.float32)
        spec = true + (rng.randn(N, 17, 3) * rng.choice([10.0, 60.0, 200.0], (N, 17, 1))).astype(np.float32)
        swap = rng.rand(N, 17) < 0.1                       # some left/right switches
        spec[swap] = true[:, mirror][swap] + (rng.randn(int(swap.sum()), 3) * 20).astype(np.float32)
        deep = rng.rand(N, 17) < 0.1                       # some pure depth errors
        spec[deep, 2] += 400.0
        valid = rng.rand(N, 17) < 0.85
        q, _ = np.linalg.qr(rng.randn(N, 3, 3))
        rot = q.astype(np.float32)
        # depth_train.py:522-527 executed with the reference's utils.analyze
        s = np.einsum("Bij,BCj->BCi", rot, spec)
        t = np.einsum("Bij,BCj->BCi", rot, true)
        stats = U.analyze(s, t, valid, mirror, thresh)
        batches.append(stats)
        out.update({f"b{b}_spec": spec, f"b{b}_true": true, f"b{b}_valid": valid, f"b{b}_rot": rot})
        for k, v in stats.items():
            out[f"b{b}_{k}"] = np.array(v, np.float64)
    epoch = U.parse_epoch(batches)
    for k, v in epoch.items():
        out[f"epoch_{k}"] = np.array(v, np.float64)
    out["mirror"] = mirror
    out["thresh"] = np.array([thresh["solid"], thresh["close"], thresh["rough"]], np.float64)
    out["n_batches"] = np.array(3)
    return out


# ------------------------------------------------------------------ input pipeline
def _reference_functions(path, names):
    """Compile selected top-level functions straight from a reference source file (the module itself does not
    import here: jpeg4py, matplotlib, pickle5 ... are absent) -- the reference's code runs, nothing is copied."""
    import ast
    import re
    text = open(path).read()
    ns = {"np": np}
    try:
        tree = ast.parse(text)
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                exec(compile(ast.Module([node], []), path, "exec"), ns)
    except SyntaxError:                 # a Python-2 file (back_project.py): run just the text of the wanted functions
        for n in names:
            m = re.search(r"^def %s\(.*?(?=^\S)" % n, text, flags=re.S | re.M)
            exec(compile(m.group(0), path, "exec"), ns)
    return [ns[n] for n in names]


def gen_pipeline(ref):
    import copy
    import torchvision.transforms as transforms
    CL, U = ref["cameralib"], ref["utils"]
    if not hasattr(np, "float"):
        np.float = float                      # the reference predates numpy 1.24 (depth_datasets.py:42)
    enhance_ntu, enhance_pku = _reference_functions(os.path.join(REF, "depth_datasets.py"), ["enhance_ntu", "enhance_pku"])
    transform = transforms.Compose([transforms.ToTensor(),
                                    transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    rng = np.random.RandomState(7)
    out, names = {}, []
    side = 48
    for name, (Hs, Ws, f, bbox, flip, zoom) in dict(
            centre=(120, 160, 130.0, (50, 20, 40, 80), False, 1.0),
            corner_flip=(120, 160, 130.0, (100, 50, 50, 60), True, 0.9),       # crop leaves the frame: border pixels
            wide=(90, 128, 95.0, (10, 30, 90, 40), False, 1.15)).items():
        K = np.array([[f, 0, Ws / 2], [0, f * 1.02, Hs / 2], [0, 0, 1]], np.float32)
        camera = CL.Camera(intrinsic_matrix=K, world_up=(0, -1, 0))
        bbox = np.array(bbox, np.float64)
        # get_input_image, depth_datasets.py:153-198, executed on the reference's Camera
        center = bbox[:2] + bbox[2:] / 2
        width, height = np.array([bbox[2] / 2, 0]), np.array([0, bbox[3] / 2])
        far_side = np.stack([center - height, center + height]) if bbox[2] < bbox[3] else np.stack([center - width, center + width])
        new_cam = copy.deepcopy(camera)
        new_cam.turn_towards(center)
        new_cam.undistort()
        new_cam.square_pixels()
        far = new_cam.world_to_image(camera.image_to_world(far_side))
        new_cam.zoom(side / np.linalg.norm(far[0] - far[1]))
        new_cam.center_principal_point((side, side))
        new_cam.zoom(zoom)
        if flip:
            new_cam.horizontal_flip()
        yy, xx = np.mgrid[:Hs, :Ws]
        color = np.clip(np.stack([xx * 255.0 / Ws, yy * 255.0 / Hs, (xx + yy) % 256], -1) + rng.randint(-40, 40, (Hs, Ws, 3)), 0, 255).astype(np.uint8)
        depth = (rng.rand(Hs, Ws).astype(np.float32) * 0.08 + 0.002) * (rng.rand(Hs, Ws) > 0.2)
        depth = depth.astype(np.float32)
        col_crop = CL.reproject_image(color, camera, new_cam, (side, side))
        dep_crop = CL.reproject_image(depth, camera, new_cam, (side, side))
        hom = po.homography(camera.intrinsic_matrix, camera.R, new_cam.intrinsic_matrix, new_cam.R)
        assert np.array_equal(po.remap_bilinear(color, hom, (side, side)), col_crop), name
        out.update({f"{name}_color": color, f"{name}_depth": depth, f"{name}_hom": hom, f"{name}_K": K,
                    f"{name}_K_old": camera.intrinsic_matrix, f"{name}_R_old": camera.R,
                    f"{name}_K_new": new_cam.intrinsic_matrix, f"{name}_R_new": new_cam.R,
                    f"{name}_color_crop": col_crop, f"{name}_depth_crop": dep_crop[..., 0],
                    f"{name}_color_out": np_(transform(col_crop.copy()))})
        d = dep_crop.squeeze()
        out[f"{name}_ntu_exp"] = enhance_ntu(d, True)
        out[f"{name}_ntu_lin"] = enhance_ntu(d, False)
        out[f"{name}_pku_exp"] = enhance_pku(d, True)
        out[f"{name}_todepth_ntu_exp"] = enhance_ntu(U.to_depth(d, camera), True)      # depth_datasets.py:214-217
        names.append(name)
    out["names"] = np.array(names)
    return out


def gen_shapes(ref):
    """KA7: parameter counts and output shapes of the full-size nets (no forward needed for counts)."""
    out = {}
    for kind, J in (("partial_depthnet", 17), ("partial_fusionnet", 17), ("fusionnet", 17)):
        cfg = po.net_config(side_in=256, num_joints=J)
        net = build_reference_net(ref, kind, "resnet50", cfg)
        out[f"{kind}_nparams"] = np.array(sum(p.numel() for p in net.parameters()), np.int64)
        out[f"{kind}_keys"] = np.array(list(net.state_dict().keys()))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = import_reference()
    which = sys.argv[1:] or ["ka", "pconv", "head", "to_depth", "shapes", "nets", "distill", "pipeline", "metrics", "head2d"]
    table = dict(ka=gen_known_answers, pconv=gen_pconv_cases, head=gen_head, to_depth=gen_to_depth,
                 shapes=gen_shapes, nets=gen_nets, distill=gen_distill, pipeline=gen_pipeline, metrics=gen_metrics, head2d=gen_head2d)
    for name in which:
        data = table[name](ref)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB", flush=True)


if __name__ == "__main__":
    main()
