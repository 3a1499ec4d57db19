"""Reference arm of bench.py: the reference's OWN modules (imported unmodified from ``baseline/_ref``, staged by
``baseline/stage_reference.py``) driven through the training step of depth_train.py:384-456.

``depth_train.Trainer`` itself cannot be constructed outside the author's cluster (it opens
/globalwork/liu/metadata.json, depth_train.py:58, and needs the private datasets), so the step loop around the imported
model + ``utils.to_heatmap`` / ``utils.decode`` is restated here in ~25 lines -- the same restatement
``oracle/make_golden.py`` used to produce the fixtures.  Benchmark harness only; never imported by the product package.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    return os.path.exists(os.path.join(REF_DIR, "partial_fusionnet.py"))


def import_reference():
    """The reference modules on the hot path; the three third-party imports this image lacks (none is touched by
    the path) get empty stand-ins."""
    for name in ("imageio", "transforms3d", "pyyolo"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    return {n: importlib.import_module(n) for n in
            ("partial_conv", "partial_depthnet", "partial_fusionnet", "depthnet", "fusionnet", "utils")}


def build_net(ref, kind, model, cfg):
    net = getattr(ref[kind], model)(cfg, False)
    if kind == "partial_fusionnet":
        # documented 2-line fix: the file ships with its two stems swapped (partial_fusionnet.py:202-203 vs :251,:257)
        # and raises TypeError as is; RGB stem plain, depth stem partial is what manual_update (:293) intends
        net.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        net.conv2 = ref["partial_conv"].PartialConv(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
    return net


class RefStep:
    """One optimisation step = depth_train.py:384-456, non-half branch (Adam lr 5e-5, wd 4e-5, clip-norm 5)."""

    def __init__(self, ref, kind, model, cfg, device, key_index, autocast=None, channels_last=False, loss_div=10.0,
                 depth_range=1000.0):
        self.ref, self.kind, self.cfg, self.device = ref, kind, cfg, device
        self.net = build_net(ref, kind, model, cfg).to(device).train()
        if channels_last:
            self.net = self.net.to(memory_format=torch.channels_last)
        self.opt = torch.optim.Adam(self.net.parameters(), 5e-5, weight_decay=4e-5)
        self.crit = nn.SmoothL1Loss(reduction="mean")
        self.key_index, self.loss_div, self.depth_range = key_index, loss_div, depth_range
        self.autocast, self.channels_last = autocast, channels_last
        self.side_out = (cfg.side_in - 1) // cfg.stride + 1

    def step(self, batch):
        U, cfg = self.ref["utils"], self.cfg
        color, depth, true_cam, true_val = batch
        if self.channels_last:
            color = color.contiguous(memory_format=torch.channels_last)
        ctx = torch.autocast(self.device.type, dtype=self.autocast) if self.autocast is not None else _Null()
        with ctx:
            if self.kind in ("fusionnet", "partial_fusionnet"):
                cam_feat, _ = self.net(color, depth)
            else:
                cam_feat, _ = self.net(depth if (self.kind == "partial_depthnet" or cfg.depth_only) else color)
        heat = U.to_heatmap(cam_feat.float(), cfg.depth, cfg.num_joints, self.side_out, self.side_out)
        rel = U.decode(heat, self.depth_range)
        k = self.key_index
        rel = rel - rel[:, k:k + 1]
        spec = rel + true_cam[:, k:k + 1]
        sel = true_val.view(-1)
        loss = self.crit(spec.view(-1, 3)[sel] / self.loss_div, true_cam.view(-1, 3)[sel] / self.loss_div)
        self.opt.zero_grad()
        loss.backward()
        nn.utils.clip_grad_norm_(list(self.net.parameters()), 5.0)
        self.opt.step()
        return loss


class _Null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False
