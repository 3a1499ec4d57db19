"""The five reference networks (depthnet / partial_depthnet / fusionnet / partial_fusionnet /
resnet) as ONE table-driven module on the libb2pose kernels.

Module and parameter names, registration order and ``[K, C, R, S]`` shapes follow the reference
(partial_depthnet.py:160-229, partial_fusionnet.py:184-274, fusionnet.py:143-240,
depthnet.py:119-200, resnet.py:122-210) so ``state_dict()`` interchanges with it.  Inside,
activations are NHWC, every conv+BN(+ReLU)(+residual) is a single fused autograd node, the veil
is threaded next to the activations and applied in the BN epilogue (so the next PartialConv reads
pre-masked input), and the outputs are returned as logical-NCHW views.

``partial_fusionnet`` is built with the *intended* stems (RGB plain 3->64, depth PartialConv
1->64): the reference file has them swapped and raises TypeError (SURVEY.md note 3).
"""
import math

import torch
from torch import nn

from . import ops
from .layers import BatchNorm2d, Conv2d, PartialConv, conv_bn


def stage_strides(net_stride):
    """Per-stage stride and dilation for a given network stride (partial_depthnet.py:169-175)."""
    lg = math.log2(net_stride)
    s2 = int(min(max(lg, 2), 3) - 1)
    s3 = int(min(max(lg, 3), 4) - 2)
    s4 = int(min(max(lg, 4), 5) - 3)
    d2 = 3 - s2
    d3 = d2 * (3 - s3)
    d4 = d3 * (3 - s4)
    return (1, s2, s3, s4), (1, d2, d3, d4)


class _Block(nn.Module):
    def __init__(self, partial, skip_relu):
        super().__init__()
        self.partial = partial
        self.skip_relu = skip_relu

    def _holder(self, x):
        """Shared dict that lets the block's first conv+BN node absorb the shortcut's gradient with a TMA reduce-add
        instead of a separate add kernel (see ops.ConvBNFn): the identity shortcut's gradient is parked by the block's
        LAST node, a down-sampling shortcut's by the shortcut node itself."""
        if not (torch.is_grad_enabled() and x.requires_grad):
            return None
        return {} if self.downsample is None else {"open": True}

    def _residual(self, x, holder=None):
        """Shortcut branch, called right AFTER the block's first conv+BN node was created (autograd runs later-created
        nodes first, so the shortcut's backward precedes the first node's and can hand it its dx through `holder`).
        A down-sampling shortcut (1x1 conv + BN) runs on its own stream beside the block's main path (forward here,
        backward through autograd's stream affinity); `_join` is called before it is consumed."""
        self._side = None
        if self.downsample is None:
            return x
        if ops.TWO_STREAMS and x.is_cuda:
            cur, side = torch.cuda.current_stream(x.device), ops.shortcut_stream(x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                res, _ = conv_bn(x, None, self.downsample[0], self.downsample[1], relu=False, park_holder=holder)
            self._side = (cur, side)
            return res
        res, _ = conv_bn(x, None, self.downsample[0], self.downsample[1], relu=False, park_holder=holder)
        return res

    def _join(self, res):
        if self._side is not None:
            cur, side = self._side
            cur.wait_stream(side)
            res.record_stream(cur)
            self._side = None
        return res


class BasicBlock(_Block):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, partial=False, skip_relu=False):
        super().__init__(partial, skip_relu)
        Conv = PartialConv if partial else Conv2d
        self.conv1 = Conv(in_channels=inplanes, out_channels=planes, kernel_size=3, stride=stride,
                          dilation=dilation, padding=dilation, bias=False)
        self.bn1 = BatchNorm2d(planes)
        self.conv2 = Conv(in_channels=planes, out_channels=planes, kernel_size=3, padding=1, bias=False)
        self.bn2 = BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward_nhwc(self, x, veil):
        h = self._holder(x)
        out, veil = conv_bn(x, veil, self.conv1, self.bn1, relu=True, mask_output=True, dx_holder=h)
        res = self._residual(x, h)
        out, veil = conv_bn(out, veil, self.conv2, self.bn2, relu=not self.skip_relu, residual=self._join(res),
                            premasked=True, res_holder=h if self.downsample is None else None)
        return out, veil


class Bottleneck(_Block):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, partial=False, skip_relu=False):
        super().__init__(partial, skip_relu)
        Conv = PartialConv if partial else Conv2d
        self.conv1 = Conv(in_channels=inplanes, out_channels=planes, kernel_size=1, bias=False)
        self.bn1 = BatchNorm2d(planes)
        self.conv2 = Conv(in_channels=planes, out_channels=planes, kernel_size=3, stride=stride,
                          padding=dilation, dilation=dilation, bias=False)
        self.bn2 = BatchNorm2d(planes)
        self.conv3 = Conv(in_channels=planes, out_channels=planes * 4, kernel_size=1, bias=False)
        self.bn3 = BatchNorm2d(planes * 4)
        self.downsample = downsample
        self.stride = stride

    def forward_nhwc(self, x, veil):
        h = self._holder(x)
        out, veil = conv_bn(x, veil, self.conv1, self.bn1, relu=True, mask_output=True, dx_holder=h)
        res = self._residual(x, h)
        out, veil = conv_bn(out, veil, self.conv2, self.bn2, relu=True, mask_output=True, premasked=True)
        out, veil = conv_bn(out, veil, self.conv3, self.bn3, relu=not self.skip_relu, residual=self._join(res),
                            premasked=True, res_holder=h if self.downsample is None else None)
        return out, veil


class Fusion(nn.Module):
    """1x1 conv over the channel concatenation of the two streams + BN + ReLU (fusionnet.py:130-140)."""

    def __init__(self, inplanes):
        super().__init__()
        self.conv = Conv2d(inplanes * 2, inplanes, kernel_size=1, bias=False)
        self.bn = BatchNorm2d(inplanes)

    def forward_nhwc(self, x, y):
        # the concatenation (fusionnet.py:137) is never materialised: the kernels read the two streams as they are
        z, _ = conv_bn(x, None, self.conv, self.bn, relu=True, x2=y)
        return z


KINDS = ("depthnet", "partial_depthnet", "fusionnet", "partial_fusionnet", "resnet")
ARCH = {"resnet18": (BasicBlock, (2, 2, 2, 2)), "resnet50": (Bottleneck, (3, 4, 6, 3))}


class ResNet(nn.Module):
    """Dilated ResNet-18/50 pose backbone emitting a depth*num_joints volumetric heat-map."""

    def __init__(self, kind, block, layers, args):
        assert kind in KINDS
        allowed = [16, 32] if kind == "resnet" else [4, 8, 16, 32]       # resnet.py:126
        assert args.stride in allowed
        if kind == "partial_depthnet":
            assert args.depth_only                                         # partial_depthnet.py:164
        super().__init__()
        self.kind = kind
        self.fused = kind in ("fusionnet", "partial_fusionnet")
        self.partial = kind.startswith("partial_")
        self.early_dist = bool(getattr(args, "early_dist", False)) and kind in ("depthnet", "fusionnet")
        self.skip_relu = bool(getattr(args, "skip_relu", False)) and kind in ("depthnet", "fusionnet")
        self.depth, self.num_joints = args.depth, args.num_joints
        self.compute_dtype = torch.float32
        strides, dils = stage_strides(args.stride)
        planes = (64, 128, 256, 512)

        if self.fused:
            self.conv1 = Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
            DepthStem = PartialConv if self.partial else Conv2d
            self.conv2 = DepthStem(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
            self.bn1 = BatchNorm2d(64)
            self.bn2 = BatchNorm2d(64)
        else:
            if kind == "resnet":
                cin = 4 if getattr(args, "extra_channel", False) else 3
            elif kind == "partial_depthnet":
                cin = 1
            else:
                cin = 1 if args.depth_only else 3
            Stem = PartialConv if self.partial else Conv2d
            self.conv1 = Stem(cin, 64, kernel_size=7, stride=2, padding=3, bias=False)
            self.bn1 = BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)     # geometry record only

        part12 = self.partial and not self.fused
        self.inplanes = 64
        self.layer1 = self._make_layer(block, planes[0], layers[0], partial=part12)
        self.layer2 = self._make_layer(block, planes[1], layers[1], strides[1], dils[1], partial=part12)
        if self.fused:
            self.fusion = Fusion(self.inplanes)
        self.layer3 = self._make_layer(block, planes[2], layers[2], strides[2], dils[2], skip_relu=self.skip_relu)
        self.layer4 = self._make_layer(block, planes[3], layers[3], strides[3], dils[3], skip_relu=self.skip_relu)
        if self.fused:
            self.inplanes = 64
            self.layer5 = self._make_layer(block, planes[0], layers[0], partial=self.partial)
            self.layer6 = self._make_layer(block, planes[1], layers[1], strides[1], dils[1], partial=self.partial)

        for m in self.modules():                    # fan_out Kaiming-normal, BN affine (1, 0)
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, (2.0 / n) ** 0.5)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

        head_in = 512 * block.expansion
        if kind == "resnet":
            self.cam_regressor = Conv2d(head_in, args.depth * args.num_joints, kernel_size=3, padding=1)
            self.mat_regressor = Conv2d(head_in, args.num_joints, kernel_size=3, padding=1) \
                if getattr(args, "joint_space", False) else None
        else:
            self.regressor = Conv2d(head_in, args.depth * args.num_joints, 3, padding=1)

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1, partial=False, skip_relu=False):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                Conv2d(self.inplanes, planes * block.expansion, 1, stride=stride, bias=False),
                BatchNorm2d(planes * block.expansion))
        seq = [block(self.inplanes, planes, stride, dilation, downsample, partial=partial,
                     skip_relu=skip_relu and blocks == 1)]
        self.inplanes = planes * block.expansion
        for i in range(1, blocks):
            seq.append(block(self.inplanes, planes, partial=partial, skip_relu=skip_relu and i == blocks - 1))
        return nn.Sequential(*seq)

    # ---- dtype policy: masters stay fp32, .half()/.bfloat16() select bf16 tensor-core compute ----
    def half(self):
        self.compute_dtype = torch.bfloat16
        return self

    def bfloat16(self):
        self.compute_dtype = torch.bfloat16
        return self

    def float(self):
        self.compute_dtype = torch.float32
        return super().float()

    def freeze_batchnorm(self):                     # depthnet.py:158-161
        for module in self.modules():
            if isinstance(module, nn.BatchNorm2d):
                module.eval()

    # ---- forward ----
    @staticmethod
    def _run(layer, x, veil):
        for blk in layer:
            x, veil = blk.forward_nhwc(x, veil)
        return x, veil

    def _stem(self, inp, conv, bn):
        x = ops.to_nhwc(inp, self.compute_dtype)
        partial = isinstance(conv, PartialConv)
        veil = ops.veil_from_depth(x) if partial else None          # partial_depthnet.py:215
        # veil = (x != 0), so x * veil == x exactly: the stem reads its input as already masked
        x, veil = conv_bn(x, veil, conv, bn, relu=True, premasked=True)
        return ops.MaxPoolFn.apply(x, veil)                          # x and veil pooled together (:219-220)

    def forward(self, x, y=None):
        if self.fused:
            if y is None:
                raise TypeError("forward() of a fusion net takes (color, depth)")
            if ops.TWO_STREAMS and y.is_cuda:
                # depth trunk on a second stream, concurrent with the RGB trunk (see ops.branch_stream)
                main, side = torch.cuda.current_stream(y.device), ops.branch_stream(y.device)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    b, veil = self._stem(y, self.conv2, self.bn2)
                    b, veil = self._run(self.layer5, b, veil)
                    b, veil = self._run(self.layer6, b, veil)
                a, _ = self._stem(x, self.conv1, self.bn1)
                a, _ = self._run(self.layer1, a, None)
                a, _ = self._run(self.layer2, a, None)
                main.wait_stream(side)
                b.record_stream(main)            # allocated on the side stream, consumed on this one
            else:
                a, _ = self._stem(x, self.conv1, self.bn1)
                b, veil = self._stem(y, self.conv2, self.bn2)
                a, _ = self._run(self.layer1, a, None)
                b, veil = self._run(self.layer5, b, veil)
                a, _ = self._run(self.layer2, a, None)
                b, veil = self._run(self.layer6, b, veil)
            f = self.fusion.forward_nhwc(a, b)
        else:
            f, veil = self._stem(x, self.conv1, self.bn1)
            f, veil = self._run(self.layer1, f, veil)
            f, veil = self._run(self.layer2, f, veil)
        # f = input of layer3: the boundary between the shallow (stems, layer1/2/5/6, fusion) and the deep part; a
        # data-parallel Trainer runs the backward pass in two stages around it so that the all-reduce of the deep
        # gradients (~90 % of the parameters) overlaps the shallow backward (trainer.Trainer._fwd_bwd_deep)
        self._boundary = None
        if getattr(self, "_mark_boundary", False):
            # cut the autograd graph here: stage 1 runs loss.backward() down to the detached leaf, stage 2 continues
            # from `f` with the leaf's gradient (backward(inputs=[f]) would also RUN the node that produced f, and the
            # second stage would then accumulate that node's parameter gradients twice)
            cut = f.detach().requires_grad_()
            self._boundary, f = (f, cut), cut
        m, _ = self._run(self.layer3, f, None)
        n, _ = self._run(self.layer4, torch.relu(m) if self.skip_relu else m, None)
        top = torch.relu(n) if self.skip_relu else n
        if self.kind == "resnet":
            cam = self.cam_regressor.forward_nhwc(top).permute(0, 3, 1, 2)
            if self.mat_regressor is not None:
                return cam, self.mat_regressor.forward_nhwc(top).permute(0, 3, 1, 2)
            return cam
        z = self.regressor.forward_nhwc(top)
        last = m if self.early_dist else n
        return z.permute(0, 3, 1, 2), last.permute(0, 3, 1, 2)


# ------------------------------------------------------------------ ImageNet-checkpoint surgery
def _load_toy(args, kind):
    if kind in ("depthnet", "fusionnet") and getattr(args, "depth_host", False):
        return torch.load(args.host_path, map_location="cpu")["model"]
    return torch.load(args.model_path, map_location="cpu")


def load_pretrained(model, kind, args):
    """`pretrain=True` behaviour of the reference builders (partial_depthnet.py:232-257,
    depthnet.py:203-229, fusionnet.py:243-297, resnet.py:213-262): start from an ImageNet ResNet
    checkpoint, slice / replicate the stem to the input channel count, clone the RGB trunk into
    the depth stream, drop unknown keys."""
    own = model.state_dict()
    toy = {k: v.clone() for k, v in _load_toy(args, kind).items()}
    stem = toy["conv1.weight"]
    manual = {}
    if kind in ("fusionnet", "partial_fusionnet"):
        for key in own:
            for dst, src in (("bn2", "bn1"), ("layer5", "layer1"), ("layer6", "layer2")):
                if key.startswith(dst) and key.replace(dst, src, 1) in toy:
                    manual[key] = toy[key.replace(dst, src, 1)].clone()
        manual["conv2.weight"] = stem[:, :1].clone()
        if getattr(args, "depth_host", False) and kind == "fusionnet":
            toy = {k: v.clone() for k, v in torch.load(args.model_path, map_location="cpu").items()}
        missing = [k for k in own if k not in toy and k not in manual and not k.endswith("num_batches_tracked")]
        assert all(k.startswith(("fusion", "regressor")) for k in missing), missing
    elif kind == "resnet":
        if getattr(args, "extra_channel", False):
            widened = own["conv1.weight"].clone()
            widened[:, :3] = stem
            toy["conv1.weight"] = widened
    else:
        if kind == "partial_depthnet" or args.depth_only:
            toy["conv1.weight"] = stem[:, :1].clone()
        if kind == "depthnet" and getattr(args, "depth_host", False):
            toy["conv1.weight"] = (toy["conv1.weight"] / 3).repeat(1, 3, 1, 1)
    for key in list(toy):
        if key not in own:
            print("key [", key, "] deleted")
            del toy[key]
    own.update(manual)
    own.update(toy)
    model.load_state_dict(own)
    return model


def build(kind, model_name, args, pretrain):
    block, layers = ARCH[model_name]
    model = ResNet(kind, block, list(layers), args)
    return load_pretrained(model, kind, args) if pretrain else model
