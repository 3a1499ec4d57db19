"""Drop-in for the reference's legacy ``resnet`` module (resnet.py:213-262): ``resnet18(args)`` /
``resnet50(args)`` with ``args.pretrain``; forward returns ``cam_feat`` or ``(cam_feat, mat_feat)``."""
from . import nets
from .nets import BasicBlock, Bottleneck  # noqa: F401

KIND = "resnet"


class ResNet(nets.ResNet):
    def __init__(self, block, layers, args):
        super().__init__(KIND, block, layers, args)


def _build(block, layers, args):
    model = ResNet(block, layers, args)
    return nets.load_pretrained(model, KIND, args) if getattr(args, "pretrain", False) else model


def resnet18(args):
    return _build(BasicBlock, [2, 2, 2, 2], args)


def resnet50(args):
    return _build(Bottleneck, [3, 4, 6, 3], args)
