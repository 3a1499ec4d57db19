"""Drop-in for the reference's ``fusionnet`` module: ``resnet18(args, pretrain)`` / ``resnet50(args, pretrain)``
(fusionnet.py:300-305), resolved by name from ``depth_main.create_model`` (depth_main.py:36-45)."""
from . import nets
from .layers import PartialConv  # noqa: F401  (the reference module re-exports it)
from .nets import BasicBlock, Bottleneck, Fusion  # noqa: F401

KIND = "fusionnet"


class ResNet(nets.ResNet):
    def __init__(self, block, layers, args):
        super().__init__(KIND, block, layers, args)


def build_resnet(block, layers, args, pretrain):
    model = ResNet(block, layers, args)
    return nets.load_pretrained(model, KIND, args) if pretrain else model


def resnet18(args, pretrain):
    return build_resnet(BasicBlock, [2, 2, 2, 2], args, pretrain)


def resnet50(args, pretrain):
    return build_resnet(Bottleneck, [3, 4, 6, 3], args, pretrain)
