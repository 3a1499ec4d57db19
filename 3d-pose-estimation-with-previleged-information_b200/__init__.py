"""b2pose: B200-native depth-stream hot path of 3D-Pose-Estimation-with-Previleged-Information.

Public surface = the reference's module surface for this path:
``PartialConv`` / ``PartialConv2d``, the ``depthnet`` / ``partial_depthnet`` / ``fusionnet`` /
``partial_fusionnet`` / ``resnet`` constructor modules, ``utils.to_heatmap`` / ``decode`` /
``to_depth`` and the ``Trainer`` loops; underneath, hand-written sm_100a kernels in
``libb2pose.so`` (C ABI in ``include/b2pose.h``).  Nothing here falls back to CPU or eager
PyTorch math: a missing library or a non-CUDA tensor raises.
"""
from . import _lib, ops, layers, nets, utils, pipeline, mat_utils, back_project  # noqa: F401
from . import partial_conv, depthnet, partial_depthnet, fusionnet, partial_fusionnet, resnet  # noqa: F401
from .layers import PartialConv, PartialConv2d, Conv2d, BatchNorm2d  # noqa: F401
from .trainer import Trainer, train_args, synthetic_batch  # noqa: F401
from .utils import to_heatmap, decode, to_depth, heatmap_coords, pose_loss, mpjpe, mimic_loss, get_attention, analyze, parse_epoch, MetricAccumulator  # noqa: F401
