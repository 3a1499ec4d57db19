"""Drop-in for the hot-path functions of the reference's ``utils`` module:
``to_heatmap`` (utils.py:154-175), ``decode`` (utils.py:178-194), ``to_depth`` (utils.py:68-75),
plus the fused form the Trainer uses (``heatmap_coords``) and the MPJPE of ``analyze``
(utils.py:253-262).  All device work runs in libb2pose kernels."""
import numpy as np
import torch

from . import ops


def to_heatmap(feat, depth, num_joints, height, width):
    """feat [N, depth*num_joints, H, W] (channel = d*num_joints + j) -> softmax over the H*W*D
    voxels of every (sample, joint), returned as [N, num_joints, H, W, depth] fp32."""
    return ops.ToHeatmapFn.apply(feat, depth, num_joints, height, width)


def decode(heatmap, depth_range):
    """Soft-argmax of a [N, J, H, W, D] heat-map -> [N, J, 3] = (x<-W, y<-H, z<-D) * depth_range."""
    return ops.DecodeFn.apply(heatmap, depth_range)


def heatmap_coords(feat, depth, num_joints, depth_range):
    """decode(to_heatmap(feat)) in a single pass over the logits (never materialises the heat-map)."""
    return ops.HeadFn.apply(feat, depth, num_joints, depth_range)


def pose_loss(coords, true_cam, true_val, key_index, loss_div=10.0, criterion="SmoothL1"):
    """Root-relative shift + masked mean loss of depth_train.py:397-405 -> (loss, spec_cam)."""
    return ops.PoseLossFn.apply(coords, true_cam, true_val, key_index, loss_div, criterion)


def to_depth(image, depth_cam):
    """Ray length -> z-depth, ``image / sqrt(|image_to_camera(u, v)|^2 + 1)`` (utils.py:68-75 with
    cameralib.Camera.image_to_camera, no-distortion branch cameralib.py:192-194).

    image: [H, W] (or [..., H, W]) numpy array or CUDA tensor; depth_cam: an object with
    ``intrinsic_matrix`` (like cameralib.Camera) or a 3x3 matrix.  Returns the input's kind."""
    K = getattr(depth_cam, "intrinsic_matrix", depth_cam)
    dist = getattr(depth_cam, "distortion_coeffs", None)
    if dist is not None and np.any(np.asarray(dist) != 0):
        raise NotImplementedError("to_depth: lens distortion is not built (NTU depth cameras have none, "
                                  "get_depth_cams.py:89-90)")
    if torch.is_tensor(image):
        return ops.unproject_depth(image, K)
    dev = torch.device("cuda", torch.cuda.current_device())
    out = ops.unproject_depth(torch.as_tensor(np.ascontiguousarray(image, np.float32)).to(dev), K)
    return out.cpu().numpy()


def mpjpe(spec_cam, true_cam, valid):
    """``cam_mean`` of utils.analyze (utils.py:253-262): mean joint distance (mm) over valid joints."""
    d = torch.linalg.norm(spec_cam.float() - true_cam.float(), dim=-1).reshape(-1)
    return float(d[valid.reshape(-1).bool()].mean())
