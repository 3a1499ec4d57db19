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


def mimic_loss(teach_last, last_feat, atten_map, sigmoid=False, bin_dist=False):
    """Trainer.distill (depth_train.py:115-129): the feature-mimic term of the distillation step.
    ``bin_dist`` selects the BCE variant (:117-121), ``sigmoid`` the squashed L2 (:123), default the
    attention-weighted L2 norm per sample, averaged over the batch."""
    mode = ops.MIMIC_MODES["bce" if bin_dist else ("sigmoid" if sigmoid else "l2")]
    return ops.MimicLossFn.apply(teach_last, last_feat, atten_map, mode)


def get_attention(side_in, stride, image_coords, attention):
    """utils.get_attention (utils.py:14-42).  image_coords: (J, 2) numpy array -> numpy (1, S', S') like the
    reference (computed on the GPU); a CUDA tensor [N, J, 2] -> CUDA tensor [N, 1, S', S']."""
    side_out = (side_in - 1) // stride + 1
    if torch.is_tensor(image_coords) and image_coords.is_cuda:
        if not attention:
            return torch.ones((image_coords.shape[0], 1, side_out, side_out), device=image_coords.device)
        return ops.attention_map(image_coords, side_in, side_out)
    if not attention:
        return np.ones((1, side_out, side_out))
    dev = torch.device("cuda", torch.cuda.current_device())
    c = torch.as_tensor(np.ascontiguousarray(image_coords, np.float32)).to(dev)[None]
    return ops.attention_map(c, side_in, side_out)[0].cpu().numpy().astype(np.float64)


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


METRIC_KEYS = ("solid", "close", "depth", "jitter", "switch", "fail")


class MetricAccumulator:
    """Epoch-level ``utils.analyze`` + ``utils.parse_epoch`` (utils.py:197-262) on the device: every batch
    is one kernel launch into a 10-double accumulator; ``result()`` is the single host read-back.  The
    batch-size-weighted means of parse_epoch equal the ratios of the epoch totals kept here."""

    def __init__(self, mirror, thresh, device):
        self.thresh = (float(thresh["solid"]), float(thresh["close"]), float(thresh["rough"]))
        self.mirror = None if mirror is None else torch.as_tensor(np.asarray(mirror, np.int32)).to(device)
        self.acc = torch.zeros(10, dtype=torch.float64, device=device)

    def update(self, spec_cam, true_cam, valid_mask, back_rotate=None):
        L = ops.L
        L.require_cuda(spec_cam, true_cam, valid_mask, back_rotate)
        spec = spec_cam.detach().float().contiguous()
        true = true_cam.detach().float().contiguous()
        valid = valid_mask.to(torch.uint8).contiguous()
        rot = None if back_rotate is None else back_rotate.detach().float().contiguous()
        N, J, _ = spec.shape
        L.call("b2_pose_metrics", L.ptr(spec), L.ptr(true), L.ptr(valid), L.ptr(rot), L.ptr(self.mirror), N, J,
               self.thresh[0], self.thresh[1], self.thresh[2], L.ptr(self.acc), L.stream())

    def result(self):
        a = self.acc.cpu().numpy()
        n = max(a[0], 1.0)
        out = {k: float(a[4 + i] / n) for i, k in enumerate(METRIC_KEYS)}
        out.update(score_pck=float(a[2] / n), score_auc=float(a[3] / n), cam_mean=float(a[1] / n), batch_size=int(a[0]))
        return out


def analyze(spec_cam, true_cam, valid_mask, mirror, thresh, back_rotate=None):
    """utils.analyze (utils.py:234-262) for CUDA tensors: dict(batch_size, score_pck, score_auc, cam_mean and the
    error taxonomy solid / close / depth / jitter / switch / fail).  ``back_rotate`` [N,3,3] applies the
    einsum of the test loops (depth_train.py:522-523) first."""
    acc = MetricAccumulator(mirror, thresh, spec_cam.device)
    acc.update(spec_cam, true_cam, valid_mask, back_rotate)
    return acc.result()


def parse_epoch(stats):
    """utils.parse_epoch (utils.py:224-231): batch-size weighted means of a list of ``analyze`` dicts."""
    keys = ("solid", "close", "jitter", "depth", "switch", "fail", "score_pck", "score_auc", "cam_mean", "batch_size")
    values = np.array([[patch[key] for patch in stats] for key in keys], np.float64)
    return dict(zip(keys[:-1], np.sum(values[-1] * values[:-1], axis=1) / np.sum(values[-1])))


def mpjpe(spec_cam, true_cam, valid):
    """``cam_mean`` of utils.analyze (utils.py:253-262): mean joint distance (mm) over valid joints."""
    d = torch.linalg.norm(spec_cam.float() - true_cam.float(), dim=-1).reshape(-1)
    return float(d[valid.reshape(-1).bool()].mean())
