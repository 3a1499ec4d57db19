"""On-device input pipeline: the per-sample CPU work of
``depth_datasets.Dataset.parse_sample`` (depth_datasets.py:199-237) as batched GPU kernels --
homography crop (``cameralib.reproject_image_fast``, cameralib.py:667-711) fused with
ToTensor + Normalize for the colour frame, and with ``utils.to_depth`` + ``enhance_ntu`` /
``enhance_pku`` (depth_datasets.py:39-56) for the depth frame.  Camera bookkeeping (turn_towards,
zoom, flip ... ) stays on the host: it is a handful of 3x3 products per sample.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

MEAN = (0.485, 0.456, 0.406)          # depth_datasets.py:78-79
DEV = (0.229, 0.224, 0.225)
VEIL_THRESHOLD = {"ntu": 0.1, "pku": 0.5}      # enhance_ntu / enhance_pku, depth_datasets.py:42,52


def homography(old_camera, new_camera):
    """float32 3x3 of cameralib.reproject_image_fast (cameralib.py:672-674): destination pixel -> source
    pixel.  Cameras: objects with ``intrinsic_matrix`` and ``R`` (like cameralib.Camera) or (K, R) pairs."""
    def kr(cam):           # dtypes are kept: the products must round as the reference's do (float32 members,
        if isinstance(cam, (tuple, list)):          # float64 intrinsics after square_pixels() / zoom())
            return np.asarray(cam[0]), np.asarray(cam[1])
        return np.asarray(cam.intrinsic_matrix), np.asarray(cam.R)
    k0, r0 = kr(old_camera)
    k1, r1 = kr(new_camera)
    return ((k0 @ r0) @ np.linalg.inv(k1 @ r1)).astype(np.float32)


def _homs(h, n, device):
    h = torch.as_tensor(np.ascontiguousarray(h, np.float32)) if not torch.is_tensor(h) else h.float()
    h = h.reshape(-1, 9)
    if h.shape[0] != n:
        raise ValueError("need one 3x3 homography per image (%d), got %d" % (n, h.shape[0]))
    return h.to(device).contiguous()


def crop_normalize_rgb(frames, homographies, side_in, mean=MEAN, std=DEV):
    """uint8 frames [N, Hs, Ws, 3] (CUDA) -> normalised fp32 [N, 3, side_in, side_in]:
    ``transform(reproject_image(image, camera, new_cam, (side_in, side_in)))`` of depth_datasets.py:196,212."""
    L.require_cuda(frames)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise TypeError("frames must be a uint8 [N, H, W, 3] tensor")
    frames = frames.contiguous()
    N, Hs, Ws, _ = frames.shape
    hom = _homs(homographies, N, frames.device)
    out = torch.empty((N, 3, side_in, side_in), dtype=torch.float32, device=frames.device)
    L.call("b2_remap_normalize_rgb", L.ptr(frames), N, Hs, Ws, L.ptr(hom), int(side_in),
           (C.c_float * 3)(*[float(v) for v in mean]), (C.c_float * 3)(*[float(v) for v in std]), L.ptr(out), L.stream())
    return out


def _cams(intrinsics, n, device):
    if intrinsics is None:
        return None
    k = np.asarray(intrinsics, np.float32).reshape(-1, 3, 3)
    if k.shape[0] == 1 and n > 1:
        k = np.repeat(k, n, axis=0)
    if k.shape[0] != n:
        raise ValueError("need one intrinsic matrix per image (%d), got %d" % (n, k.shape[0]))
    rows = [np.concatenate([np.linalg.inv(m[:2, :2]).astype(np.float32).reshape(-1), m[:2, 2]]) for m in k]
    return torch.as_tensor(np.stack(rows).astype(np.float32)).to(device).contiguous()


def crop_enhance_depth(frames, homographies, side_in, data_name="ntu", nexponent=True, to_depth_intrinsics=None,
                       enhance=True):
    """fp32 depth frames [N, Hs, Ws] (CUDA) -> [N, 1, side_in, side_in]: reproject_image -> (utils.to_depth with
    ``to_depth_intrinsics`` [N,3,3] or one 3x3) -> enhance_<data_name> (depth_datasets.py:197,214-217).
    ``homographies`` None: the frames are already cropped to side_in."""
    L.require_cuda(frames)
    if frames.dtype != torch.float32 or frames.dim() != 3:
        raise TypeError("depth frames must be a float32 [N, H, W] tensor")
    if data_name not in VEIL_THRESHOLD:
        raise ValueError("data_name must be one of %s" % sorted(VEIL_THRESHOLD))
    frames = frames.contiguous()
    N, Hs, Ws = frames.shape
    hom = None
    if homographies is not None:
        hom = _homs(homographies, N, frames.device)
    elif Hs != side_in or Ws != side_in:
        raise ValueError("without homographies the frames must already be %dx%d" % (side_in, side_in))
    cam = _cams(to_depth_intrinsics, N, frames.device)
    out = torch.empty((N, 1, side_in, side_in), dtype=torch.float32, device=frames.device)
    L.call("b2_remap_enhance_depth", L.ptr(frames), N, Hs, Ws, L.ptr(hom), int(side_in), L.ptr(cam),
           float(VEIL_THRESHOLD[data_name]), int(bool(nexponent)), int(bool(enhance)), L.ptr(out), L.stream())
    return out


def enhance_ntu(image, nexponent):
    """depth_datasets.enhance_ntu (:39-46) on a CUDA tensor [..., H, W] (H == W) -> [..., 1, H, W]."""
    return _enhance(image, "ntu", nexponent)


def enhance_pku(image, nexponent):
    """depth_datasets.enhance_pku (:49-56)."""
    return _enhance(image, "pku", nexponent)


def _enhance(image, name, nexponent):
    L.require_cuda(image)
    h, w = image.shape[-2:]
    if h != w:
        raise ValueError("square crops only (side_in x side_in), got %dx%d" % (h, w))
    flat = image.float().reshape(-1, h, w)
    out = crop_enhance_depth(flat, None, h, name, nexponent)
    return out.reshape(tuple(image.shape[:-2]) + (1, h, w))
