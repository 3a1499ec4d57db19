"""Drop-in for ``back_project.projectPoints`` (back_project.py:12-36): world points -> distorted pixel
coordinates of a CMU-Panoptic camera dict (K, R, t, distCoef), computed on the GPU."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _arr(v, n):
    a = np.asarray(v, np.float64).reshape(-1)
    if a.size != n:
        raise ValueError("expected %d values, got %d" % (n, a.size))
    return (C.c_float * n)(*[float(x) for x in a])


def projectPoints(X, cam):
    """X: (3, N) numpy array / matrix or CUDA tensor; cam: dict with 'K' (3x3), 'R' (3x3), 't' (3x1),
    'distCoef' (5,).  Returns (3, N) = (u, v, camera z) of the input's kind."""
    as_numpy = not torch.is_tensor(X)
    if as_numpy:
        dev = torch.device("cuda", torch.cuda.current_device())
        Xt = torch.as_tensor(np.ascontiguousarray(np.asarray(X, np.float32))).to(dev)
    else:
        L.require_cuda(X)
        Xt = X.float().contiguous()
    if Xt.dim() != 2 or Xt.shape[0] != 3:
        raise ValueError("X must be 3 x N")
    n = Xt.shape[1]
    out = torch.empty_like(Xt)
    L.call("b2_project_points", L.ptr(Xt), n, _arr(cam["R"], 9), _arr(cam["t"], 3), _arr(cam["K"], 9),
           _arr(cam["distCoef"], 5), L.ptr(out), L.stream())
    return out.cpu().numpy().astype(np.float64) if as_numpy else out
