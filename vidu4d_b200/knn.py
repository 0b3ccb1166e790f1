"""`distCUDA2` of the reference's simple-knn submodule (gs/submodules/simple-knn/simple_knn.cu, ext.cpp): mean squared
distance of every point to its three nearest neighbours, used once to initialise the surfel scales
(gs/scene/gaussian_model.py:139-140: `scales = log(sqrt(clamp_min(distCUDA2(points), 1e-7)))`)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """points (P,3) float32 CUDA tensor -> (P,) float32: (d0 + d1 + d2) / 3 over the three nearest OTHER points."""
    lib = _capi.load()
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor")      # the reference's extension has no CPU path either
    pts = points.detach().to(torch.float32).contiguous()
    P = int(pts.shape[0])
    out = torch.empty((P,), dtype=torch.float32, device=pts.device)
    if P == 0:
        return out
    lo, hi = pts.amin(0).cpu(), pts.amax(0).cpu()                # one host sync: this is an initialisation-time routine
    bmin = (C.c_float * 3)(*[float(v) for v in lo]); bmax = (C.c_float * 3)(*[float(v) for v in hi])
    cells = int(lib.sr_knn_cells(P, bmin, bmax, None, None))
    scratch = torch.empty((2 * P + 3 * cells,), dtype=torch.int32, device=pts.device)
    with torch.cuda.device(pts.device):
        rc = lib.sr_knn_mean_dist2(P, pts.data_ptr(), bmin, bmax, out.data_ptr(), scratch.data_ptr(),
                                   torch.cuda.current_stream(pts.device).cuda_stream)
    _capi.check(rc, "sr_knn_mean_dist2")
    return out
