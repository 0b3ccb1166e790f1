"""Harness around the UNMODIFIED reference extension built by oracle/build_ref.sh (oracle/_ref/_C.so).

TEST / BENCH INFRASTRUCTURE ONLY (parity checker on the GPU box, reference arm of bench.py).
Nothing in vidu4d_b200/ imports this.  The pybind module is loaded straight from the .so; the reference's
Python wrapper is not copied -- the thin autograd-free calls below follow
RAST/diff_surfel_rasterization/__init__.py:60-98,109-143 (argument order of _C.rasterize_gaussians /
_C.rasterize_gaussians_backward).

decode_*() read the reference's opaque byte buffers; layout from RAST/cuda_rasterizer/rasterizer_impl.cu:155-194
and rasterizer_impl.h:21-27 (every sub-array start rounded up to 128 B).
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_ref", "_C.so")
_mod = None


def available() -> bool:
    return os.path.exists(SO_PATH)


def load():
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError(f"{SO_PATH} missing: run `bash oracle/build_ref.sh` where /root/reference exists")
        spec = importlib.util.spec_from_file_location("_C", SO_PATH)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def _e(dev):
    return torch.empty((0,), dtype=torch.float32, device=dev)


def forward(means3D, opacities, scales, rotations, shs=None, colors_precomp=None, *, sh_degree, W, H, tanfovx,
            tanfovy, bg, viewmatrix, projmatrix, campos, debug=False):
    """All tensor args are CUDA float32 tensors.  Returns dict with the 7 outputs of _C.rasterize_gaussians."""
    C = load()
    dev = means3D.device
    out = C.rasterize_gaussians(
        bg, means3D, colors_precomp if colors_precomp is not None else _e(dev), opacities, scales, rotations, 1.0,
        _e(dev), viewmatrix, projmatrix, float(tanfovx), float(tanfovy), int(H), int(W),
        shs if shs is not None else _e(dev), int(sh_degree), campos, False, bool(debug))
    keys = ("num_rendered", "color", "allmap", "radii", "geomBuffer", "binningBuffer", "imgBuffer")
    return dict(zip(keys, out))


def backward(fw, means3D, scales, rotations, shs=None, colors_precomp=None, *, dL_dcolor, dL_dallmap, sh_degree,
             tanfovx, tanfovy, bg, viewmatrix, projmatrix, campos, debug=False):
    C = load()
    dev = means3D.device
    out = C.rasterize_gaussians_backward(
        bg, means3D, fw["radii"], colors_precomp if colors_precomp is not None else _e(dev), scales, rotations, 1.0,
        _e(dev), viewmatrix, projmatrix, float(tanfovx), float(tanfovy), dL_dcolor, dL_dallmap,
        shs if shs is not None else _e(dev), int(sh_degree), campos, fw["geomBuffer"], int(fw["num_rendered"]),
        fw["binningBuffer"], fw["imgBuffer"], bool(debug))
    keys = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dtransMat", "dL_dsh", "dL_dscales",
            "dL_drotations")
    return dict(zip(keys, out))


class _Cursor:
    def __init__(self, buf: torch.Tensor):
        self.buf = buf
        self.base = buf.data_ptr()
        self.cur = self.base

    def take(self, count, dtype):
        itemsize = torch.empty((), dtype=dtype).element_size()
        start = (self.cur + 127) & ~127
        off = start - self.base
        self.cur = start + count * itemsize
        return self.buf[off:off + count * itemsize].view(dtype)


def decode_geom(geomBuffer, P):
    c = _Cursor(geomBuffer)
    d = {}
    d["depths"] = c.take(P, torch.float32)
    d["clamped"] = c.take(3 * P, torch.uint8).view(P, 3)
    d["internal_radii"] = c.take(P, torch.int32)
    d["means2D"] = c.take(2 * P, torch.float32).view(P, 2)
    d["transMat"] = c.take(9 * P, torch.float32).view(P, 9)
    d["normal_opacity"] = c.take(4 * P, torch.float32).view(P, 4)
    d["rgb"] = c.take(3 * P, torch.float32).view(P, 3)
    d["tiles_touched"] = c.take(P, torch.int32)
    return d


def decode_binning(binningBuffer, R):
    c = _Cursor(binningBuffer)
    d = {}
    d["point_list"] = c.take(R, torch.int32)
    d["point_list_unsorted"] = c.take(R, torch.int32)
    d["keys"] = c.take(R, torch.int64)
    d["keys_unsorted"] = c.take(R, torch.int64)
    return d


def decode_image(imgBuffer, W, H):
    N = W * H
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    c = _Cursor(imgBuffer)
    d = {}
    d["final_T"] = c.take(3 * N, torch.float32).view(3, H, W)
    d["n_contrib"] = c.take(2 * N, torch.int32).view(2, H, W)
    d["ranges"] = c.take(2 * N, torch.int32).view(N, 2)[:tiles]
    return d


def to_np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


# ---------------------------------------------------------------------------------------------
# autograd wrapper over the reference `_C` (for bench.py --impl reference and the API-level parity test):
# same forward/backward contract as RAST/diff_surfel_rasterization/__init__.py:44-156,172-222.
# ---------------------------------------------------------------------------------------------
class _RefRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        C = load()
        out = C.rasterize_gaussians(rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                                    cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, float(rs.tanfovx), float(rs.tanfovy),
                                    rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = out
        ctx.rs = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        C = load()
        rs = ctx.rs
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer = ctx.saved_tensors
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, rs.image_height, rs.image_width), device=means3D.device)
        if grad_depth is None:
            grad_depth = torch.zeros((8, rs.image_height, rs.image_width), device=means3D.device)
        g = C.rasterize_gaussians_backward(rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier,
                                           cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, float(rs.tanfovx),
                                           float(rs.tanfovy), grad_out_color.contiguous(), grad_depth.contiguous(), sh,
                                           rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered, binningBuffer,
                                           imgBuffer, rs.debug)
        gm2, gcol, gop, gm3, gcov, gsh, gsc, grot = g
        return (gm3, gm2, gsh if sh.numel() else None, gcol if colors_precomp.numel() else None, gop, gsc, grot,
                None, None)


class RefGaussianRasterizer(torch.nn.Module):
    """Drop-in for GaussianRasterizer backed by the reference extension (bench reference arm / parity tests)."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        dev = means3D.device
        e = lambda: torch.empty((0,), dtype=torch.float32, device=dev)  # noqa: E731
        return _RefRasterize.apply(means3D, means2D, shs if shs is not None else e(),
                                   colors_precomp if colors_precomp is not None else e(), opacities,
                                   scales if scales is not None else e(), rotations if rotations is not None else e(),
                                   cov3D_precomp if cov3D_precomp is not None else e(), self.raster_settings)
