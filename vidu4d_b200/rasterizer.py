"""Python host side of the surfel rasterizer: the reference's operator interface, re-implemented over
the C ABI of libsurfel_raster.so.

Mirrors, name for name, RAST/diff_surfel_rasterization/__init__.py (RAST = gs/submodules/
diff-surfel-rasterization of yikaiw/Vidu4D):

    GaussianRasterizationSettings   __init__.py:158-170
    GaussianRasterizer              __init__.py:172-222   (forward, markVisible)
    rasterize_gaussians             __init__.py:21-42
    _RasterizeGaussians             __init__.py:44-156    (autograd.Function)
    _C.rasterize_gaussians / _C.rasterize_gaussians_backward / _C.mark_visible
                                    RAST/rasterize_points.h:18-68, RAST/ext.cpp:15-19

so `gs/gaussian_renderer/__init__.py:14` (`from diff_surfel_rasterization import ...`) works unchanged
when the repo root is on sys.path (see the top-level `diff_surfel_rasterization/` shim).

PyTorch is plumbing here: it owns device memory and streams; every kernel is in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _capi

# --------------------------------------------------------------------------------------------
# capacity policy for the instance ("binning") buffer
# --------------------------------------------------------------------------------------------
_CAP_ALIGN = 256           # capacities are multiples of this so that bytes -> capacity is invertible
_cap_hint: dict = {}       # (device index, W, H) -> last seen num_rendered
_sort_global: set = set()  # keys for which a tile outgrew the shared-memory sort: use the global onesweep path
_local_sort_default = False  # the tile-local sort (SR_FLAG_LOCAL_SORT) is opt-in: see DESIGN.md section 3.2
_sync_mode = True          # True: read num_rendered back after every forward (like the reference)
_pending: list = []        # nosync mode: (pinned host word, key, capacity) awaiting check_overflow()
_host_pool: list = []      # pinned uint32[2] words, pre-allocated so that a forward never allocates pinned memory
_host_next = 0             #   (cudaHostAlloc is illegal during CUDA-graph capture)
# The module state above belongs to the (single) thread that issues forwards; autograd's backward thread only calls
# rasterize_gaussians_backward, which touches none of it.


def set_sync_mode(flag: bool):
    """sync=True (default): each forward reads `num_rendered` back (one stream sync, as the reference does at
    rasterizer_impl.cu:282) and transparently re-runs with a larger buffer on overflow.
    sync=False: no host<->device synchronisation at all; call check_overflow() once per step."""
    global _sync_mode
    _sync_mode = bool(flag)


def _round_cap(n: int) -> int:
    n = max(int(n), _CAP_ALIGN)
    return (n + _CAP_ALIGN - 1) // _CAP_ALIGN * _CAP_ALIGN


def _pick_capacity(key, P: int) -> int:
    r = _cap_hint.get(key)
    if r is None:
        return _round_cap(max(4 * P, 4096))
    return _round_cap(int(r * 1.5) + 4096)     # headroom for frame-to-frame growth in nosync mode (136 B / instance)


_bytes_to_cap: dict = {}


def _capacity_from_bytes(nbytes: int, W: int, H: int) -> int:
    """Inverse of sr_binning_bytes(., W, H) for capacities that are multiples of _CAP_ALIGN."""
    key = (nbytes, W, H)
    if key in _bytes_to_cap:
        return _bytes_to_cap[key]
    lib = _capi.load()
    lo, hi = 1, 1 << 22   # in units of _CAP_ALIGN
    while lo < hi:
        mid = (lo + hi) // 2
        if lib.sr_binning_bytes(mid * _CAP_ALIGN, W, H) < nbytes:
            lo = mid + 1
        else:
            hi = mid
    cap = lo * _CAP_ALIGN
    if lib.sr_binning_bytes(cap, W, H) != nbytes:
        raise _capi.SurfelRasterError("binningBuffer was not produced by this library's forward")
    _bytes_to_cap[key] = cap
    return cap


def reserve_host_slots(n: int):
    """Pre-allocate pinned status words for the next `n` nosync forwards (needed before CUDA-graph capture)."""
    while len(_host_pool) - _host_next < n:
        _host_pool.append(torch.zeros((2,), dtype=torch.int32).pin_memory())


def _host_slot():
    global _host_next
    if _host_next >= len(_host_pool):
        _host_pool.append(torch.zeros((2,), dtype=torch.int32).pin_memory())
    h = _host_pool[_host_next]
    _host_next += 1
    return h


def check_overflow(keep: bool = False):
    """nosync mode: synchronise once, verify that no forward since the last call overflowed its instance
    buffer, and refresh the capacity hints.  Raises SurfelRasterError if a frame was dropped.
    keep=True leaves the watch list in place (a captured CUDA graph rewrites the same status words on every
    replay: call check_overflow(keep=True) after each replay)."""
    global _host_next
    bad, prefilter = None, False
    if _pending:
        torch.cuda.synchronize()
    for host, key, cap in _pending:
        r, status = int(host[0]), int(host[1])
        _cap_hint[key] = max(_cap_hint.get(key, 0), r)
        if status & _capi.SR_STATUS_SORT_CAP:
            _sort_global.add(key)
            bad = (r, cap)
        if status & _capi.SR_STATUS_OVERFLOW:
            bad = (r, cap)
        if status & _capi.SR_STATUS_PREFILTER:
            prefilter = True
    if not keep:
        _pending.clear()
        _host_next = 0
    if prefilter:
        raise RuntimeError("Point is filtered although prefiltered is set. This shouldn't happen!")
    if bad:
        raise _capi.SurfelRasterError(
            f"instance buffer overflow (or a tile beyond the shared-memory sort) in nosync mode: num_rendered={bad[0]}, "
            f"capacity={bad[1]}; the frame was not rendered. Hints updated -- re-run the step.")


def _ptr(t):
    """Device pointer of a tensor, or None for the 'not provided' empty tensor (the reference's kernels test
    for a null data pointer the same way, forward.cu:215,247)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")   # CHECK_INPUT, rasterize_points.cu:27-28
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _as_float(x) -> float:
    # the reference passes CUDA 0-d tensors to pybind `float` params (gaussian_renderer/__init__.py:36-43)
    return float(x.item()) if torch.is_tensor(x) else float(x)


class _CNamespace:
    """Stand-in for the reference's pybind module `_C` with the same three entry points."""

    @staticmethod
    def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                            transMat_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                            image_width, sh, degree, campos, prefiltered, debug):
        lib = _capi.load()
        if means3D.ndim != 2 or means3D.shape[1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        if scales.ndim != 2 or scales.shape[1] != 2:
            raise RuntimeError("scales must have dimensions (num_points, 2)")
        if rotations.ndim != 2 or rotations.shape[1] != 4:
            raise RuntimeError("rotations must have dimensions (num_points, 4)")
        if transMat_precomp is not None and transMat_precomp.numel() != 0:
            raise RuntimeError("precomputed transMat / cov3D is not supported (the reference path for it is "
                               "broken: forward.cu:214-218 leaves the normal uninitialised)")
        dev = means3D.device
        P, H, W = int(means3D.shape[0]), int(image_height), int(image_width)
        background = _f32c(background, "background"); means3D = _f32c(means3D, "means3D")
        colors = _f32c(colors, "colors"); opacity = _f32c(opacity, "opacity")
        scales = _f32c(scales, "scales"); rotations = _f32c(rotations, "rotations")
        viewmatrix = _f32c(viewmatrix, "viewmatrix"); projmatrix = _f32c(projmatrix, "projmatrix")
        sh = _f32c(sh, "sh"); campos = _f32c(campos, "campos")
        M = int(sh.shape[1]) if sh is not None and sh.numel() != 0 else 0

        with torch.cuda.device(dev):
            out_color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
            out_others = torch.empty((8, H, W), dtype=torch.float32, device=dev)
            radii = torch.empty((P,), dtype=torch.int32, device=dev)
            geom = torch.empty((lib.sr_geom_bytes(P),), dtype=torch.uint8, device=dev)
            img = torch.empty((lib.sr_image_bytes(W, H),), dtype=torch.uint8, device=dev)
            nr_dev = torch.empty((2,), dtype=torch.int32, device=dev)
            stream = torch.cuda.current_stream(dev)
            key = (dev.index, W, H)
            cap = _pick_capacity(key, P)
            while True:
                fr = _capi.SrFrame(P, int(degree), M, W, H, _as_float(tan_fovx), _as_float(tan_fovy),
                                   float(scale_modifier), int(bool(prefiltered)), int(bool(debug)),
                                   _capi.SR_FLAG_LOCAL_SORT if (_local_sort_default and key not in _sort_global) else 0)
                binning = torch.empty((lib.sr_binning_bytes(cap, W, H),), dtype=torch.uint8, device=dev)
                nosync = (not _sync_mode) and key in _cap_hint
                host = _host_slot() if nosync else torch.empty((2,), dtype=torch.int32, pin_memory=True)
                rc = lib.sr_forward(
                    C.byref(fr), _ptr(background), _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity),
                    _ptr(scales), _ptr(rotations), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                    out_color.data_ptr(), out_others.data_ptr(), _ptr(radii), geom.data_ptr(),
                    binning.data_ptr(), img.data_ptr(), cap, nr_dev.data_ptr(), host.data_ptr(),
                    stream.cuda_stream)
                _capi.check(rc, "sr_forward")
                if nosync:
                    _pending.append((host, key, cap))
                    num_rendered = -1
                    break
                stream.synchronize()
                num_rendered, status = int(host[0]), int(host[1])
                _cap_hint[key] = max(_cap_hint.get(key, 0), num_rendered) if not _sync_mode else num_rendered
                if status & _capi.SR_STATUS_PREFILTER:
                    raise RuntimeError("Point is filtered although prefiltered is set. This shouldn't happen!")
                if status & _capi.SR_STATUS_SORT_CAP:
                    _sort_global.add(key)          # some tile is too long for the shared-memory sort: global path
                    continue
                if status & _capi.SR_STATUS_OVERFLOW:
                    cap = _round_cap(int(num_rendered * 1.1) + 4096)
                    continue
                break
        return num_rendered, out_color, out_others, radii, geom, binning, img

    @staticmethod
    def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                     transMat_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                     dL_dout_color, dL_dout_others, sh, degree, campos, geomBuffer, R,
                                     binningBuffer, imageBuffer, debug):
        lib = _capi.load()
        dev = means3D.device
        P = int(means3D.shape[0])
        H, W = int(dL_dout_color.shape[1]), int(dL_dout_color.shape[2])
        background = _f32c(background, "background"); means3D = _f32c(means3D, "means3D")
        colors = _f32c(colors, "colors"); scales = _f32c(scales, "scales"); rotations = _f32c(rotations, "rotations")
        viewmatrix = _f32c(viewmatrix, "viewmatrix"); projmatrix = _f32c(projmatrix, "projmatrix")
        sh = _f32c(sh, "sh"); campos = _f32c(campos, "campos")
        dL_dout_color = _f32c(dL_dout_color, "dL_dout_color"); dL_dout_others = _f32c(dL_dout_others, "dL_dout_others")
        M = int(sh.shape[1]) if sh is not None and sh.numel() != 0 else 0
        with torch.cuda.device(dev):
            mk = (lambda *s: torch.empty(s, dtype=torch.float32, device=dev)) if P > 0 else \
                 (lambda *s: torch.zeros(s, dtype=torch.float32, device=dev))
            dL_dmeans3D, dL_dmeans2D, dL_dcolors = mk(P, 3), mk(P, 3), mk(P, 3)
            dL_dopacity, dL_dtransMat, dL_dsh = mk(P, 1), mk(P, 9), mk(P, M, 3)
            dL_dscales, dL_drotations = mk(P, 2), mk(P, 4)
            if P > 0:
                cap = _capacity_from_bytes(int(binningBuffer.numel()), W, H)
                fr = _capi.SrFrame(P, int(degree), M, W, H, _as_float(tan_fovx), _as_float(tan_fovy),
                                   float(scale_modifier), 0, int(bool(debug)), 0)
                rc = lib.sr_backward(
                    C.byref(fr), _ptr(background), _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(scales),
                    _ptr(rotations), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), _ptr(radii),
                    _ptr(dL_dout_color), _ptr(dL_dout_others), geomBuffer.data_ptr(), binningBuffer.data_ptr(),
                    imageBuffer.data_ptr(), cap, dL_dmeans2D.data_ptr(), dL_dcolors.data_ptr(),
                    dL_dopacity.data_ptr(), dL_dmeans3D.data_ptr(), dL_dtransMat.data_ptr(),
                    dL_dsh.data_ptr() if M > 0 else None, dL_dscales.data_ptr(), dL_drotations.data_ptr(),
                    torch.cuda.current_stream(dev).cuda_stream)
                _capi.check(rc, "sr_backward")
        return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dscales, dL_drotations

    @staticmethod
    def mark_visible(means3D, viewmatrix, projmatrix):
        lib = _capi.load()
        dev = means3D.device
        P = int(means3D.shape[0])
        present = torch.zeros((P,), dtype=torch.bool, device=dev)
        if P:
            means3D = _f32c(means3D, "means3D"); viewmatrix = _f32c(viewmatrix, "viewmatrix")
            projmatrix = _f32c(projmatrix, "projmatrix")
            with torch.cuda.device(dev):
                rc = lib.sr_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix), present.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream)
            _capi.check(rc, "sr_mark_visible")
        return present


_C = _CNamespace()


def cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        args = (raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
                raster_settings.image_height, raster_settings.image_width, sh, raster_settings.sh_degree,
                raster_settings.campos, raster_settings.prefiltered, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        # only inputs + private buffers are saved (never our own outputs): the caller edits `color` in place
        # before backward (deformable_gaussian.py:188-190)
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        num_rendered = ctx.num_rendered
        raster_settings = ctx.raster_settings
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer = ctx.saved_tensors
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, raster_settings.image_height, raster_settings.image_width),
                                         dtype=torch.float32, device=means3D.device)
        if grad_depth is None:
            grad_depth = torch.zeros((8, raster_settings.image_height, raster_settings.image_width),
                                     dtype=torch.float32, device=means3D.device)
        args = (raster_settings.bg, means3D, radii, colors_precomp, scales, rotations, raster_settings.scale_modifier,
                cov3Ds_precomp, raster_settings.viewmatrix, raster_settings.projmatrix, raster_settings.tanfovx,
                raster_settings.tanfovy, grad_out_color, grad_depth, sh, raster_settings.sh_degree,
                raster_settings.campos, geomBuffer, num_rendered, binningBuffer, imgBuffer, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                grads8 = _C.rasterize_gaussians_backward(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            grads8 = _C.rasterize_gaussians_backward(*args)
        grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales, grad_rotations = grads8
        # the reference returns a gradient for every input slot, including the "not provided" empty tensors
        if colors_precomp.numel() == 0:
            grad_colors_precomp = None
        if sh.numel() == 0:
            grad_sh = None
        if cov3Ds_precomp.numel() == 0:
            grad_cov3Ds_precomp = None
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                grad_rotations, grad_cov3Ds_precomp, None)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        dev = means3D.device
        empty = lambda: torch.empty((0,), dtype=torch.float32, device=dev)  # noqa: E731
        if shs is None:
            shs = empty()
        if colors_precomp is None:
            colors_precomp = empty()
        if scales is None:
            scales = empty()
        if rotations is None:
            rotations = empty()
        if cov3D_precomp is None:
            cov3D_precomp = empty()
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings)
