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
_sync_mode = True          # True: read num_rendered back after every forward (like the reference)
_pending: list = []        # nosync mode: (pinned host words (frames,2), key, capacity, stream) awaiting check_overflow()
_MAX_PENDING = 4096        # nosync forwards that may accumulate before check_overflow() is forced
_host_pool: list = []      # pinned (64,2) int32 status blocks, pre-allocated so that a forward never allocates pinned memory
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


_HOST_SLOT_FRAMES = 64      # frames one pinned status block can hold


def reserve_host_slots(n: int):
    """Pre-allocate pinned status words for the next `n` nosync forwards (needed before CUDA-graph capture)."""
    while len(_host_pool) - _host_next < n:
        _host_pool.append(torch.zeros((_HOST_SLOT_FRAMES, 2), dtype=torch.int32).pin_memory())


def _host_slot(frames: int = 1):
    """Pinned {num_rendered, status} words for one (batched) forward: (frames, 2) int32."""
    global _host_next
    if frames > _HOST_SLOT_FRAMES:
        return torch.zeros((frames, 2), dtype=torch.int32).pin_memory()
    if _host_next >= len(_host_pool):
        _host_pool.append(torch.zeros((_HOST_SLOT_FRAMES, 2), dtype=torch.int32).pin_memory())
    h = _host_pool[_host_next]
    _host_next += 1
    return h[:frames]


def check_overflow(keep: bool = False, sync: bool = True):
    """nosync mode: synchronise once, verify that no forward since the last call overflowed its instance
    buffer, and refresh the capacity hints.  Raises SurfelRasterError if a frame was dropped.
    keep=True leaves the watch list in place (a captured CUDA graph rewrites the same status words on every
    replay: call check_overflow(keep=True) after each replay).  sync=False: the caller has already waited for the
    forwards in question (e.g. on an event recorded behind them) -- used to check step k-1 while step k is in flight."""
    global _host_next
    bad, prefilter = None, False
    # Every device a pending forward ran on is synchronised (not just the current one): the asynchronous D2H copies of the
    # status words have landed.  Device-wide rather than per stream, because a forward recorded into a CUDA graph replays
    # on whatever stream the graph is launched on, not on the stream it was captured on.
    devices, bare = set(), []
    for _, _, _, stream in _pending:
        d = getattr(stream, "device", None)
        if d is not None:
            devices.add(d)
        elif stream not in bare:
            bare.append(stream)
    if sync:
        for d in devices:
            torch.cuda.synchronize(d)
        for stream in bare:
            stream.synchronize()
    for host, key, cap, stream in _pending:
        for r, status in host.tolist():
            _cap_hint[key] = max(_cap_hint.get(key, 0), r)
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
            f"instance buffer overflow in nosync mode: num_rendered={bad[0]}, "
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
    t = t.contiguous()
    if t.data_ptr() % 16 != 0 and t.numel() != 0:
        # the kernels use 128-bit loads (rotations, SH rows) and 64-bit loads (scales); a contiguous view at an odd
        # storage offset (e.g. `rot[1:]`, a slice of a flat parameter buffer) is re-homed; the reference's scalar glm
        # loads accept such tensors, so must we
        t = t.clone()
    return t


def _batch_inputs(M: int, **named):
    """Normalise the per-surfel inputs of a batched call.  Each value is (tensor, trailing shape or None); a tensor is
    either shared by the M frames (its single-frame shape) or per frame (leading M).  Returns (contiguous float32
    tensors, per-frame strides in floats: 0 = shared)."""
    out, strides = {}, {}
    for name, (t, tail) in named.items():
        if t is None or t.numel() == 0:
            out[name], strides[name] = None, 0
            continue
        t = _f32c(t, name)
        base_ndim = 3 if name == "sh" else 2
        if t.ndim == base_ndim + 1:
            if t.shape[0] != M:
                raise RuntimeError(f"{name}: leading dimension {t.shape[0]} != number of frames {M}")
            strides[name] = int(t[0].numel()) if M > 1 else 0
        elif t.ndim == base_ndim:
            strides[name] = 0
        else:
            raise RuntimeError(f"{name} must have {base_ndim} (shared) or {base_ndim + 1} (per frame) dimensions")
        if tail is not None and tuple(t.shape[-len(tail):]) != tail:
            raise RuntimeError(f"{name} must have trailing dimensions {tail}")
        out[name] = t
    return out, strides


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
                                   float(scale_modifier), int(bool(prefiltered)), int(bool(debug)), 0)
                binning = torch.empty((lib.sr_binning_bytes(cap, W, H),), dtype=torch.uint8, device=dev)
                nosync = (not _sync_mode) and key in _cap_hint
                host = _host_slot() if nosync else torch.empty((1, 2), dtype=torch.int32, pin_memory=True)
                rc = lib.sr_forward(
                    C.byref(fr), _ptr(background), _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity),
                    _ptr(scales), _ptr(rotations), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                    out_color.data_ptr(), out_others.data_ptr(), _ptr(radii), geom.data_ptr(),
                    binning.data_ptr(), img.data_ptr(), cap, nr_dev.data_ptr(), host.data_ptr(),
                    stream.cuda_stream)
                _capi.check(rc, "sr_forward")
                if nosync:
                    _pending.append((host, key, cap, stream))
                    if len(_pending) > _MAX_PENDING:
                        check_overflow()      # an unchecked backlog must not grow without bound (nor hide dropped frames)
                    num_rendered = -1
                    break
                stream.synchronize()
                num_rendered, status = int(host[0, 0]), int(host[0, 1])
                _cap_hint[key] = max(_cap_hint.get(key, 0), num_rendered) if not _sync_mode else num_rendered
                if status & _capi.SR_STATUS_PREFILTER:
                    raise RuntimeError("Point is filtered although prefiltered is set. This shouldn't happen!")
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

    # ---- batched entry points: M frames in one launch set -------------------------------------
    @staticmethod
    def rasterize_gaussians_batch(background, means3D, colors, opacity, scales, rotations, scale_modifier, viewmatrix,
                                  projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                                  prefiltered=False, debug=False):
        """M frames at once.  viewmatrix / projmatrix (M,4,4), campos (M,3).  Every per-surfel input is either shared by
        all frames -- the single-frame shape, e.g. means3D (P,3) -- or per frame with a leading M, e.g. (M,P,3) (Stage 3:
        each frame rasterizes its own warped copy of the surfels).  Returns (num_rendered (list of M ints, or -1 in nosync
        mode), out_color (M,3,H,W), out_others (M,8,H,W), radii (M,P), geomBuffer, binningBuffer, imgBuffer)."""
        lib = _capi.load()
        M = int(viewmatrix.shape[0])
        if viewmatrix.ndim != 3 or viewmatrix.shape[1:] != (4, 4) or campos.shape != (M, 3):
            raise RuntimeError("batched cameras: viewmatrix must be (M,4,4) and campos (M,3)")
        dev = means3D.device
        H, W = int(image_height), int(image_width)
        background = _f32c(background, "background"); viewmatrix = _f32c(viewmatrix, "viewmatrix")
        projmatrix = _f32c(projmatrix, "projmatrix"); campos = _f32c(campos, "campos")
        ins, strides = _batch_inputs(M, means3D=(means3D, (3,)), sh=(sh, None), colors=(colors, (3,)), opacity=(opacity, (1,)),
                                     scales=(scales, (2,)), rotations=(rotations, (4,)))
        means3D, sh, colors, opacity, scales, rotations = (ins[k] for k in ("means3D", "sh", "colors", "opacity", "scales", "rotations"))
        P = int(means3D.shape[-2])
        Msh = int(sh.shape[-2]) if sh is not None and sh.numel() != 0 else 0
        bt = _capi.SrBatch(M, 0, strides["means3D"], strides["sh"], strides["colors"], strides["opacity"], strides["scales"],
                           strides["rotations"])
        with torch.cuda.device(dev):
            out_color = torch.empty((M, 3, H, W), dtype=torch.float32, device=dev)
            out_others = torch.empty((M, 8, H, W), dtype=torch.float32, device=dev)
            radii = torch.empty((M, P), dtype=torch.int32, device=dev)
            geom = torch.empty((M * lib.sr_geom_bytes(P),), dtype=torch.uint8, device=dev)
            img = torch.empty((M * lib.sr_image_bytes(W, H),), dtype=torch.uint8, device=dev)
            nr_dev = torch.empty((M, 2), dtype=torch.int32, device=dev)
            stream = torch.cuda.current_stream(dev)
            key = (dev.index, W, H)
            cap = _pick_capacity(key, P)
            while True:
                fr = _capi.SrFrame(P, int(degree), Msh, W, H, _as_float(tan_fovx), _as_float(tan_fovy),
                                   float(scale_modifier), int(bool(prefiltered)), int(bool(debug)), 0)
                binning = torch.empty((M * lib.sr_binning_bytes(cap, W, H),), dtype=torch.uint8, device=dev)
                nosync = (not _sync_mode) and key in _cap_hint
                host = _host_slot(M) if nosync else torch.empty((M, 2), dtype=torch.int32, pin_memory=True)
                rc = lib.sr_forward_batch(
                    C.byref(fr), C.byref(bt), _ptr(background), _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity),
                    _ptr(scales), _ptr(rotations), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                    out_color.data_ptr(), out_others.data_ptr(), _ptr(radii), geom.data_ptr(),
                    binning.data_ptr(), img.data_ptr(), cap, nr_dev.data_ptr(), host.data_ptr(), stream.cuda_stream)
                _capi.check(rc, "sr_forward_batch")
                if nosync:
                    _pending.append((host, key, cap, stream))
                    num_rendered = -1
                    break
                stream.synchronize()
                num_rendered = [int(v) for v in host[:, 0]]
                status = 0
                for v in host[:, 1]:
                    status |= int(v)
                _cap_hint[key] = max(_cap_hint.get(key, 0), max(num_rendered)) if not _sync_mode else max(num_rendered)
                if status & _capi.SR_STATUS_PREFILTER:
                    raise RuntimeError("Point is filtered although prefiltered is set. This shouldn't happen!")
                if status & _capi.SR_STATUS_OVERFLOW:
                    cap = _round_cap(int(max(num_rendered) * 1.1) + 4096)
                    continue
                break
        return num_rendered, out_color, out_others, radii, geom, binning, img

    @staticmethod
    def rasterize_gaussians_backward_batch(background, means3D, radii, colors, scales, rotations, scale_modifier, viewmatrix,
                                           projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_others, sh, degree, campos,
                                           geomBuffer, binningBuffer, imageBuffer, grad_scale=None, debug=False,
                                           sum_shared=False, opacity_shared=True, want_transmat=True, outs=None):
        """Backward of rasterize_gaussians_batch.  dL_dout_color (M,3,H,W), dL_dout_others (M,8,H,W); grad_scale: optional
        0-d device tensor multiplying both (the upstream scalar of a fused loss).  Returns the eight gradient tensors
        (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dscales, dL_drotations), each with a
        leading M: (M,P,.).  With sum_shared=True the gradients of inputs the frames SHARE come back summed over the frames
        with the input's own shape (P,.) -- the per-surfel kernel accumulates them in place, nothing (M,P,.)-sized is
        written or reduced; `opacity_shared` says whether the forward's opacities were shared (they are not an input of
        the backward); dL_dmeans2D stays (M,P,3); want_transmat=False skips the (M,P,9) by-product.
        outs: optional {"dL_dmeans3D" | "dL_dsh" | "dL_dopacity" | "dL_dscales" | "dL_drotations" | "dL_dcolors": tensor}
        of pre-allocated, suitably shaped and aligned outputs (e.g. views of the flat buffer an all-reduce runs over)."""
        lib = _capi.load()
        M = int(viewmatrix.shape[0])
        dev = means3D.device
        H, W = int(dL_dout_color.shape[-2]), int(dL_dout_color.shape[-1])
        background = _f32c(background, "background"); viewmatrix = _f32c(viewmatrix, "viewmatrix")
        projmatrix = _f32c(projmatrix, "projmatrix"); campos = _f32c(campos, "campos")
        dL_dout_color = _f32c(dL_dout_color, "dL_dout_color"); dL_dout_others = _f32c(dL_dout_others, "dL_dout_others")
        ins, strides = _batch_inputs(M, means3D=(means3D, (3,)), sh=(sh, None), colors=(colors, (3,)),
                                     scales=(scales, (2,)), rotations=(rotations, (4,)))
        means3D, sh, colors, scales, rotations = (ins[k] for k in ("means3D", "sh", "colors", "scales", "rotations"))
        P = int(means3D.shape[-2])
        Msh = int(sh.shape[-2]) if sh is not None and sh.numel() != 0 else 0
        op_stride = 0 if (opacity_shared or M == 1) else P
        bt = _capi.SrBatch(M, _capi.SR_BATCH_SUM_SHARED if sum_shared else 0, strides["means3D"], strides["sh"], strides["colors"],
                           op_stride, strides["scales"], strides["rotations"])
        with torch.cuda.device(dev):
            mk0 = (lambda *s: torch.empty(s, dtype=torch.float32, device=dev)) if P > 0 else \
                  (lambda *s: torch.zeros(s, dtype=torch.float32, device=dev))

            def mk(name, shared, *tail):   # (M,P,.) per frame, or (P,.) when the input is shared and the kernel sums over frames
                shape = (P,) + tail if (sum_shared and shared and M > 1) else (M, P) + tail
                o = outs.get(name) if outs else None
                if o is not None:
                    if (o.numel() != int(torch.Size(shape).numel()) or not o.is_contiguous() or o.dtype != torch.float32
                            or o.data_ptr() % 16 != 0):
                        raise RuntimeError(f"outs[{name!r}] must be a contiguous, 16-byte aligned float32 tensor of {shape}")
                    return o.view(shape)
                return mk0(*shape)
            dL_dmeans2D = mk0(M, P, 3)
            dL_dmeans3D, dL_dcolors = mk("dL_dmeans3D", strides["means3D"] == 0, 3), mk("dL_dcolors", strides["colors"] == 0, 3)
            dL_dopacity = mk("dL_dopacity", op_stride == 0, 1)
            dL_dtransMat = mk0(M, P, 9) if (want_transmat or (M == 1 and not sum_shared)) else None
            dL_dsh = mk("dL_dsh", strides["sh"] == 0, Msh, 3)
            dL_dscales, dL_drotations = mk("dL_dscales", strides["scales"] == 0, 2), mk("dL_drotations", strides["rotations"] == 0, 4)
            if P > 0:
                cap = _capacity_from_bytes(int(binningBuffer.numel()) // M, W, H)
                fr = _capi.SrFrame(P, int(degree), Msh, W, H, _as_float(tan_fovx), _as_float(tan_fovy),
                                   float(scale_modifier), 0, int(bool(debug)), 0)
                gs = None
                if grad_scale is not None:
                    gs = grad_scale.to(device=dev, dtype=torch.float32).reshape(()).contiguous()
                rc = lib.sr_backward_batch(
                    C.byref(fr), C.byref(bt), _ptr(background), _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(scales),
                    _ptr(rotations), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), _ptr(radii),
                    _ptr(dL_dout_color), _ptr(dL_dout_others), gs.data_ptr() if gs is not None else None,
                    geomBuffer.data_ptr(), binningBuffer.data_ptr(), imageBuffer.data_ptr(), cap,
                    dL_dmeans2D.data_ptr(), dL_dcolors.data_ptr(), dL_dopacity.data_ptr(), dL_dmeans3D.data_ptr(),
                    dL_dtransMat.data_ptr() if dL_dtransMat is not None else None, dL_dsh.data_ptr() if Msh > 0 else None,
                    dL_dscales.data_ptr(), dL_drotations.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
                _capi.check(rc, "sr_backward_batch")
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


class BatchRasterizationSettings(NamedTuple):
    """GaussianRasterizationSettings for M frames: viewmatrix / projmatrix are (M,4,4), campos is (M,3)."""
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
    prefiltered: bool = False
    debug: bool = False


def _reduce_like(grad, inp):
    """(M,P,.) per-frame gradient -> the input's shape: inputs shared by all frames get the sum over frames."""
    if inp is None or inp.numel() == 0:
        return None
    if grad.ndim == inp.ndim:                 # per-frame input, or already summed over the frames by the kernel
        return grad.view(inp.shape)
    return grad.sum(dim=0).view(inp.shape)


class _RasterizeGaussiansBatch(torch.autograd.Function):
    """M frames in one launch set.  Inputs are shared (single-frame shape) or per frame (leading M);
    means2D is (M,P,3) and only receives the densification proxy gradient, as in the single-frame op."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, rs):
        num_rendered, color, allmap, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians_batch(
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, rs.viewmatrix, rs.projmatrix,
            rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        ctx.rs = rs
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, radii, sh, opacities, geomBuffer, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii, allmap

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_allmap):
        rs = ctx.rs
        colors_precomp, means3D, scales, rotations, radii, sh, opacities, geomBuffer, binningBuffer, imgBuffer = ctx.saved_tensors
        M = int(rs.viewmatrix.shape[0])
        H, W = rs.image_height, rs.image_width
        if grad_color is None:
            grad_color = torch.zeros((M, 3, H, W), dtype=torch.float32, device=means3D.device)
        if grad_allmap is None:
            grad_allmap = torch.zeros((M, 8, H, W), dtype=torch.float32, device=means3D.device)
        g2d, gcol, gop, g3d, gtm, gsh, gsc, grot = _C.rasterize_gaussians_backward_batch(
            rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, rs.viewmatrix, rs.projmatrix,
            rs.tanfovx, rs.tanfovy, grad_color, grad_allmap, sh, rs.sh_degree, rs.campos, geomBuffer, binningBuffer, imgBuffer,
            None, rs.debug, sum_shared=True, opacity_shared=opacities.ndim == 2, want_transmat=False)
        return (_reduce_like(g3d, means3D), g2d, _reduce_like(gsh, sh), _reduce_like(gcol, colors_precomp),
                _reduce_like(gop, opacities), _reduce_like(gsc, scales), _reduce_like(grot, rotations), None)


def rasterize_gaussians_batch(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, raster_settings):
    """Batched twin of rasterize_gaussians: returns (color (M,3,H,W), radii (M,P), allmap (M,8,H,W))."""
    dev = means3D.device
    empty = lambda: torch.empty((0,), dtype=torch.float32, device=dev)  # noqa: E731
    if (shs is None) == (colors_precomp is None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    return _RasterizeGaussiansBatch.apply(means3D, means2D, empty() if shs is None else shs,
                                          empty() if colors_precomp is None else colors_precomp, opacities, scales, rotations,
                                          raster_settings)


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
