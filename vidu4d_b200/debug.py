"""Views into the library's opaque buffers (sr_debug_view) -- used by the parity tests to compare
(tile|depth) keys, the sorted surfel list and the tile ranges with the reference, bit for bit."""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi
from .rasterizer import _capacity_from_bytes


def decode(geomBuffer, binningBuffer, imgBuffer, P: int, W: int, H: int, num_rendered: int) -> dict:
    lib = _capi.load()
    cap = _capacity_from_bytes(int(binningBuffer.numel()), W, H)
    L = _capi.SrDebugLayout()
    _capi.check(lib.sr_debug_view(P, W, H, cap, C.byref(L)), "sr_debug_view")
    N, R = W * H, int(num_rendered)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)

    def view(buf, off, count, dtype):
        sz = torch.empty((), dtype=dtype).element_size()
        return buf[off:off + count * sz].view(dtype)

    ctl = view(binningBuffer, L.sort_ctl, 64, torch.int32)
    sel = int(ctl[0].item())
    out = {
        "surfel_rec": view(geomBuffer, L.surfel_rec, P * 20, torch.float32).view(P, 20),
        "depths": view(geomBuffer, L.depths, P, torch.float32),
        "tiles_touched": view(geomBuffer, L.tiles_touched, P, torch.int32),
        "point_offsets": view(geomBuffer, L.point_offsets, P, torch.int32),
        "clamped": view(geomBuffer, L.clamped, P, torch.uint8),
        "keys_unsorted": view(binningBuffer, L.keys[0], R, torch.int64),   # valid only if no pass wrote back into ping
        "keys": view(binningBuffer, L.keys[sel], R, torch.int64),
        "point_list": view(binningBuffer, L.values[sel], R, torch.int32),
        "inst_rec": view(binningBuffer, L.inst_rec, R * 20, torch.float32).view(R, 20),
        "sort_ctl": ctl,
        "final_T": view(imgBuffer, L.final_T, 3 * N, torch.float32).view(3, H, W),
        "n_contrib": view(imgBuffer, L.n_contrib, 2 * N, torch.int32).view(2, H, W),
        "ranges": view(imgBuffer, L.ranges, 2 * tiles, torch.int32).view(tiles, 2),
    }
    return out
