"""`diff_surfel_rasterization._C` -- the importable stand-in for the reference's pybind module (RAST/ext.cpp:15-19,
RAST/setup.py:18-23): the same three entry points with the same signatures and return tuples
(RAST/rasterize_points.h:18-68), implemented over the C ABI of libsurfel_raster.so (vidu4d_b200/rasterizer.py)."""
from vidu4d_b200.rasterizer import _C as _impl

rasterize_gaussians = _impl.rasterize_gaussians
rasterize_gaussians_backward = _impl.rasterize_gaussians_backward
mark_visible = _impl.mark_visible

__all__ = ["rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"]
