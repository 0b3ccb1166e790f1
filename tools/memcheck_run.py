"""Small forward+backward frames through the C ABI for `compute-sanitizer --tool memcheck` (no pytest, no oracle):
ragged image sizes, big surfels (many tiles per surfel), precomputed colours, every SH size, both sort paths."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from vidu4d_b200 import rasterizer as R
from vidu4d_b200.synthetic import object_scene, orbit_view, projection_matrix

dev = torch.device("cuda:0")
e = torch.empty((0,), device=dev)
P4 = projection_matrix(0.5, 0.5).astype(np.float64)


def run(P, W, H, M=16, deg=3, scale=1.0, precomp=False, view=0, seed=0):
    sc = object_scene(P, seed=seed, opacity="trained", center=(0, 0, 0)); t = sc.to_torch(dev)
    Rm, tt = orbit_view(view, 64); W2C = np.eye(4); W2C[:3, :3] = Rm; W2C[:3, 3] = tt
    vm = torch.from_numpy(W2C.T.astype(np.float32)).to(dev); pm = torch.from_numpy((W2C.T @ P4).astype(np.float32)).to(dev)
    cp = torch.from_numpy((-Rm.T @ tt).astype(np.float32)).to(dev)
    bg = torch.rand(3, device=dev)
    shs = t["shs"][:, :M].contiguous() if not precomp else e
    cols = torch.rand((P, 3), device=dev) if precomp else e
    scales = t["scales"] * scale
    o = R._C.rasterize_gaussians(bg, t["means3D"], cols, t["opacities"], scales, t["rotations"], 1.0, e, vm, pm, 0.5, 0.5, H, W,
                                 shs, deg, cp, False, False)
    dLc = torch.randn((3, H, W), device=dev); dLo = torch.randn((8, H, W), device=dev) * 0.1
    g = R._C.rasterize_gaussians_backward(bg, t["means3D"], o[3], cols, scales, t["rotations"], 1.0, e, vm, pm, 0.5, 0.5, dLc, dLo,
                                          shs, deg, cp, o[4], o[0], o[5], o[6], False)
    torch.cuda.synchronize()
    assert all(torch.isfinite(x).all() for x in g if x.numel())
    print(f"ok P={P} {W}x{H} M={M} deg={deg} scale={scale} precomp={precomp} R={int(o[0])}", flush=True)


run(5000, 80, 48, M=4, deg=1)
run(3000, 70, 50, precomp=True)
run(2000, 64, 64, scale=8.0)          # big surfels: dozens of tiles each, long lists
run(20000, 200, 120, view=7)
run(1, 32, 32)
for M, deg in ((1, 0), (9, 2), (16, 3)):
    run(4000, 96, 64, M=M, deg=deg, view=3)
R.set_sync_mode(False)
run(8000, 128, 128, view=11); run(8000, 128, 128, view=12)
R.check_overflow()
print("memcheck_run done")
