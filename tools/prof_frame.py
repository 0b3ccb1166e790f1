"""Minimal device-resident fwd+bwd loop of the headline frame for ncu captures (few launches, no torch glue)."""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from vidu4d_b200 import rasterizer as R
from vidu4d_b200.synthetic import object_scene, orbit_view, projection_matrix
ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=4); ap.add_argument("--surfels", type=int, default=300000)
ap.add_argument("--res", type=int, default=512); ap.add_argument("--opacity", default="trained"); ap.add_argument("--impl", default="ours")
a = ap.parse_args()
dev = torch.device("cuda:0")
sc = object_scene(a.surfels, seed=0, opacity=a.opacity, center=(0, 0, 0)); t = sc.to_torch(dev)
e = torch.empty((0,), device=dev); bg = torch.zeros(3, device=dev)
g = torch.Generator(device=dev).manual_seed(1234)
dLc = torch.randn((3, a.res, a.res), device=dev, generator=g); dLo = torch.randn((8, a.res, a.res), device=dev, generator=g) * 0.1
P = projection_matrix(0.5, 0.5).astype(np.float64)
if a.impl == "ours":
    C = R._C
else:
    from oracle import ref_ext; C = ref_ext.load()
for f in range(a.frames):
    Rm, tt = orbit_view(f, 64); W2C = np.eye(4); W2C[:3, :3] = Rm; W2C[:3, 3] = tt
    vm = torch.from_numpy(W2C.T.astype(np.float32)).to(dev); pm = torch.from_numpy((W2C.T @ P).astype(np.float32)).to(dev)
    cp = torch.from_numpy((-Rm.T @ tt).astype(np.float32)).to(dev)
    o = C.rasterize_gaussians(bg, t["means3D"], e, t["opacities"], t["scales"], t["rotations"], 1.0, e, vm, pm, 0.5, 0.5, a.res, a.res, t["shs"], 3, cp, False, False)
    C.rasterize_gaussians_backward(bg, t["means3D"], o[3], e, t["scales"], t["rotations"], 1.0, e, vm, pm, 0.5, 0.5, dLc, dLo, t["shs"], 3, cp, o[4], o[0], o[5], o[6], False)
torch.cuda.synchronize()
print("done R=", o[0])
