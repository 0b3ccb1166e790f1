"""One batched step (F frames, forward + backward) of the headline workload for ncu captures: few launches, no torch glue
beyond the buffers.  `--warm 1` runs an untimed step first (pass `-s <launches>` to ncu to skip it)."""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from vidu4d_b200 import rasterizer as R
from vidu4d_b200.synthetic import object_scene, orbit_view, projection_matrix
ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=8); ap.add_argument("--surfels", type=int, default=300000)
ap.add_argument("--res", type=int, default=512); ap.add_argument("--steps", type=int, default=1); ap.add_argument("--warm", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
sc = object_scene(a.surfels, seed=0, opacity="trained", center=(0, 0, 0)); t = sc.to_torch(dev)
e = torch.empty((0,), device=dev); bg = torch.zeros(3, device=dev); F = a.frames
g = torch.Generator(device=dev).manual_seed(1234)
dLc = torch.randn((F, 3, a.res, a.res), device=dev, generator=g); dLo = torch.randn((F, 8, a.res, a.res), device=dev, generator=g) * 0.1
P = projection_matrix(0.5, 0.5).astype(np.float64)
vms, pms, cps = [], [], []
for f in range(F):
    Rm, tt = orbit_view(f, 64); W2C = np.eye(4); W2C[:3, :3] = Rm; W2C[:3, 3] = tt
    vms.append(W2C.T.astype(np.float32)); pms.append((W2C.T @ P).astype(np.float32)); cps.append((-Rm.T @ tt).astype(np.float32))
vms = torch.from_numpy(np.stack(vms)).to(dev); pms = torch.from_numpy(np.stack(pms)).to(dev); cps = torch.from_numpy(np.stack(cps)).to(dev)
C = R._C
for s in range(a.warm + a.steps):
    o = C.rasterize_gaussians_batch(bg, t["means3D"], e, t["opacities"], t["scales"], t["rotations"], 1.0, vms, pms, 0.5, 0.5, a.res, a.res, t["shs"], 3, cps)
    gr = C.rasterize_gaussians_backward_batch(bg, t["means3D"], o[3], e, t["scales"], t["rotations"], 1.0, vms, pms, 0.5, 0.5, dLc, dLo, t["shs"], 3, cps, o[4], o[5], o[6])
torch.cuda.synchronize()
print("done R=", o[0])
