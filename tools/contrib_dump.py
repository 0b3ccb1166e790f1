"""Dump the forward's per-pixel contribution masks of the headline frame (for offline scheduling studies):
gpurun_out/contrib_dump.npz with ranges (tiles,2), sub_last (tiles*8), masks ((R>>5)+tiles+1, 8, 32) uint32."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from vidu4d_b200 import rasterizer as R, _capi, debug
from vidu4d_b200.synthetic import object_scene, orbit_view, projection_matrix
P, res = 300000, 512
dev = torch.device("cuda:0")
sc = object_scene(P, seed=0, opacity="trained", center=(0, 0, 0)); t = sc.to_torch(dev)
e = torch.empty((0,), device=dev); bg = torch.zeros(3, device=dev)
Pm = projection_matrix(0.5, 0.5).astype(np.float64)
Rm, tt = orbit_view(0, 64); W2C = np.eye(4); W2C[:3, :3] = Rm; W2C[:3, 3] = tt
vm = torch.from_numpy(W2C.T.astype(np.float32)).to(dev); pm = torch.from_numpy((W2C.T @ Pm).astype(np.float32)).to(dev)
cp = torch.from_numpy((-Rm.T @ tt).astype(np.float32)).to(dev)
o = R._C.rasterize_gaussians(bg, t["means3D"], e, t["opacities"], t["scales"], t["rotations"], 1.0, e, vm, pm, 0.5, 0.5, res, res, t["shs"], 3, cp, False, False)
torch.cuda.synchronize()
Rn = int(o[0]); binb = o[5]; img = o[6]
cap = R._capacity_from_bytes(int(binb.numel()), res, res)
L = _capi.SrDebugLayout(); _capi.check(_capi.load().sr_debug_view(P, res, res, cap, C.byref(L)), "dbg")
tiles = (res // 16) ** 2
off = (L.inst_rec + cap * 80 + 255) // 256 * 256
nst = (Rn >> 5) + tiles + 1
masks = binb[off:off + nst * 1024].view(torch.int32).view(nst, 8, 32).cpu().numpy().astype(np.uint32)
d = debug.decode(o[4], binb, img, P, res, res, Rn)
ranges = d["ranges"].cpu().numpy()
n_contrib = d["n_contrib"].cpu().numpy()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "contrib_dump.npz"), ranges=ranges, masks=masks, n_contrib=n_contrib, R=Rn)
print("dumped", Rn, masks.shape, int(np.unpackbits(masks.view(np.uint8)).sum()))
