"""Debug helper: render_loss_batch vs the per-frame render_fused + torch losses path at the headline size, per loss term."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from vidu4d_b200.renderer import BatchCameras, MiniCam, PipelineParams, render_fused, render_loss_batch
from vidu4d_b200.synthetic import SurfelCloud, object_scene, orbit_view, projection_matrix
dev = torch.device("cuda:0")
P, RES, F = int(sys.argv[1]) if len(sys.argv) > 1 else 300000, int(sys.argv[2]) if len(sys.argv) > 2 else 512, 4
scene = object_scene(P, seed=0, opacity="trained", center=(0, 0, 0))
Pm = projection_matrix(0.5, 0.5).astype(np.float64)
vms, pms, cps = [], [], []
for f in range(F):
    R, t = orbit_view(f, 64); W2C = np.eye(4); W2C[:3, :3] = R; W2C[:3, 3] = t
    vms.append(W2C.T.astype(np.float32)); pms.append((W2C.T @ Pm).astype(np.float32)); cps.append((-R.T @ t).astype(np.float32))
cam = torch.from_numpy(np.stack([np.stack(vms), np.stack(pms)], 1)).to(dev); cp = torch.from_numpy(np.stack(cps)).to(dev)
tg = torch.rand((F, 3, RES, RES), generator=torch.Generator().manual_seed(5)).to(dev)
fov = 2 * float(np.arctan(0.5)); bg = torch.zeros(3, device=dev); pipe = PipelineParams()
for wr, ln, ld in ((1.0, 0.0, 0.0), (0.0, 0.05, 0.0), (0.0, 0.0, 0.01), (1.0, 0.05, 0.01)):
    c1 = SurfelCloud(scene, dev)
    out = render_loss_batch(BatchCameras(RES, RES, fov, fov, cam[:, 0], cam[:, 1], cp), c1, pipe, bg, tg, w_rgb=wr, lambda_normal=ln, lambda_dist=ld)
    out["loss"].backward()
    c2 = SurfelCloud(scene, dev)
    tot = 0.0
    for f in range(F):
        o = render_fused(MiniCam(RES, RES, fov, fov, 0.01, 100.0, cam[f, 0], cam[f, 1], cp[f]), c2, pipe, bg)
        loss = wr * (o["render"] - tg[f]).abs().mean() + ln * (1.0 - (o["rend_normal"] * o["surf_normal"]).sum(0)).mean() + ld * o["rend_dist"].mean()
        loss.backward(); tot += float(loss)
    print(f"w_rgb={wr} l_n={ln} l_d={ld}: loss fused {float(out['loss']):.6f} eager {tot:.6f}")
    for n, a, b in zip(("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"), c1.flat_params(), c2.flat_params()):
        d = (a.grad - b.grad).double().norm() / (b.grad.double().norm() + 1e-30)
        print(f"    {n:9s} rel l2 diff {float(d):.3e}   |g| {float(b.grad.abs().max()):.3e}")
