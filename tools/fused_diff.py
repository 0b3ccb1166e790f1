import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from vidu4d_b200.renderer import PipelineParams, make_camera, render, render_fused
from vidu4d_b200.synthetic import SurfelCloud, object_scene
dev = torch.device("cuda:0")
cam = make_camera(160, 112, 2 * np.arctan(0.5), 2 * np.arctan(0.35), device=dev)
bg = torch.tensor([0.1, 0.3, 0.2], device=dev)
outs = []
for fn in (render, render_fused):
    cloud = SurfelCloud(object_scene(6000, seed=11), dev)
    outs.append(fn(cam, cloud, PipelineParams(depth_ratio=0.3), bg))
for k in outs[0]:
    a, b = outs[0][k], outs[1][k]
    if a.dtype != torch.float32: continue
    d = (a - b).abs()
    i = int(d.argmax())
    print(k, tuple(a.shape), "max|d|", float(d.max()), "at", np.unravel_index(i, a.shape), "a", float(a.reshape(-1)[i]), "b", float(b.reshape(-1)[i]), "n>1e-5:", int((d > 1e-5).sum()))
