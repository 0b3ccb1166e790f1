"""Generate tests/golden/post_*.npz from the REFERENCE's own post-processing code, run on the CPU in this container.

The reference's render() post-processing (gs/gaussian_renderer/__init__.py:121-162) sits inside render() behind the CUDA
rasterizer call, and gs/utils/point_utils.py hard-codes device='cuda'.  This script does not copy either: at run time
it reads the two files from /root/reference, takes (i) the source TEXT of depths_to_points / depth_to_normal and
(ii) lines 121-145 of render() (from `render_alpha = allmap[1:2]` to the surf_normal line), replaces the cuda device
strings by cpu, and exec()s them on a seeded `allmap` with torch autograd giving the gradients.  Inputs, outputs and
gradients are stored as small fixtures that pin oracle/post_oracle.py (tests/test_post_oracle.py).

    python tests/golden/make_post_golden.py        # needs /root/reference; writes tests/golden/post_*.npz
"""
import ast
import math
import os
import sys
import textwrap
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def _func_src(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    out = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            out.append(ast.get_source_segment(src, node))
    return "\n\n".join(out)


def reference_post():
    pu = _func_src(os.path.join(REF, "gs/utils/point_utils.py"), {"depths_to_points", "depth_to_normal"})
    pu = pu.replace("device='cuda'", "device='cpu'").replace(".float().cuda()", ".float()")
    ns = {"torch": torch, "math": math}
    exec(pu, ns)
    lines = open(os.path.join(REF, "gs/gaussian_renderer/__init__.py")).read().splitlines()
    a = next(i for i, l in enumerate(lines) if l.strip().startswith("render_alpha = allmap[1:2]"))
    b = next(i for i, l in enumerate(lines) if l.strip().startswith("surf_normal = surf_normal * (render_alpha).detach()"))
    body = textwrap.dedent("\n".join(lines[a:b + 1]))
    code = compile(body, "render_post_lines", "exec")

    def run(allmap, viewpoint_camera, pipe):
        loc = {"allmap": allmap, "viewpoint_camera": viewpoint_camera, "pipe": pipe, "torch": torch,
               "depth_to_normal": ns["depth_to_normal"]}
        exec(code, loc)
        return {"acc": loc["render_alpha"], "rend_normal": loc["render_normal"], "rend_dist": loc["render_dist"],
                "render_depth_median": loc["render_depth_median"], "render_depth_expected": loc["render_depth_expected"],
                "surf_depth": loc["surf_depth"], "surf_normal": loc["surf_normal"]}
    return run


def make_case(seed, W, H, depth_ratio, rigid):
    sys.path.insert(0, ROOT)
    from vidu4d_b200.synthetic import random_rotation
    rng = np.random.default_rng(seed)
    # a plausible allmap: a blob of coverage with alpha in (0,1), zero outside (alpha = 0 -> 0/0 in the expected depth)
    yy, xx = np.mgrid[0:H, 0:W]
    r2 = ((xx - W / 2) / (0.35 * W)) ** 2 + ((yy - H / 2) / (0.4 * H)) ** 2
    alpha = np.clip(1.2 - r2, 0.0, 0.97) * rng.uniform(0.7, 1.0, size=(H, W))
    alpha[r2 > 1.15] = 0.0
    depth = (1.0 + 0.3 * np.sin(xx / 7.0) * np.cos(yy / 5.0) + 0.02 * rng.normal(size=(H, W)))
    allmap = np.zeros((8, H, W), np.float32)
    allmap[0] = depth * alpha
    allmap[1] = alpha
    n = rng.normal(size=(3, H, W)); n /= np.linalg.norm(n, axis=0, keepdims=True)
    allmap[2:5] = n * alpha
    allmap[5] = depth * (alpha > 0) * (1 + 0.01 * rng.normal(size=(H, W)))
    allmap[6] = np.abs(rng.normal(size=(H, W))) * 0.01 * alpha
    allmap[7] = alpha * 0.5
    if rigid:
        R = random_rotation(rng); t = rng.normal(size=3) * 0.3
    else:
        R = np.eye(3); t = np.zeros(3)
    W2C = np.eye(4); W2C[:3, :3] = R; W2C[:3, 3] = t
    wvt = W2C.T.astype(np.float32)
    tanx, tany = 0.5, 0.5 * H / W
    cam = types.SimpleNamespace(world_view_transform=torch.from_numpy(wvt), image_width=W, image_height=H,
                                FoVx=2 * math.atan(tanx), FoVy=2 * math.atan(tany))
    pipe = types.SimpleNamespace(depth_ratio=depth_ratio)
    wts = {k: rng.normal(size=(c, H, W)).astype(np.float32) for k, c in
           (("acc", 1), ("rend_normal", 3), ("rend_dist", 1), ("render_depth_median", 1), ("render_depth_expected", 1),
            ("surf_depth", 1), ("surf_normal", 3))}
    run = reference_post()
    am = torch.from_numpy(allmap).requires_grad_(True)
    out = run(am, cam, pipe)
    loss = sum((out[k] * torch.from_numpy(w)).sum() for k, w in wts.items())
    loss.backward()
    res = {"in_allmap": allmap, "in_wvt": wvt, "in_tan": np.array([tanx, tany], np.float32),
           "in_depth_ratio": np.array([depth_ratio], np.float32), "ref_grad_allmap": am.grad.numpy()}
    for k, w in wts.items():
        res["w_" + k] = w
        res["ref_" + k] = out[k].detach().numpy()
    return res


CASES = {"post_id_48x40": dict(seed=1, W=48, H=40, depth_ratio=0.0, rigid=False),
         "post_rigid_56x36_r03": dict(seed=2, W=56, H=36, depth_ratio=0.3, rigid=True)}

if __name__ == "__main__":
    for name, kw in CASES.items():
        r = make_case(**kw)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **r)
        print(name, {k: v.shape for k, v in r.items() if k.startswith("ref_")})
