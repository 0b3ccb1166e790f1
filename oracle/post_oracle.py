"""TEST INFRASTRUCTURE -- float64 NumPy oracle of the reference's render() post-processing and Stage-3 image losses.

Only tests/, __graft_entry__.smoke() and bench.py's checker legs may import this module; nothing under
vidu4d_b200/ does (tests/test_host_cpu.py asserts it).

Restates, in float64 NumPy with hand-written vector-Jacobian products (pinned to finite differences by
tests/test_post_oracle.py, so it does not depend on autograd or on any product code):

  post_forward / post_backward
      gs/gaussian_renderer/__init__.py:121-162   alpha, rotated normal, nan_to_num'd median / expected depth,
                                                 distortion, surf_depth mix, surf_normal * alpha.detach()
      gs/utils/point_utils.py:9-37               depths_to_points, depth_to_normal
  stage3_losses / stage3_losses_backward
      lab4d/engine/model.py:674-692              masked L1 on rgb where vis2d > 0 (mean over ALL elements)
      lab4d/engine/model.py:649-653              (rendered mask - gt mask)^2 * balance weight  ("fg" field type)
      lab4d/engine/model.py:817-842              normal loss lambda_n * mean(1 - <rend_normal, surf_normal>),
                                                 distortion loss lambda_d * mean(rend_dist)

Conventions: allmap is (8,H,W) with the reference's channel order (auxiliary.h:25-30): 0 depth, 1 alpha,
2..4 normal, 5 median depth, 6 distortion, 7 median weight.  world_view_transform is W2C transposed.
"""
from __future__ import annotations

import numpy as np

F32_LOWEST = float(np.finfo(np.float32).min)


def _nan_to_num00(x):
    """torch.nan_to_num(x, 0, 0) on a float32 tensor: nan -> 0, +inf -> 0, -inf -> float32 lowest."""
    y = np.array(x, dtype=np.float64, copy=True)
    y[np.isnan(x)] = 0.0
    y[np.isposinf(x)] = 0.0
    y[np.isneginf(x)] = F32_LOWEST
    return y


def rays(wvt, W, H, tanx, tany):
    """rays_d (H,W,3), rays_o (3,) of depths_to_points (point_utils.py:9-21): integer pixel grid, K^-1, c2w."""
    wvt = np.asarray(wvt, np.float64)
    c2w = np.linalg.inv(wvt.T)
    fx, fy = W / (2.0 * tanx), H / (2.0 * tany)
    K = np.array([[fx, 0.0, W / 2.0], [0.0, fy, H / 2.0], [0.0, 0.0, 1.0]])
    gx, gy = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64), indexing="xy")
    pts = np.stack([gx, gy, np.ones_like(gx)], -1).reshape(-1, 3)
    d = pts @ np.linalg.inv(K).T @ c2w[:3, :3].T
    return d.reshape(H, W, 3), c2w[:3, 3].copy()


def _normal_from_depth(sd, rays_d, rays_o):
    """depth_to_normal (point_utils.py:23-37).  Returns (normal (H,W,3), cache for the vjp)."""
    H, W = sd.shape
    pts = sd[..., None] * rays_d + rays_o
    dx = pts[2:, 1:-1] - pts[:-2, 1:-1]
    dy = pts[1:-1, 2:] - pts[1:-1, :-2]
    n = np.cross(dx, dy)
    ln = np.linalg.norm(n, axis=-1, keepdims=True)
    den = np.maximum(ln, 1e-12)                     # F.normalize eps
    out = np.zeros((H, W, 3))
    out[1:-1, 1:-1] = n / den
    return out, (dx, dy, n, ln, den)


def post_forward(allmap, wvt, tanx, tany, depth_ratio=0.0):
    """allmap (8,H,W) -> dict of the seven post-processed maps, float64 (single-plane depths, not tiled x3)."""
    a = np.asarray(allmap)
    _, H, W = a.shape
    a64 = a.astype(np.float64)
    wvt = np.asarray(wvt, np.float64)
    alpha = a64[1]
    M = wvt[:3, :3].T                                # rend_normal = n (row vector) @ world_view_transform[:3,:3].T
    rn = np.einsum("chw,cd->dhw", a64[2:5], M)
    m32 = a[5].astype(np.float32)
    med = np.where(np.isfinite(m32), a64[5], _nan_to_num00(m32))        # finite entries carried in float64
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (a[0].astype(np.float32) / a[1].astype(np.float32))   # the reference divides in float32: inf/nan pattern
        q64 = a64[0] / a64[1]
    ex = _nan_to_num00(q)
    fin = np.isfinite(q)
    ex = np.where(fin, q64, ex)                      # finite entries carried in float64
    sd = ex * (1.0 - depth_ratio) + depth_ratio * med
    rd, ro = rays(wvt, W, H, tanx, tany)
    nrm, _ = _normal_from_depth(sd, rd, ro)
    sn = np.transpose(nrm, (2, 0, 1)) * alpha[None]
    return {"acc": alpha[None], "rend_normal": rn, "rend_dist": a64[6][None], "render_depth_median": med[None],
            "render_depth_expected": ex[None], "surf_depth": sd[None], "surf_normal": sn}


def post_backward(allmap, wvt, tanx, tany, depth_ratio, grads):
    """Vector-Jacobian product of post_forward.  grads: dict with any of the post_forward keys -> upstream gradients
    of the same shapes.  Returns dL/dallmap (8,H,W) float64.  alpha is detached inside surf_normal (render():144)."""
    a = np.asarray(allmap)
    _, H, W = a.shape
    a64 = a.astype(np.float64)
    wvt = np.asarray(wvt, np.float64)
    g = {k: np.asarray(v, np.float64) for k, v in grads.items()}
    z1, z3 = np.zeros((1, H, W)), np.zeros((3, H, W))
    g_acc, g_rn, g_dist = g.get("acc", z1)[0], g.get("rend_normal", z3), g.get("rend_dist", z1)[0]
    g_med, g_ex = g.get("render_depth_median", z1)[0].copy(), g.get("render_depth_expected", z1)[0].copy()
    g_sd, g_sn = g.get("surf_depth", z1)[0].copy(), g.get("surf_normal", z3)
    out = np.zeros((8, H, W))
    alpha = a64[1]
    # ---- surf_normal = normalize(cross(dx, dy)) * alpha.detach()
    fw = post_forward(allmap, wvt, tanx, tany, depth_ratio)
    sd = fw["surf_depth"][0]
    rd, ro = rays(wvt, W, H, tanx, tany)
    _, (dx, dy, n, ln, den) = _normal_from_depth(sd, rd, ro)
    gn_full = np.transpose(g_sn, (1, 2, 0)) * alpha[..., None]          # gradient w.r.t. the (H,W,3) unit normal map
    gu = gn_full[1:-1, 1:-1]
    u = n / den
    big = ln > 1e-12
    gn = np.where(big, (gu - u * (u * gu).sum(-1, keepdims=True)) / den, gu / 1e-12)
    gdx = np.cross(dy, gn)
    gdy = np.cross(gn, dx)
    gp = np.zeros((H, W, 3))
    gp[2:, 1:-1] += gdx
    gp[:-2, 1:-1] -= gdx
    gp[1:-1, 2:] += gdy
    gp[1:-1, :-2] -= gdy
    g_sd = g_sd + (gp * rd).sum(-1)
    # ---- surf_depth = expected (1 - r) + r median
    g_ex += g_sd * (1.0 - depth_ratio)
    g_med += g_sd * depth_ratio
    with np.errstate(divide="ignore", invalid="ignore"):
        q = a[0].astype(np.float32) / a[1].astype(np.float32)
        fin = np.isfinite(q)
        out[0] = np.where(fin, g_ex / alpha, 0.0)
        out[1] = g_acc + np.where(fin, -g_ex * a64[0] / (alpha * alpha), 0.0)
    out[2:5] = np.einsum("dhw,cd->chw", g_rn, wvt[:3, :3].T)
    out[5] = np.where(np.isfinite(a[5]), g_med, 0.0)
    out[6] = g_dist
    return out


# ------------------------------------------------------------------------------------------------ Stage-3 losses
def mask_balance_wt(mask, vis2d):
    """dvr_model.get_mask_balance_wt (lab4d/engine/model.py:597-611) for one detected frame: scalar 1 or (H,W)."""
    mask = np.asarray(mask, np.float64)
    vis = np.asarray(vis2d, np.float64)
    if mask.sum() > 0 and (1 - mask).sum() > 0:
        pos = vis.sum() / mask[vis > 0].sum()
        neg = vis.sum() / (1 - mask[vis > 0]).sum()
        return 0.5 * pos * mask + 0.5 * neg * (1 - mask)
    return np.ones_like(mask)


def stage3_losses(color, allmap, wvt, tanx, tany, depth_ratio, target_rgb, vis2d, mask_gt, mask_wt,
                  w_rgb=1.0, w_mask=1.0, lambda_normal=0.05, lambda_dist=0.01, bkgd=None):
    """The image-space losses of one Stage-3 frame (float64).  Returns (total, dict of the four terms).
    color (3,H,W) rendered image, target_rgb (3,H,W), vis2d / mask_gt / mask_wt (H,W); bkgd (3,) = the learnable
    background composited under the render first (lab4d/nnutils/deformable_gaussian.py:188-190)."""
    c = np.asarray(color, np.float64)
    fw = post_forward(allmap, wvt, tanx, tany, depth_ratio)
    if bkgd is not None:
        c = c + (1.0 - fw["acc"]) * np.asarray(bkgd, np.float64)[:, None, None]
    vis = (np.asarray(vis2d) > 0)
    l1 = (np.abs(c - np.asarray(target_rgb, np.float64)) * vis[None]).mean()          # zeros where vis2d == 0, mean over all
    lmask = (((fw["acc"][0] - np.asarray(mask_gt, np.float64)) ** 2) * np.asarray(mask_wt, np.float64)).mean()
    lnorm = (1.0 - (fw["rend_normal"] * fw["surf_normal"]).sum(0)).mean()
    ldist = fw["rend_dist"].mean()
    terms = {"rgb": w_rgb * l1, "mask": w_mask * lmask, "normal": lambda_normal * lnorm, "dist": lambda_dist * ldist}
    return sum(terms.values()), terms


def stage3_losses_backward(color, allmap, wvt, tanx, tany, depth_ratio, target_rgb, vis2d, mask_gt, mask_wt,
                           w_rgb=1.0, w_mask=1.0, lambda_normal=0.05, lambda_dist=0.01, bkgd=None):
    """d total / d color (3,H,W), d total / d allmap (8,H,W) and (if bkgd is given) d total / d bkgd (3,), float64."""
    c = np.asarray(color, np.float64)
    _, H, W = c.shape
    N = H * W
    fw = post_forward(allmap, wvt, tanx, tany, depth_ratio)
    vis = (np.asarray(vis2d) > 0)
    bk = None if bkgd is None else np.asarray(bkgd, np.float64)
    if bk is not None:
        c = c + (1.0 - fw["acc"]) * bk[:, None, None]
    g_color = w_rgb * np.sign(c - np.asarray(target_rgb, np.float64)) * vis[None] / (3.0 * N)
    g_acc_bk = 0.0 if bk is None else -(g_color * bk[:, None, None]).sum(0)
    grads = {
        "acc": (w_mask * 2.0 * (fw["acc"][0] - np.asarray(mask_gt, np.float64)) * np.asarray(mask_wt, np.float64) / N + g_acc_bk)[None],
        "rend_normal": -lambda_normal * fw["surf_normal"] / N,
        "surf_normal": -lambda_normal * fw["rend_normal"] / N,
        "rend_dist": np.full((1, H, W), lambda_dist / N),
    }
    g_allmap = post_backward(allmap, wvt, tanx, tany, depth_ratio, grads)
    if bk is None:
        return g_color, g_allmap
    return g_color, g_allmap, (g_color * (1.0 - fw["acc"])).sum((1, 2))
