"""CPU tests pinning oracle/post_oracle.py (float64 NumPy restatement of the reference's render() post-processing and
Stage-3 image losses):
  (1) against fixtures produced by the REFERENCE's own code (tests/golden/make_post_golden.py executes the source text
      of gs/utils/point_utils.py and of render() lines 121-145 from /root/reference on the CPU);
  (2) its hand-written vector-Jacobian products against central finite differences in float64.
"""
import os

import numpy as np
import pytest

from .conftest import GOLDEN_DIR
from oracle import post_oracle as po

POST_CASES = ["post_id_48x40", "post_rigid_56x36_r03"]
KEYS = ("acc", "rend_normal", "rend_dist", "render_depth_median", "render_depth_expected", "surf_depth", "surf_normal")


def _load(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: g[k] for k in g.files}


@pytest.mark.parametrize("name", POST_CASES)
def test_post_oracle_matches_reference_outputs(name):
    g = _load(name)
    tanx, tany = [float(v) for v in g["in_tan"]]
    fw = po.post_forward(g["in_allmap"], g["in_wvt"], tanx, tany, float(g["in_depth_ratio"][0]))
    for k in KEYS:
        ref = g["ref_" + k].astype(np.float64)
        # the reference computes in float32; surf_normal normalises a cross product of central differences, which
        # amplifies float32 rounding where the surface is nearly flat in camera space
        tol = 2e-4 if k == "surf_normal" else 2e-6
        assert np.abs(fw[k] - ref).max() <= tol * max(1.0, np.abs(ref).max()), k


@pytest.mark.parametrize("name", POST_CASES)
def test_post_oracle_vjp_matches_reference_autograd(name):
    g = _load(name)
    tanx, tany = [float(v) for v in g["in_tan"]]
    grads = {k: g["w_" + k] for k in KEYS}
    ga = po.post_backward(g["in_allmap"], g["in_wvt"], tanx, tany, float(g["in_depth_ratio"][0]), grads)
    ref = g["ref_grad_allmap"].astype(np.float64)
    # autograd of 0/0 under nan_to_num yields NaN in the reference (div backward of a masked-out 0 gradient): the
    # rasterizer backward never reads those pixels' depth gradient (no contributor there); compare where finite
    fin = np.isfinite(ref)
    assert fin.mean() > 0.5
    scale = np.abs(ref[fin]).max()
    assert np.abs(ga[fin] - ref[fin]).max() <= 3e-4 * scale
    # the oracle (like the CUDA kernel) passes no gradient through the non-finite quotient: plane 0 gets 0, plane 1 only
    # the direct `acc` term
    assert not fin[0].all() and fin[2:].all()
    assert np.abs(ga[0][~fin[0]]).max(initial=0.0) == 0.0
    np.testing.assert_array_equal(ga[1][~fin[1]], grads["acc"][0].astype(np.float64)[~fin[1]])


def test_post_oracle_vjp_matches_finite_differences():
    g = _load("post_rigid_56x36_r03")
    tanx, tany = [float(v) for v in g["in_tan"]]
    rng = np.random.default_rng(0)
    am = g["in_allmap"].astype(np.float64)
    wts = {k: g["w_" + k].astype(np.float64) for k in KEYS}

    def f(a):
        fw = po.post_forward(a, g["in_wvt"], tanx, tany, 0.3)
        # alpha is DETACHED inside surf_normal (render():144): hold that factor at the unperturbed alpha
        with np.errstate(divide="ignore", invalid="ignore"):
            fw["surf_normal"] = np.where(a[1] > 0, fw["surf_normal"] / a[1] * am[1], 0.0)
        return sum((fw[k] * wts[k]).sum() for k in KEYS)
    ga = po.post_backward(am, g["in_wvt"], tanx, tany, 0.3, wts)
    cov = np.argwhere(am[1] > 0.2)
    for _ in range(40):
        y, x = cov[rng.integers(len(cov))]
        c = int(rng.integers(0, 7))
        h = 1e-6
        ap, an = am.copy(), am.copy()
        ap[c, y, x] += h; an[c, y, x] -= h
        fd = (f(ap) - f(an)) / (2 * h)
        assert abs(fd - ga[c, y, x]) <= 1e-5 * max(1.0, abs(fd)), (c, y, x, fd, ga[c, y, x])


def test_stage3_loss_oracle_gradients_match_finite_differences():
    g = _load("post_id_48x40")
    tanx, tany = [float(v) for v in g["in_tan"]]
    rng = np.random.default_rng(3)
    am = g["in_allmap"].astype(np.float64)
    H, W = am.shape[1:]
    color = rng.uniform(0, 1, size=(3, H, W))
    target = rng.uniform(0, 1, size=(3, H, W))
    vis = (rng.uniform(size=(H, W)) > 0.2).astype(np.float64)
    mask = (am[1] > 0.5).astype(np.float64)
    wt = po.mask_balance_wt(mask, vis)
    args = (g["in_wvt"], tanx, tany, 0.0, target, vis, mask, wt)
    kw = dict(w_rgb=0.8, w_mask=0.1, lambda_normal=0.05, lambda_dist=100.0)
    total, terms = po.stage3_losses(color, am, *args, **kw)
    assert np.isfinite(total) and set(terms) == {"rgb", "mask", "normal", "dist"}
    gc, ga = po.stage3_losses_backward(color, am, *args, **kw)
    cov = np.argwhere(am[1] > 0.2)
    for _ in range(30):
        y, x = cov[rng.integers(len(cov))]
        c = int(rng.choice([0, 2, 3, 4, 5, 6]))    # plane 1 (alpha) is detached inside surf_normal: checked below
        h = 1e-6
        ap, an = am.copy(), am.copy()
        ap[c, y, x] += h; an[c, y, x] -= h
        fd = (po.stage3_losses(color, ap, *args, **kw)[0] - po.stage3_losses(color, an, *args, **kw)[0]) / (2 * h)
        assert abs(fd - ga[c, y, x]) <= 1e-6 + 1e-4 * abs(fd), (c, y, x, fd, ga[c, y, x])
    kw0 = dict(kw, lambda_normal=0.0)              # without the normal term alpha is not detached anywhere
    ga0 = po.stage3_losses_backward(color, am, *args, **kw0)[1]
    for _ in range(10):
        y, x = cov[rng.integers(len(cov))]
        h = 1e-6
        ap, an = am.copy(), am.copy()
        ap[1, y, x] += h; an[1, y, x] -= h
        fd = (po.stage3_losses(color, ap, *args, **kw0)[0] - po.stage3_losses(color, an, *args, **kw0)[0]) / (2 * h)
        assert abs(fd - ga0[1, y, x]) <= 1e-6 + 1e-4 * abs(fd)
    for _ in range(10):
        y, x = cov[rng.integers(len(cov))]
        c = int(rng.integers(0, 3))
        h = 1e-7
        cp_, cn = color.copy(), color.copy()
        cp_[c, y, x] += h; cn[c, y, x] -= h
        fd = (po.stage3_losses(cp_, am, *args, **kw)[0] - po.stage3_losses(cn, am, *args, **kw)[0]) / (2 * h)
        assert abs(fd - gc[c, y, x]) <= 1e-9 + 1e-4 * abs(fd)
