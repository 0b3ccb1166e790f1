"""GPU parity of the fused bob-skinning warp (csrc/warp.cu via vidu4d_b200.warp) against the float64 oracle
oracle/warp_oracle.py (pinned to the reference's own code by tests/test_warp_oracle.py) and against the reference
fixtures directly: outputs and every gradient within 1e-4 relative."""
import os

import numpy as np
import pytest
import torch

from .conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev(built):
    from vidu4d_b200 import _capi
    _capi.load()
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _q2dq(q, tr):
    from oracle import warp_oracle as wo
    t4 = torch.cat((torch.zeros_like(tr[..., :1]), tr), -1)
    return q, 0.5 * wo.qmul(t4, q)


@pytest.mark.parametrize("name", ["warp_b25_m3", "warp_b7_m2_nodelta"])
def test_warp_kernel_against_reference_fixture(name, dev):
    from vidu4d_b200.warp import bob_warp
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    t = {k[3:]: torch.tensor(v, device=dev, requires_grad=True) for k, v in g.items() if k.startswith("in_")}
    rest = _q2dq(t["rest_q"], t["rest_t"])
    art = _q2dq(t["art_q"], t["art_t"])
    xc, rc, ent = bob_warp(t["xyz"], t["rot"], rest, art, t["log_gauss"], (t["cam_q"], t["cam_t"]), t.get("delta"))
    assert _rel(xc.detach().cpu(), g["ref_xyz_cam"]) <= TOL and _rel(rc.detach().cpu(), g["ref_rot_cam"]) <= TOL
    assert _rel(ent.detach().cpu(), g["ref_entropy"][0]) <= TOL
    w = lambda k: torch.tensor(g[k], device=dev)  # noqa: E731
    ((xc * w("w_xyz")).sum() + (rc * w("w_rot")).sum() + (ent * w("w_ent")).sum()).backward()
    for k, v in t.items():
        ref = g.get("ref_grad_" + k)
        if ref is not None:
            assert _rel(v.grad.cpu(), ref) <= 3e-4, k       # the fixture itself is float32 autograd


@pytest.mark.parametrize("P,B,M,with_delta", [(30000, 25, 4, True), (5000, 40, 2, False), (777, 3, 9, True)])
def test_warp_kernel_against_float64_oracle(P, B, M, with_delta, dev):
    from oracle import warp_oracle as wo
    from tests.golden.make_warp_golden import make_inputs
    from vidu4d_b200.warp import bob_warp
    inp = make_inputs(seed=100 + B, P=P, B=B, M=M, with_delta=with_delta)
    # ---- oracle, float64 on the CPU
    t64 = {k: v.double().requires_grad_(True) for k, v in inp.items() if v is not None}
    rest = _q2dq(t64["rest_q"], t64["rest_t"]); art = _q2dq(t64["art_q"], t64["art_t"])
    M_ = t64["art_q"].shape[0]
    rest_m = (rest[0][None].expand(M_, -1, -1), rest[1][None].expand(M_, -1, -1))
    se3 = wo.dq_mul(art, wo.dq_inverse(rest_m))
    xo, ro, eo = wo.bob_warp(t64["xyz"], t64["rot"], rest, torch.exp(-t64["log_gauss"]), se3, (t64["cam_q"], t64["cam_t"]), t64.get("delta"))
    gen = torch.Generator().manual_seed(3)
    wx, wr, we = torch.randn(xo.shape, generator=gen), torch.randn(ro.shape, generator=gen), torch.randn(eo.shape, generator=gen)
    ((xo * wx.double()).sum() + (ro * wr.double()).sum() + (eo * we.double()).sum()).backward()
    # ---- kernel
    t = {k: v.to(dev).requires_grad_(True) for k, v in inp.items() if v is not None}
    xc, rc, ent = bob_warp(t["xyz"], t["rot"], _q2dq(t["rest_q"], t["rest_t"]), _q2dq(t["art_q"], t["art_t"]), t["log_gauss"],
                           (t["cam_q"], t["cam_t"]), t.get("delta"))
    assert _rel(xc.detach().cpu(), xo.detach()) <= TOL and _rel(rc.detach().cpu(), ro.detach()) <= TOL
    assert _rel(ent.detach().cpu(), eo.detach()) <= TOL
    ((xc * wx.to(dev)).sum() + (rc * wr.to(dev)).sum() + (ent * we.to(dev)).sum()).backward()
    for k, v in t.items():
        assert _rel(v.grad.cpu(), t64[k].grad) <= TOL, k


def test_warp_feeds_the_batched_rasterizer(dev):
    """The warp's (M,P,3) / (M,P,4) outputs are the per-frame inputs of the batched rasterizer: gradients of an image loss
    reach the canonical surfels and the bone tables through both kernels."""
    from tests.golden.make_warp_golden import make_inputs
    from vidu4d_b200 import rasterizer as RZ
    from vidu4d_b200.synthetic import SurfelCloud, object_scene
    from vidu4d_b200.warp import bob_warp
    P, B, M, W, H = 8000, 25, 2, 128, 96
    scene = object_scene(P, seed=4, center=(0.0, 0.0, 0.0))
    cloud = SurfelCloud(scene, dev)
    inp = make_inputs(seed=9, P=P, B=B, M=M, with_delta=False)
    t = {k: v.to(dev).requires_grad_(True) for k, v in inp.items() if v is not None and k not in ("xyz", "rot")}
    xc, rc, ent = bob_warp(cloud.get_xyz, cloud.get_rotation, _q2dq(t["rest_q"], t["rest_t"]), _q2dq(t["art_q"], t["art_t"]),
                           t["log_gauss"], (t["cam_q"], t["cam_t"]))
    eye = torch.eye(4, device=dev)[None].expand(M, -1, -1).contiguous()
    from vidu4d_b200.synthetic import projection_matrix
    pm = torch.from_numpy(projection_matrix(0.5, 0.375)).to(dev)[None].expand(M, -1, -1).contiguous()
    rs = RZ.BatchRasterizationSettings(H, W, 0.5, 0.375, torch.zeros(3, device=dev), 1.0, eye, pm, 3, torch.zeros((M, 3), device=dev))
    m2d = torch.zeros((M, P, 3), device=dev, requires_grad=True)
    color, radii, allmap = RZ.rasterize_gaussians_batch(xc, m2d, cloud.get_features, None, cloud.get_opacity, cloud.get_scaling,
                                                        torch.nn.functional.normalize(rc, dim=-1), rs)
    assert int((radii > 0).sum()) > P // 4
    (color.mean() + allmap[:, 1].mean() + 1e-3 * ent.mean()).backward()
    for p in cloud.flat_params():
        assert p.grad is not None and torch.isfinite(p.grad).all()
    assert float(cloud._xyz.grad.abs().sum()) > 0 and float(t["art_q"].grad.abs().sum()) > 0 and float(t["cam_t"].grad.abs().sum()) > 0
