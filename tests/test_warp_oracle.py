"""CPU: oracle/warp_oracle.py (float64 torch restatement of the Stage-3 bob-skinning warp) against fixtures produced by the
REFERENCE's own source text (tests/golden/make_warp_golden.py): outputs and, through autograd, every gradient."""
import os

import numpy as np
import pytest
import torch

from .conftest import GOLDEN_DIR
from oracle import warp_oracle as wo

CASES = ["warp_b25_m3", "warp_b7_m2_nodelta"]


def _inputs(g, dtype=torch.float64, requires_grad=False):
    t = {}
    for k in ("xyz", "rot", "rest_q", "rest_t", "log_gauss", "art_q", "art_t", "cam_q", "cam_t", "delta"):
        if "in_" + k in g:
            t[k] = torch.tensor(g["in_" + k], dtype=dtype, requires_grad=requires_grad)
    return t


def oracle_from_leaves(t):
    """The same leaves as the reference chain (quaternion + translation per bone / frame) -> warp_oracle.bob_warp inputs."""
    def q2dq(q, tr):   # quaternion_translation_to_dual_quaternion (quat_transform.py:294-301): q_d = 0.5 * t (x) q
        t4 = torch.cat((torch.zeros_like(tr[..., :1]), tr), -1)
        return q, 0.5 * wo.qmul(t4, q)
    rest = q2dq(t["rest_q"], t["rest_t"])
    art = q2dq(t["art_q"], t["art_t"])
    M = t["art_q"].shape[0]
    rest_m = (rest[0][None].expand(M, -1, -1), rest[1][None].expand(M, -1, -1))
    se3 = wo.dq_mul(art, wo.dq_inverse(rest_m))
    return wo.bob_warp(t["xyz"], t["rot"], rest, torch.exp(-t["log_gauss"]), se3, (t["cam_q"], t["cam_t"]), t.get("delta")), se3


@pytest.mark.parametrize("name", CASES)
def test_warp_oracle_matches_reference_outputs_and_gradients(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    t = _inputs(g, requires_grad=True)
    (xc, rc, ent), se3 = oracle_from_leaves(t)
    for mine, ref, tol in ((se3[0], g["ref_se3_r"], 1e-6), (se3[1], g["ref_se3_d"], 1e-6), (xc, g["ref_xyz_cam"], 2e-6),
                           (rc, g["ref_rot_cam"], 2e-6), (ent, g["ref_entropy"][0], 2e-5)):
        assert np.abs(mine.detach().numpy() - ref).max() <= tol * max(1.0, np.abs(ref).max())
    loss = (xc * torch.tensor(g["w_xyz"], dtype=torch.float64)).sum() + (rc * torch.tensor(g["w_rot"], dtype=torch.float64)).sum() \
        + (ent * torch.tensor(g["w_ent"], dtype=torch.float64)).sum()
    loss.backward()
    for k, v in t.items():
        ref = g.get("ref_grad_" + k)
        if ref is None:
            continue
        err = np.abs(v.grad.numpy() - ref).max() / (np.abs(ref).max() + 1e-30)
        assert err <= 2e-4, (k, err)            # the reference ran in float32


def test_warp_oracle_basic_properties():
    g = dict(np.load(os.path.join(GOLDEN_DIR, "warp_b25_m3.npz")))
    t = _inputs(g)
    (xc, rc, ent), _ = oracle_from_leaves(t)
    # orientations stay unit quaternions; identity articulation + identity camera is the identity map
    assert np.abs(rc.norm(dim=-1).numpy() - 1.0).max() < 1e-6     # inputs are float32-normalised
    t2 = dict(t)
    t2["art_q"], t2["art_t"] = t["rest_q"][None].expand(3, -1, -1), t["rest_t"][None].expand(3, -1, -1)
    t2["cam_q"] = torch.tensor([[1.0, 0, 0, 0]] * 3, dtype=torch.float64); t2["cam_t"] = torch.zeros((3, 3), dtype=torch.float64)
    (x0, r0, _), _ = oracle_from_leaves(t2)
    assert np.abs(x0.numpy() - t["xyz"].numpy()[None]).max() < 1e-6 and np.abs(r0.numpy() - t["rot"].numpy()[None]).max() < 1e-6
    assert (ent.numpy() >= 0).all()
