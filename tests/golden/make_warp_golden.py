"""Generate tests/golden/warp_*.npz from the REFERENCE's own warp code, run on the CPU in this container.

Nothing is copied: at run time this script loads lab4d/utils/quat_transform.py from /root/reference by path.  Its compiled
`quaternion` CUDA helper (lab4d/third_party/quaternion) cannot be built here, and the file's own CPU branch only handles
4-vectors, while the warp multiplies 3-vectors (zero real part) with quaternions -- something only the CUDA helper does
(quaternion.cu:40-57: a 3-component operand gets w = 0).  quaternion_mul / quaternion_conjugate are therefore re-pointed
to a dispatcher over the file's OWN pure-torch kernels `_quaternion_mul`, `_quaternion_4D_mul_3D`, `_quaternion_3D_mul_4D`
and `_quaternion_conjugate` (same semantics as the CUDA helper).  The script then
exec()s the source text of get_bone_coords (lab4d/utils/transforms.py), dual_quaternion_skinning (lab4d/utils/geom_utils.py),
cross_entropy_skin_loss (lab4d/utils/loss_utils.py) and DeformableGaussian.apply_qt_to_gaussian
(lab4d/nnutils/deformable_gaussian.py) on seeded inputs, chaining them exactly as SkinningWarp.forward
(lab4d/nnutils/warping.py:378-444) and forward_warp (deformable_gaussian.py:1395-1434) do.  Outputs and autograd gradients are
stored as fixtures that pin oracle/warp_oracle.py.

    python tests/golden/make_warp_golden.py        # needs /root/reference
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def _load_reference():
    stub = types.ModuleType("quaternion")
    stub.quaternion_conjugate = lambda q: (_ for _ in ()).throw(RuntimeError("CUDA-only helper"))
    stub.quaternion_mul = lambda a, b: (_ for _ in ()).throw(RuntimeError("CUDA-only helper"))
    sys.modules["quaternion"] = stub
    spec = importlib.util.spec_from_file_location("ref_quat_transform", os.path.join(REF, "lab4d/utils/quat_transform.py"))
    qt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(qt)

    def qmul(a, b):                      # the CUDA helper's semantics (quaternion.cu:40-57) over the file's own torch kernels
        shape = torch.broadcast_shapes(a.shape[:-1], b.shape[:-1])
        a = a.expand(shape + a.shape[-1:]); b = b.expand(shape + b.shape[-1:])
        if a.shape[-1] == 3 and b.shape[-1] == 4:
            return qt._quaternion_3D_mul_4D(a, b)
        if a.shape[-1] == 4 and b.shape[-1] == 3:
            return qt._quaternion_4D_mul_3D(a, b)
        return qt._quaternion_mul(a, b)
    qt.quaternion_mul = qmul
    qt.quaternion_conjugate = qt._quaternion_conjugate
    ns = {k: getattr(qt, k) for k in dir(qt) if not k.startswith("__")}
    ns.update(torch=torch, F=F)

    def grab(path, name, in_class=None):
        src = open(os.path.join(REF, path)).read()
        tree = ast.parse(src)
        body = tree.body
        if in_class:
            body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == in_class).body
        node = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
        node.decorator_list = []
        import textwrap
        seg = textwrap.dedent(ast.get_source_segment(src, node))
        seg = seg.replace("@staticmethod\n", "")
        exec(seg, ns)
    grab("lab4d/utils/transforms.py", "get_bone_coords")
    grab("lab4d/utils/geom_utils.py", "dual_quaternion_skinning")
    grab("lab4d/utils/loss_utils.py", "cross_entropy_skin_loss")
    grab("lab4d/nnutils/deformable_gaussian.py", "apply_qt_to_gaussian", in_class="DeformableGaussian")
    return ns


def rand_unit_quat(rng, *shape):
    q = rng.normal(size=shape + (4,))
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def make_inputs(seed, P, B, M, with_delta):
    rng = np.random.default_rng(seed)
    xyz = rng.normal(size=(P, 3)) * 0.15
    rot = rand_unit_quat(rng, P)
    centres = rng.normal(size=(B, 3)) * 0.15
    rest_q = rand_unit_quat(rng, B)
    log_gauss = np.log(0.03) + 0.3 * rng.normal(size=(B, 3))          # init_gauss_scale = 0.03 (warping.py:349)
    # per frame: small articulation about the rest pose
    art_q = rand_unit_quat(rng, M, B) * 0.15 + rest_q[None]
    art_q /= np.linalg.norm(art_q, axis=-1, keepdims=True)
    art_t = centres[None] + 0.03 * rng.normal(size=(M, B, 3))
    cam_q = rand_unit_quat(rng, M) * 0.2 + np.array([1.0, 0, 0, 0])
    cam_q /= np.linalg.norm(cam_q, axis=-1, keepdims=True)
    cam_t = np.array([0.0, 0.0, 1.0]) + 0.05 * rng.normal(size=(M, 3))
    delta = np.maximum(rng.normal(size=(P, B)), 0.0) * 0.1 if with_delta else None
    f = lambda a: None if a is None else torch.tensor(a, dtype=torch.float32)  # noqa: E731
    return dict(xyz=f(xyz), rot=f(rot), rest_q=f(rest_q), rest_t=f(centres), log_gauss=f(log_gauss), art_q=f(art_q), art_t=f(art_t),
                cam_q=f(cam_q), cam_t=f(cam_t), delta=f(delta))


def run_reference(ns, inp):
    """SkinningWarp.forward (forward direction, return_qt) + forward_warp, on leaf tensors requiring grad."""
    t = {k: (v.clone().requires_grad_(True) if v is not None else None) for k, v in inp.items()}
    M, B = t["art_q"].shape[:2]
    P = t["xyz"].shape[0]
    q2dq = ns["quaternion_translation_to_dual_quaternion"]
    rest = q2dq(t["rest_q"], t["rest_t"])                                     # (B,4) x2
    art = q2dq(t["art_q"], t["art_t"])                                        # (M,B,4) x2
    rest_m = (rest[0][None].expand(M, -1, -1), rest[1][None].expand(M, -1, -1))
    se3 = ns["dual_quaternion_mul"](art, ns["dual_quaternion_inverse"](rest_m))   # warping.py:410-413
    xyz4 = t["xyz"][None, :, None, :].expand(M, -1, -1, -1)                   # (M,N,1,3) as forward_warp asserts
    articulation = (rest_m[0][:, None, None].expand(xyz4.shape[:3] + (-1, -1)),
                    rest_m[1][:, None, None].expand(xyz4.shape[:3] + (-1, -1)))
    xyz_bone = ns["get_bone_coords"](xyz4, articulation) / t["log_gauss"].exp().view(1, 1, 1, B, 3)   # skinning.py:126-142
    dist2 = xyz_bone.pow(2).sum(-1)
    skin = -(dist2 + t["delta"][None, :, None, :]) if t["delta"] is not None else -dist2
    q, tr = ns["dual_quaternion_skinning"](se3, xyz4, skin.softmax(-1), return_qt=True)
    entropy = ns["cross_entropy_skin_loss"](skin)                              # (M,N,1)
    rot4 = t["rot"][None, :, None, :].expand(M, -1, -1, -1)
    xyz_t, rot_t = ns["apply_qt_to_gaussian"](xyz4, rot4, q, tr, M)
    qc = t["cam_q"][:, None].repeat(1, xyz_t.shape[1], 1)
    tc = t["cam_t"][:, None].repeat(1, xyz_t.shape[1], 1)
    xyz_c, rot_c = ns["apply_qt_to_gaussian"](xyz_t, rot_t, qc, tc, M)
    return t, xyz_c.reshape(M, P, 3), rot_c.reshape(M, P, 4), entropy.reshape(M, P), se3


CASES = {"warp_b25_m3": dict(seed=1, P=400, B=25, M=3, with_delta=True),
         "warp_b7_m2_nodelta": dict(seed=2, P=300, B=7, M=2, with_delta=False)}

if __name__ == "__main__":
    ns = _load_reference()
    for name, kw in CASES.items():
        inp = make_inputs(**kw)
        t, xc, rc, ent, se3 = run_reference(ns, inp)
        rng = np.random.default_rng(kw["seed"] + 50)
        wx = torch.tensor(rng.normal(size=tuple(xc.shape)), dtype=torch.float32)
        wr = torch.tensor(rng.normal(size=tuple(rc.shape)), dtype=torch.float32)
        we = torch.tensor(rng.normal(size=tuple(ent.shape[1:])), dtype=torch.float32)
        ((xc * wx).sum() + (rc * wr).sum() + (ent[0] * we).sum()).backward()
        out = {"in_" + k: v.numpy() for k, v in inp.items() if v is not None}
        out.update(ref_xyz_cam=xc.detach().numpy(), ref_rot_cam=rc.detach().numpy(), ref_entropy=ent.detach().numpy(),
                   ref_se3_r=se3[0].detach().numpy(), ref_se3_d=se3[1].detach().numpy(), w_xyz=wx.numpy(), w_rot=wr.numpy(), w_ent=we.numpy())
        for k, v in t.items():
            if v is not None and v.grad is not None:
                out["ref_grad_" + k] = v.grad.numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("ref_")})
