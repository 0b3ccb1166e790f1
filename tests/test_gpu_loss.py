"""GPU parity of the fused Stage-3 inner loop (render_loss_batch: batched rasterizer + csrc/loss.cu) against the float64
oracle of the reference's loss code (oracle/post_oracle.py: lab4d/engine/model.py:649-653,674-692,817-842 on top of
gs/gaussian_renderer/__init__.py:121-145): the four loss terms per frame and every parameter gradient, 1e-4."""
import numpy as np
import pytest
import torch

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


@pytest.mark.parametrize("with_bkgd,depth_ratio", [(True, 0.0), (False, 0.3)])
def test_render_loss_batch_matches_oracle(with_bkgd, depth_ratio, dev):
    from oracle import post_oracle as po
    from vidu4d_b200 import rasterizer as RZ
    from vidu4d_b200.renderer import PipelineParams, make_camera, render_loss_batch, stack_cameras
    from vidu4d_b200.synthetic import SurfelCloud, object_scene, orbit_view
    M, P, W, H = 3, 6000, 160, 112
    tanx, tany = 0.5, 0.35
    cams = []
    for f in range(M):
        R, t = orbit_view(5 * f + 2, 64)
        cams.append(make_camera(W, H, 2 * np.arctan(tanx), 2 * np.arctan(tany), R=R.T, T=t, device=dev))
    bc = stack_cameras(cams)
    # hand the matrices in as a STRIDED view (as bench.py does: one (M,2,4,4) upload holding view and projection matrices)
    both = torch.stack((bc.world_view_transform, bc.full_proj_transform), 1)
    bc.world_view_transform, bc.full_proj_transform = both[:, 0], both[:, 1]
    scene = object_scene(P, seed=21, center=(0.0, 0.0, 0.0))
    gen = torch.Generator(device=dev).manual_seed(7)
    target = torch.rand((M, 3, H, W), device=dev, generator=gen)
    vis = (torch.rand((M, H, W), device=dev, generator=gen) > 0.15).float()
    shift = 0.01 * torch.randn((M, P, 3), device=dev, generator=gen)
    bg = torch.zeros(3, device=dev)
    kw = dict(w_rgb=0.8, w_mask=0.1, lambda_normal=0.05, lambda_dist=100.0)
    pipe = PipelineParams(depth_ratio=depth_ratio)

    cloud = SurfelCloud(scene, dev)
    bkgd = torch.tensor([0.3, 0.5, 0.2], device=dev, requires_grad=True) if with_bkgd else None
    # ground-truth masks from a first, gradient-free render
    with torch.no_grad():
        pre = render_loss_batch(bc, cloud, pipe, bg, target, means3D=cloud.get_xyz[None] + shift, **dict(kw, w_mask=0.0))
    mask = (pre["allmap"][:, 1] > 0.5).float()
    wt = torch.stack([torch.from_numpy(po.mask_balance_wt(mask[f].cpu().numpy(), vis[f].cpu().numpy())).float() for f in range(M)]).to(dev)
    out = render_loss_batch(bc, cloud, pipe, bg, target, vis2d=vis, mask_gt=mask, mask_wt=wt, learnable_bkgd=bkgd,
                            means3D=cloud.get_xyz[None] + shift, **kw)
    (out["loss"] * 1.7).backward()                    # a non-unit upstream scalar exercises grad_scale
    got = [p.grad.detach().cpu().numpy() for p in cloud.flat_params()]
    assert out["terms"].shape == (M, 4) and abs(float(out["terms"].sum()) - float(out["loss"].detach())) <= 1e-5 * abs(float(out["loss"].detach()))

    # ---- oracle: per frame, losses on the rasterizer's planes; VJP pushed through the single-frame rasterizer backward
    cloud2 = SurfelCloud(scene, dev)
    g_bk = np.zeros(3)
    for f in range(M):
        rs = RZ.GaussianRasterizationSettings(H, W, tanx, tany, bg, 1.0, bc.world_view_transform[f], bc.full_proj_transform[f],
                                              cloud2.active_sh_degree, bc.camera_center[f], False, False)
        m2d = torch.zeros((P, 3), device=dev, requires_grad=True)
        color, radii, allmap = RZ.GaussianRasterizer(rs)(means3D=cloud2.get_xyz + shift[f], means2D=m2d, shs=cloud2.get_features,
                                                         opacities=cloud2.get_opacity, scales=cloud2.get_scaling,
                                                         rotations=cloud2.get_rotation)
        assert torch.equal(color.detach(), out["render"][f]) and torch.equal(allmap.detach(), out["allmap"][f])
        args = (color.detach().cpu().numpy(), allmap.detach().cpu().numpy(), bc.world_view_transform[f].cpu().numpy(), tanx, tany,
                depth_ratio, target[f].cpu().numpy(), vis[f].cpu().numpy(), mask[f].cpu().numpy(), wt[f].cpu().numpy())
        okw = dict(kw, bkgd=bkgd.detach().cpu().numpy() if with_bkgd else None)
        total, terms = po.stage3_losses(*args, **okw)
        mine = out["terms"][f].cpu().numpy()
        for i, k in enumerate(("rgb", "mask", "normal", "dist")):
            assert abs(mine[i] - terms[k]) <= TOL * max(abs(terms[k]), 1e-6), (f, k, mine[i], terms[k])
        gr = po.stage3_losses_backward(*args, **okw)
        if with_bkgd:
            g_bk += 1.7 * gr[2]
        torch.autograd.backward([color, allmap], [torch.from_numpy((1.7 * gr[0]).astype(np.float32)).to(dev),
                                                  torch.from_numpy((1.7 * gr[1]).astype(np.float32)).to(dev)])
        assert _rel(out["viewspace_points"].grad[f].cpu().numpy(), m2d.grad.cpu().numpy()) <= TOL
    for a, p in zip(got, cloud2.flat_params()):
        assert _rel(a, p.grad.cpu().numpy()) <= TOL, tuple(p.shape)
    if with_bkgd:
        assert _rel(bkgd.grad.cpu().numpy(), g_bk) <= TOL
