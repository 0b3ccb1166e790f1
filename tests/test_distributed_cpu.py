"""CPU, world_size 2, gloo: the N>1 path of the hot loop -- frames sharded across ranks, canonical-surfel gradients
summed by ONE all-reduce of a flat buffer -- gives the same gradient as one rank looping over all frames.
(The per-frame renderer here is the CPU oracle: this test is about the host-side sharding/collective logic.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from .conftest import ROOT

NF, P, W, H = 4, 300, 48, 32


def _frame_grads(frame):
    from oracle import surfel_oracle as so
    from vidu4d_b200.synthetic import object_scene, orbit_view, projection_matrix
    sc = object_scene(P, seed=5, center=(0.0, 0.0, 0.0))
    R, t = orbit_view(frame, 8)
    W2C = np.eye(4); W2C[:3, :3] = R; W2C[:3, 3] = t
    vm = W2C.T.astype(np.float32)
    st = so.forward(sc.means3D, sc.opacities, sc.scales, sc.rotations, shs=sc.shs, sh_degree=3, W=W, H=H, tanfovx=0.5,
                    tanfovy=0.5, bg=(0, 0, 0), viewmatrix=vm, projmatrix=projection_matrix(0.5, 0.5), campos=(-R.T @ t))
    rng = np.random.default_rng(frame)
    g = so.backward(st, rng.normal(size=(3, H, W)), rng.normal(size=(8, H, W)))
    return [g["dL_dmeans3D"], g["dL_dsh"], g["dL_dopacity"], g["dL_dscales"], g["dL_drotations"]]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from vidu4d_b200 import distributed as D
    r, w = D.init_distributed("gloo")
    assert (r, w) == (rank, world)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((P, 3), (P, 16, 3), (P, 1), (P, 2), (P, 4))]
    fg = D.FlatGrads(params)
    mine = D.shard_frames(NF, rank, world)
    for f in mine:
        for p, g in zip(params, _frame_grads(f)):
            p.grad.add_(torch.from_numpy(g.reshape(p.shape)))      # accumulates straight into the flat buffer
    flat = fg.allreduce_(average_over=NF)
    if rank == 0:
        np.save(out, flat.numpy())
    dist.destroy_process_group()


def test_shard_frames_partition():
    from vidu4d_b200.distributed import shard_frames
    for nf, w in ((64, 8), (7, 2), (3, 4)):
        parts = [shard_frames(nf, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(nf))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_flat_grads_views():
    from vidu4d_b200.distributed import FlatGrads
    ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))]
    fg = FlatGrads(ps)
    ps[1].grad += 2.0
    assert fg.flat[6:].eq(2).all() and fg.flat[:6].eq(0).all() and fg.nbytes == 44
    (ps[0].sum() * 3).backward()          # autograd accumulates in place into the flat buffer
    assert fg.flat[:6].eq(3).all()


@pytest.mark.timeout(300)
def test_two_rank_allreduce_equals_single_rank(built, tmp_path):
    out = str(tmp_path / "flat.npy")
    port = 29600 + (os.getpid() % 300)
    mp.start_processes(_worker, args=(2, port, out), nprocs=2, join=True, start_method="spawn")
    got = np.load(out)
    want = [np.zeros(s, np.float64) for s in ((P, 3), (P, 16, 3), (P, 1), (P, 2), (P, 4))]
    for f in range(NF):
        for a, g in zip(want, _frame_grads(f)):
            a += g.reshape(a.shape)
    want = np.concatenate([a.reshape(-1) for a in want]) / NF
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-5 * scale
