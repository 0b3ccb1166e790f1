"""GPU parity of the BENCHMARKED kernels at the BENCHMARKED configuration (300 K surfels, 512x512, bench.py's scene
and cameras) against the live reference extension (oracle/_ref/_C.so, the unmodified reference compiled for sm_100a):
tile assignment / sort / ranges / contributor counts bit-exact, rendered planes and all eight gradient tensors within
1e-4 (north_star).  Skipped only when the reference .so did not travel with the repo.
"""
import numpy as np
import pytest
import torch

from .test_gpu_parity import GRADS, TOL, _assert_n_contrib, _np, _run_ours

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(built):
    from vidu4d_b200 import _capi
    _capi.load()
    return torch.device("cuda:0")


def bench_inputs(view, P=300_000, res=512):
    """The exact inputs bench.py renders: object_scene(seed 0, trained opacities) at the world origin, orbit camera
    `view` of 64 (view=None: the Stage-3 identity camera with the object at z = 1)."""
    from vidu4d_b200.synthetic import object_scene, orbit_view, projection_matrix
    tan = 0.5
    Pm = projection_matrix(tan, tan).astype(np.float64)
    if view is None:
        sc = object_scene(P, seed=0, opacity="trained", center=(0.0, 0.0, 1.0))
        vm = np.eye(4, dtype=np.float32); cp = np.zeros(3, np.float32)
    else:
        sc = object_scene(P, seed=0, opacity="trained", center=(0.0, 0.0, 0.0))
        R, t = orbit_view(view, 64)
        W2C = np.eye(4); W2C[:3, :3] = R; W2C[:3, 3] = t
        vm = W2C.T.astype(np.float32); cp = (-R.T @ t).astype(np.float32)
    pm = (vm.astype(np.float64) @ Pm).astype(np.float32)
    rng = np.random.default_rng(1234)
    return dict(means3D=sc.means3D, scales=sc.scales, rotations=sc.rotations, opacities=sc.opacities, shs=sc.shs,
                colors_precomp=np.zeros((0,), np.float32), viewmatrix=vm, projmatrix=pm, campos=cp,
                bg=np.zeros(3, np.float32), dL_dcolor=rng.normal(size=(3, res, res)).astype(np.float32),
                dL_dallmap=(0.1 * rng.normal(size=(8, res, res))).astype(np.float32),
                meta=np.array([P, res, res, 3, 0], np.int64), tanfov=np.array([tan, tan], np.float32))


@pytest.mark.parametrize("view", [None, 0, 17], ids=["identity", "orbit0", "orbit17"])
def test_headline_against_live_reference(view, dev):
    from oracle import ref_ext
    if not ref_ext.available():
        pytest.skip("oracle/_ref/_C.so not present on this box")
    from tests.golden.make_golden import run_reference
    inp = bench_inputs(view)
    ref = run_reference(inp, dev)
    r = _run_ours(inp, dev)
    R = int(ref["num_rendered"][0])
    assert R > 300_000 and r["num_rendered"] == R
    # ---- integer / index work: bit-exact
    for a, b, what in ((r["radii"], ref["radii"], "radii"), (r["keys"], ref["bin_keys"], "sorted keys"),
                       (r["point_list"], ref["bin_point_list"], "sorted surfel ids"), (r["ranges"], ref["img_ranges"], "tile ranges")):
        np.testing.assert_array_equal(_np(a), b, err_msg=what)
    _assert_n_contrib(_np(r["n_contrib"]), ref["img_n_contrib"], ref["img_ranges"], 512, 512)
    # ---- rendered planes: 1e-4 relative (north_star).  Observed: depth/alpha/normal/median planes bit-identical; colour
    # within 1 ulp on ~0.1 % of the pixels (the SH -> RGB evaluation of a surfel is value-level, not FMA-mapped, since
    # no binning decision depends on it), distortion within 1e-7 (fp32 depth mapping).  Assert a
    # bound 100x tighter than the contract so that a real regression cannot hide.
    assert np.abs(_np(r["color"]) - ref["color"]).max() <= 1e-6 * max(1.0, np.abs(ref["color"]).max())
    am, rm = _np(r["allmap"]), ref["allmap"]
    for ch in range(8):
        assert np.abs(am[ch] - rm[ch]).max() <= TOL * max(np.abs(rm[ch]).max(), 1e-30), f"allmap[{ch}]"
    for ch in (0, 1, 2, 3, 4, 5, 7):
        assert np.array_equal(am[ch], rm[ch]), f"allmap[{ch}] is expected bit-identical to the reference build"
    # ---- all eight gradient tensors: 1e-4 of the tensor's max magnitude
    for k in GRADS:
        b = ref["grad_" + k]
        if b.size == 0:
            continue
        a = _np(r["grads"][k]).reshape(b.shape)
        assert np.abs(a - b).max() <= TOL * (np.abs(b).max() + 1e-30), k
