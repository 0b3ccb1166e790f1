"""Debug helper: ours vs the live reference at the headline configuration -- where do the planes differ?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.test_gpu_headline import bench_inputs
from tests.test_gpu_parity import _run_ours, _np
from tests.golden.make_golden import run_reference
dev = torch.device("cuda:0")
for view in (None, 0):
    inp = bench_inputs(view)
    ref = run_reference(inp, dev)
    r = _run_ours(inp, dev)
    c, rc = _np(r["color"]), ref["color"]
    d = np.abs(c - rc)
    bad = np.argwhere(d.max(0) > 0)
    print("view", view, "R", r["num_rendered"], int(ref["num_rendered"][0]), "colour mismatching pixels", len(bad), "max", d.max())
    nc, rnc = _np(r["n_contrib"]).astype(np.int64) & 0xffffffff, ref["img_n_contrib"].astype(np.int64) & 0xffffffff
    print("  n_contrib[0] mismatches", int((nc[0] != rnc[0]).sum()), "keys equal", np.array_equal(_np(r["keys"]), ref["bin_keys"]),
          "point_list equal", np.array_equal(_np(r["point_list"]), ref["bin_point_list"]))
    am, rm = _np(r["allmap"]), ref["allmap"]
    for ch in range(8):
        print("  allmap", ch, "max abs diff", np.abs(am[ch] - rm[ch]).max(), "n diff", int((am[ch] != rm[ch]).sum()))
    for y, x in bad[:8]:
        print("   px", x, y, "ours", c[:, y, x], "ref", rc[:, y, x], "n_contrib", nc[0, y, x], rnc[0, y, x], "tile", (y // 16) * 32 + x // 16)
