import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.golden.make_golden import build_case, run_reference
from tests.test_gpu_parity import _run_ours, _np
from oracle import surfel_oracle as so
dev = torch.device("cuda:0")
inp = build_case(5000, 128, 128, 2)
ref = run_reference(inp, dev); r = _run_ours(inp, dev, with_grads=False)
st = so.forward(inp["means3D"], inp["opacities"], inp["scales"], inp["rotations"], shs=inp["shs"], sh_degree=3, W=128, H=128, tanfovx=0.5, tanfovy=0.5, bg=inp["bg"], viewmatrix=inp["viewmatrix"], projmatrix=inp["projmatrix"], campos=inp["campos"])
a = _np(r["n_contrib"]).astype(np.int64) & 0xffffffff; b = ref["img_n_contrib"].astype(np.int64) & 0xffffffff; c = st.n_contrib.astype(np.int64)
for pl in (0, 1):
    d = a[pl] != b[pl]
    print("plane", pl, "ours!=ref", d.sum(), "oracle!=ref", (c[pl] != b[pl]).sum(), "ours!=oracle", (a[pl] != c[pl]).sum())
    ys, xs = np.where(d)
    for y, x in list(zip(ys, xs))[:10]:
        t = (y // 16) * 8 + x // 16
        print("  px", y, x, "tile", t, "range", ref["img_ranges"][t], "ours", a[pl][y, x], "ref", b[pl][y, x], "oracle", c[pl][y, x], "last ours/ref", a[0][y, x], b[0][y, x], "T", float(r["final_T"][0, y, x]))
