import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.golden.make_golden import build_case
from tests.test_gpu_parity import _run_ours
dev = torch.device("cuda:0")
for (P, res) in ((300000, 512), (1000000, 1024)):
    inp = build_case(P, res, res, 41)
    r = _run_ours(inp, dev, with_grads=False)
    ctl = r["sort_ctl"].cpu().numpy()
    print(P, res, "R", r["num_rendered"], "sorted_sel", ctl[0], "npass", ctl[1], "skip", ctl[8:16], "src", ctl[16:24])
    rg = r["ranges"].cpu().numpy(); ln = rg[:,1]-rg[:,0]
    print("  tiles", len(ln), "nonempty", (ln>0).sum(), "max len", ln.max(), "p50/p90/p99 of nonempty", np.percentile(ln[ln>0],[50,90,99]))
