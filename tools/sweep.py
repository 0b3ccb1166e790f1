"""BASELINE.json configs[4]: tile-sort + composite sweep over surfel count x resolution (1 GPU).
For each (P, res): device-resident fwd+bwd ms per frame for ours (ONE batched call over the 4 views; also single-frame calls) and
the reference extension (if oracle/_ref/_C.so is present),
R, and the achieved algorithmic bandwidth bytes_alg / t  (bytes_alg = 1002 P + 324 R + 128 N)
against the measured HBM peak.  Writes gpurun_out/sweep.json + a markdown table."""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from vidu4d_b200 import rasterizer as RZ
from vidu4d_b200.synthetic import object_scene, orbit_view, projection_matrix

ap = argparse.ArgumentParser()
ap.add_argument("--surfels", default="10000,30000,100000,300000,1000000,3000000")
ap.add_argument("--res", default="256,512,1024,2048")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
peaks = {}
try: peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception: pass
peak = float(peaks.get("hbm_gbs", 6650.0))
ref = None
try:
    from oracle import ref_ext
    if ref_ext.available(): ref = ref_ext.load()
except Exception: pass
P64 = projection_matrix(0.5, 0.5).astype(np.float64)
rows = []
for P in [int(x) for x in a.surfels.split(",")]:
    sc = object_scene(P, seed=0, center=(0, 0, 0)); t = sc.to_torch(dev)
    for res in [int(x) for x in a.res.split(",")]:
        e = torch.empty((0,), device=dev); bg = torch.zeros(3, device=dev)
        g = torch.Generator(device=dev).manual_seed(1)
        dLc = torch.randn((3, res, res), device=dev, generator=g); dLo = torch.randn((8, res, res), device=dev, generator=g) * 0.1
        views = []
        for f in range(4):
            R, tt = orbit_view(f, 16); W2C = np.eye(4); W2C[:3, :3] = R; W2C[:3, 3] = tt
            views.append((torch.from_numpy(W2C.T.astype(np.float32)).to(dev), torch.from_numpy((W2C.T @ P64).astype(np.float32)).to(dev),
                          torch.from_numpy((-R.T @ tt).astype(np.float32)).to(dev)))
        def run(C, v):
            vm, pm, cp = views[v % 4]
            o = C.rasterize_gaussians(bg, t["means3D"], e, t["opacities"], t["scales"], t["rotations"], 1.0, e, vm, pm, 0.5, 0.5, res, res, t["shs"], 3, cp, False, False)
            C.rasterize_gaussians_backward(bg, t["means3D"], o[3], e, t["scales"], t["rotations"], 1.0, e, vm, pm, 0.5, 0.5, dLc, dLo, t["shs"], 3, cp, o[4], o[0], o[5], o[6], False)
            return o[0]
        def tm(C, sync_free):
            for i in range(3): Rn = run(C, i)
            if sync_free: RZ.check_overflow()
            torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); f_ = torch.cuda.Event(enable_timing=True)
            s.record()
            for i in range(a.iters): run(C, i)
            f_.record(); torch.cuda.synchronize()
            if sync_free: RZ.check_overflow()
            return s.elapsed_time(f_) / a.iters, Rn
        RZ.set_sync_mode(True); Rn = run(RZ._C, 0)          # establishes the capacity hint
        RZ.set_sync_mode(False)
        ms1, _ = tm(RZ._C, True)
        # the same 4 views as ONE batched launch set
        vm4 = torch.stack([v[0] for v in views]); pm4 = torch.stack([v[1] for v in views]); cp4 = torch.stack([v[2] for v in views])
        dLc4 = dLc.expand(4, -1, -1, -1).contiguous(); dLo4 = dLo.expand(4, -1, -1, -1).contiguous()
        def runb():
            o = RZ._C.rasterize_gaussians_batch(bg, t["means3D"], e, t["opacities"], t["scales"], t["rotations"], 1.0, vm4, pm4, 0.5, 0.5, res, res, t["shs"], 3, cp4)
            RZ._C.rasterize_gaussians_backward_batch(bg, t["means3D"], o[3], e, t["scales"], t["rotations"], 1.0, vm4, pm4, 0.5, 0.5, dLc4, dLo4, t["shs"], 3, cp4, o[4], o[5], o[6], sum_shared=True, want_transmat=False)
        RZ.set_sync_mode(True); runb(); RZ.set_sync_mode(False)
        for i in range(2): runb()
        RZ.check_overflow(); torch.cuda.synchronize()
        s_ = torch.cuda.Event(enable_timing=True); f_ = torch.cuda.Event(enable_timing=True)
        nb = max(2, a.iters // 2)
        s_.record()
        for i in range(nb): runb()
        f_.record(); torch.cuda.synchronize(); RZ.check_overflow()
        ms = s_.elapsed_time(f_) / (4 * nb)
        del dLc4, dLo4
        RZ.set_sync_mode(True)
        ms_ref = None
        if ref is not None:
            try: ms_ref, _ = tm(ref, False)
            except Exception as ex: ms_ref = None
        alg = 1002 * P + 324 * Rn + 128 * res * res
        row = dict(P=P, res=res, R=int(Rn), ms=round(ms, 4), ms_single_call=round(ms1, 4), fps=round(1e3 / ms, 1), ms_ref=None if ms_ref is None else round(ms_ref, 4),
                   speedup=None if ms_ref is None else round(ms_ref / ms, 2), alg_MB=round(alg / 1e6, 1),
                   GBps=round(alg / 1e9 / (ms * 1e-3), 1), frac_of_hbm_peak=round(alg / 1e9 / (ms * 1e-3) / peak, 4))
        rows.append(row); print(json.dumps(row), flush=True)
        torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(hbm_peak_gbs=peak, rows=rows), open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
with open(os.path.join(ROOT, "gpurun_out", "sweep.md"), "w") as f:
    f.write("| surfels | res | R | ours ms/frame (batch of 4) | ours ms (single-frame calls) | fps | reference ms | speed-up | alg MB | GB/s | of measured HBM peak |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write(f"| {r['P']} | {r['res']}^2 | {r['R']} | {r['ms']} | {r['ms_single_call']} | {r['fps']} | {r['ms_ref']} | {r['speedup']} | {r['alg_MB']} | {r['GBps']} | {r['frac_of_hbm_peak']} |\n")
