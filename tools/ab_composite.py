"""A/B the composite kernels: per-kernel ms/frame of the headline frame under SURFEL_GROUPS (and any other SURFEL_* switch).

Each configuration runs in a child process (the switches are read once per process)."""
import argparse, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, %r)
from vidu4d_b200 import rasterizer as R, _capi
from vidu4d_b200.synthetic import object_scene, orbit_view, projection_matrix
surfels, res, frames = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
sc = object_scene(surfels, seed=0, opacity="trained", center=(0, 0, 0)); t = sc.to_torch(dev)
e = torch.empty((0,), device=dev); bg = torch.zeros(3, device=dev)
g = torch.Generator(device=dev).manual_seed(1234)
dLc = torch.randn((3, res, res), device=dev, generator=g); dLo = torch.randn((8, res, res), device=dev, generator=g) * 0.1
P = projection_matrix(0.5, 0.5).astype(np.float64)
def frame(f):
    Rm, tt = orbit_view(f, 64); W2C = np.eye(4); W2C[:3, :3] = Rm; W2C[:3, 3] = tt
    vm = torch.from_numpy(W2C.T.astype(np.float32)).to(dev); pm = torch.from_numpy((W2C.T @ P).astype(np.float32)).to(dev)
    cp = torch.from_numpy((-Rm.T @ tt).astype(np.float32)).to(dev)
    o = R._C.rasterize_gaussians(bg, t["means3D"], e, t["opacities"], t["scales"], t["rotations"], 1.0, e, vm, pm, 0.5, 0.5, res, res, t["shs"], 3, cp, False, False)
    gr = R._C.rasterize_gaussians_backward(bg, t["means3D"], o[3], e, t["scales"], t["rotations"], 1.0, e, vm, pm, 0.5, 0.5, dLc, dLo, t["shs"], 3, cp, o[4], o[0], o[5], o[6], False)
    return o, gr
for f in range(3): frame(f)
torch.cuda.synchronize()
_capi.set_profiling(True); _capi.get_profile()
chk = 0.0
for f in range(frames):
    o, gr = frame(f)
torch.cuda.synchronize()
prof = _capi.get_profile()
out = {k: v["ms"] / frames for k, v in prof.items()}
out["_checksum"] = [float(o[1].double().sum()), float(o[2].double().sum()), float(gr[2].double().abs().sum()), float(gr[0].double().abs().sum())]
print("RESULT " + json.dumps(out))
''' % ROOT

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--surfels", type=int, default=300000); ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--configs", default="GROUPS=8;GROUPS=4")
    a = ap.parse_args()
    rows = {}
    for cfg in a.configs.split(";"):
        env = dict(os.environ)
        for kv in cfg.split(","):
            k, v = kv.split("="); env["SURFEL_" + k] = v
        r = subprocess.run([sys.executable, "-c", CHILD, str(a.surfels), str(a.res), str(a.frames)], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(cfg, "FAILED", r.stderr[-2000:]); continue
        d = json.loads(line[0][7:]); rows[cfg] = d
        print(cfg, "fwd %.4f bwd %.4f total %.4f" % (d.get("composite_fwd", 0), d.get("composite_bwd", 0), sum(v for k, v in d.items() if not k.startswith("_"))), d["_checksum"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "ab_composite.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
