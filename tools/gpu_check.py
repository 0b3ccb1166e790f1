"""Developer script (GPU box): stage-by-stage parity of the CUDA library against the unmodified reference
extension (oracle/_ref/_C.so) and the CPU oracle, plus quick timings.  Writes gpurun_out/gpu_check.json."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_ext, surfel_oracle as so
from vidu4d_b200 import rasterizer as R, debug
from vidu4d_b200.synthetic import object_scene, rigid_view, orbit_view, projection_matrix, random_rotation

dev = torch.device("cuda:0")
report = {}

def run_case(name, P, W, H, seed, rigid=False, opacity="trained", colors_precomp=False, sh_degree=3, bgv=(0.0,0.0,0.0), timing=False):
    sc = object_scene(P, seed=seed, opacity=opacity, sh_degree=sh_degree)
    vm = np.eye(4, dtype=np.float32); campos = np.zeros(3, np.float32)
    if rigid:
        rng = np.random.default_rng(seed + 100)
        # x_c = R_wc x_w + t_wc : express the camera-space scene in a rotated/translated world frame
        sc, vm, campos = rigid_view(sc, random_rotation(rng), np.array([0.3, -0.2, 0.5]))
    t = sc.to_torch(dev)
    tan = 0.5
    bg = torch.tensor(bgv, dtype=torch.float32, device=dev)
    vmt = torch.from_numpy(vm).to(dev); pm = torch.from_numpy(projection_matrix(tan, tan)).to(dev) ; pm = (vmt @ pm).contiguous()
    cp = torch.from_numpy(campos).to(dev)
    col = None; shs = t["shs"]
    if colors_precomp:
        col = torch.rand((P, 3), device=dev, generator=torch.Generator(device=dev).manual_seed(seed)); shs = None
    kw = dict(sh_degree=sh_degree, W=W, H=H, tanfovx=tan, tanfovy=tan, bg=bg, viewmatrix=vmt, projmatrix=pm, campos=cp)
    g = torch.Generator(device=dev).manual_seed(seed + 7)
    dLc = torch.randn((3, H, W), device=dev, generator=g); dLo = torch.randn((8, H, W), device=dev, generator=g)
    res = {"P": P, "W": W, "H": H}
    # ---- reference ext
    fw = ref_ext.forward(t["means3D"], t["opacities"], t["scales"], t["rotations"], shs=shs, colors_precomp=col, **kw)
    bwkw = {k: v for k, v in kw.items() if k not in ("W", "H")}
    gb = ref_ext.backward(fw, t["means3D"], t["scales"], t["rotations"], shs=shs, colors_precomp=col, dL_dcolor=dLc, dL_dallmap=dLo, **bwkw)
    Rn = int(fw["num_rendered"]); res["R_ref"] = Rn
    rb = ref_ext.decode_binning(fw["binningBuffer"], Rn); rg = ref_ext.decode_geom(fw["geomBuffer"], P); ri = ref_ext.decode_image(fw["imgBuffer"], W, H)
    # ---- ours
    e = torch.empty((0,), device=dev)
    out = R._C.rasterize_gaussians(bg, t["means3D"], col if col is not None else e, t["opacities"], t["scales"], t["rotations"], 1.0, e, vmt, pm, tan, tan, H, W, shs if shs is not None else e, sh_degree, cp, False, True)
    nr, color, allmap, radii, gbuf, bbuf, ibuf = out
    res["R_ours"] = nr
    d = debug.decode(gbuf, bbuf, ibuf, P, W, H, nr)
    res["radii_mismatch"] = int((radii != fw["radii"]).sum().item())
    vis = fw["radii"] > 0
    res["visible"] = int(vis.sum().item())
    res["depth_mismatch"] = int((d["depths"][vis].view(torch.int32) != rg["depths"][vis].view(torch.int32)).sum().item())
    res["tiles_touched_mismatch"] = int((d["tiles_touched"] != rg["tiles_touched"]).sum().item())
    rec = d["surfel_rec"]
    res["transMat_bits_mismatch"] = int((rec[vis][:, 0:9].contiguous().view(torch.int32) != rg["transMat"][vis].contiguous().view(torch.int32)).sum().item())
    res["means2D_bits_mismatch"] = int((rec[vis][:, 9:11].contiguous().view(torch.int32) != rg["means2D"][vis].contiguous().view(torch.int32)).sum().item())
    res["normal_bits_mismatch"] = int((rec[vis][:, 12:15].contiguous().view(torch.int32) != rg["normal_opacity"][vis][:, :3].contiguous().view(torch.int32)).sum().item())
    res["rgb_maxabs"] = float((rec[vis][:, 15:18] - rg["rgb"][vis]).abs().max().item()) if col is None else 0.0
    if nr == Rn:
        res["keys_mismatch"] = int((d["keys"] != rb["keys"]).sum().item())
        res["point_list_mismatch"] = int((d["point_list"] != rb["point_list"]).sum().item())
    res["ranges_mismatch"] = int((d["ranges"] != ri["ranges"]).sum().item())
    res["n_contrib_last_mismatch"] = int((d["n_contrib"][0] != ri["n_contrib"][0]).sum().item())
    res["final_T_bits_mismatch"] = int((d["final_T"].view(torch.int32) != ri["final_T"].view(torch.int32)).sum().item())
    res["color_bits_mismatch"] = int((color.view(torch.int32) != fw["color"].view(torch.int32)).sum().item())
    res["color_maxabs"] = float((color - fw["color"]).abs().max().item())
    names = ["depth", "alpha", "nx", "ny", "nz", "median_depth", "distortion", "median_weight"]
    for i, nme in enumerate(names):
        res[f"allmap_{nme}_bits_mismatch"] = int((allmap[i].view(torch.int32) != fw["allmap"][i].view(torch.int32)).sum().item())
        res[f"allmap_{nme}_maxabs"] = float((allmap[i] - fw["allmap"][i]).abs().max().item())
    # ---- backward
    go = R._C.rasterize_gaussians_backward(bg, t["means3D"], radii, col if col is not None else e, t["scales"], t["rotations"], 1.0, e, vmt, pm, tan, tan, dLc, dLo, shs if shs is not None else e, sh_degree, cp, gbuf, nr, bbuf, ibuf, True)
    gnames = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dtransMat", "dL_dsh", "dL_dscales", "dL_drotations"]
    # second reference run -> the reference's own run-to-run (atomic order) noise floor
    gb2 = ref_ext.backward(fw, t["means3D"], t["scales"], t["rotations"], shs=shs, colors_precomp=col, dL_dcolor=dLc, dL_dallmap=dLo, **bwkw)
    for nme, ours in zip(gnames, go):
        ref = gb[nme]
        if ref.numel() == 0: continue
        scale = ref.abs().max().item() + 1e-30
        res[f"{nme}_relmax"] = float((ours - ref).abs().max().item() / scale)
        res[f"{nme}_ref_selfnoise"] = float((gb2[nme] - ref).abs().max().item() / scale)
        denom = ref.abs().clamp_min(1e-3 * scale)
        res[f"{nme}_p999_elemrel"] = float(torch.quantile(((ours - ref).abs() / denom).flatten().float()[:4_000_000], 0.999).item())
    # ---- CPU oracle vs reference (small cases only)
    if P <= 20000:
        st = so.forward(sc.means3D, sc.opacities, sc.scales, sc.rotations, shs=None if colors_precomp else sc.shs, colors_precomp=None if col is None else col.cpu().numpy(), sh_degree=sh_degree, W=W, H=H, tanfovx=tan, tanfovy=tan, bg=bgv, viewmatrix=vm, projmatrix=pm.cpu().numpy(), campos=campos)
        res["oracle_R"] = st.num_rendered
        res["oracle_radii_mismatch"] = int((st.radii != fw["radii"].cpu().numpy()).sum())
        if st.num_rendered == Rn:
            res["oracle_point_list_mismatch"] = int((st.point_list.astype(np.int64) != rb["point_list"].cpu().numpy().astype(np.int64)).sum())
            res["oracle_keys_mismatch"] = int((st.keys.astype(np.int64) != rb["keys"].cpu().numpy()).sum())
        res["oracle_color_maxabs"] = float(np.abs(st.color - fw["color"].cpu().numpy()).max())
        res["oracle_allmap_maxabs"] = [float(np.abs(st.allmap[i] - fw["allmap"][i].cpu().numpy()).max()) for i in range(8)]
        og = so.backward(st, dLc.cpu().numpy(), dLo.cpu().numpy())
        for nme in gnames:
            ref = gb[nme].cpu().numpy()
            if ref.size == 0: continue
            scale = np.abs(ref).max() + 1e-30
            res[f"oracle_{nme}_relmax"] = float(np.abs(og[nme].reshape(ref.shape) - ref).max() / scale)
    if timing:
        def tm(fn, n=20):
            for _ in range(3): fn()
            torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(n): fn()
            e2.record(); torch.cuda.synchronize(); return s.elapsed_time(e2) / n
        def ours_f():
            return R._C.rasterize_gaussians(bg, t["means3D"], col if col is not None else e, t["opacities"], t["scales"], t["rotations"], 1.0, e, vmt, pm, tan, tan, H, W, shs if shs is not None else e, sh_degree, cp, False, False)
        def ours_fb():
            o = ours_f()
            R._C.rasterize_gaussians_backward(bg, t["means3D"], o[3], col if col is not None else e, t["scales"], t["rotations"], 1.0, e, vmt, pm, tan, tan, dLc, dLo, shs if shs is not None else e, sh_degree, cp, o[4], o[0], o[5], o[6], False)
        def ref_f():
            return ref_ext.forward(t["means3D"], t["opacities"], t["scales"], t["rotations"], shs=shs, colors_precomp=col, **kw)
        def ref_fb():
            f = ref_f()
            ref_ext.backward(f, t["means3D"], t["scales"], t["rotations"], shs=shs, colors_precomp=col, dL_dcolor=dLc, dL_dallmap=dLo, **bwkw)
        res["ms_ours_fwd"] = tm(ours_f); res["ms_ours_fwdbwd"] = tm(ours_fb)
        res["ms_ref_fwd"] = tm(ref_f); res["ms_ref_fwdbwd"] = tm(ref_fb)
        R.set_sync_mode(False)
        res["ms_ours_fwd_nosync"] = tm(ours_f); res["ms_ours_fwdbwd_nosync"] = tm(ours_fb)
        R.check_overflow(); R.set_sync_mode(True)
    report[name] = res
    print(name, json.dumps(res), flush=True)

cases = [
    ("tiny_id", dict(P=1000, W=64, H=64, seed=1)),
    ("small_id", dict(P=5000, W=128, H=128, seed=2)),
    ("small_rigid", dict(P=5000, W=128, H=96, seed=3, rigid=True, bgv=(0.2, 0.5, 0.7))),
    ("small_precomp", dict(P=3000, W=100, H=70, seed=4, colors_precomp=True, opacity="init")),
    ("deg1", dict(P=3000, W=128, H=128, seed=5, sh_degree=1)),
    ("c2_100k", dict(P=100000, W=512, H=512, seed=6, timing=True)),
    ("c2_rigid", dict(P=100000, W=512, H=512, seed=7, rigid=True)),
    ("hl_300k", dict(P=300000, W=512, H=512, seed=8, timing=True)),
    ("hl_300k_init", dict(P=300000, W=512, H=512, seed=9, opacity="init", timing=True)),
]
only = sys.argv[1:] 
for name, kwargs in cases:
    if only and name not in only: continue
    try:
        run_case(name, **kwargs)
    except Exception as ex:
        import traceback; traceback.print_exc()
        report[name] = {"error": repr(ex)}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(report, open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w"), indent=1)
