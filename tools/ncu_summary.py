"""Summarise an `ncu --set full` report (.ncu-rep) into one text block per kernel: duration, DRAM bytes and %, L2 %, occupancy,
issue utilisation, instruction counts, L2 reduction sectors, bank conflicts and the top stall reasons.
usage: python tools/ncu_summary.py <report.ncu-rep> [outdir]   (runs `ncu -i ... --page raw --csv` here; no GPU needed)
With outdir, writes outdir/ncu_<kernel>_<tag>.txt per kernel (tag = report basename) and a JSON with the roofline counters."""
import collections, csv, io, json, os, re, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_blocks", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max",
    "sm__cycles_elapsed.max", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_red.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")

def main():
    rep = sys.argv[1]; outdir = sys.argv[2] if len(sys.argv) > 2 else None
    tag = os.path.splitext(os.path.basename(rep))[0]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]; units = rows[1]
    ik = hdr.index("Kernel Name")
    agg = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "").strip()
        agg.setdefault(name, []).append(r)
    counters = {}
    for name, rs in agg.items():
        lines = [f"== {name}   ({len(rs)} launch(es) in the capture; values of the LAST one)"]
        r = rs[-1]
        vals = {}
        for i, h in enumerate(hdr):
            vals[h] = (r[i], units[i])
        for k in KEYS:
            if k in vals:
                lines.append(f"   {k} = {vals[k][0]} {vals[k][1]}")
        stalls = []
        for h, (v, u) in vals.items():
            m = STALL.match(h)
            if m and "per_issue_active" in h:
                try:
                    stalls.append((float(v.replace(",", "")), m.group(1) or m.group(2)))
                except ValueError:
                    pass
        for v, s in sorted(stalls, reverse=True)[:6]:
            lines.append(f"   stall.{s} = {v:.3f} per issue")
        txt = "\n".join(lines)
        print(txt)
        def num(k):
            try:
                v, u = vals[k]; x = float(v.replace(",", ""))
                return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
            except Exception:
                return None
        short = re.sub(r"_kernel.*", "", name)
        short = short.replace("composite_tile_", "composite_")
        dr, dw = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
        counters[short] = {"kernel": name, "dram_bytes_per_launch": (dr or 0) + (dw or 0), "l2_red_sectors": num("lts__t_sectors_srcunit_tex_op_red.sum"),
                           "lanes_active": num("smsp__thread_inst_executed_per_inst_executed.ratio"), "inst_executed": num("smsp__inst_executed.sum"),
                           "duration_us": num("gpu__time_duration.sum")}
        if outdir:
            os.makedirs(outdir, exist_ok=True)
            fn = os.path.join(outdir, f"ncu_{re.sub(r'[^A-Za-z0-9_]+', '_', short)}_{tag}.txt")
            open(fn, "w").write(f"# ncu --set full --clock-control none  ({tag})\n" + txt + "\n")
    if outdir:
        json.dump(counters, open(os.path.join(outdir, f"ncu_counters_{tag}.json"), "w"), indent=1)
        if len(sys.argv) > 3:      # frames per launch of the captured command: also refresh profiles/ncu_traffic.json for bench.py
            fpl = int(sys.argv[3])
            alias = {"onesweep_pass": "onesweep_passes"}
            traffic = {alias.get(k, k): dict(v, frames_per_launch=fpl, capture=tag) for k, v in counters.items()}
            json.dump(traffic, open(os.path.join(outdir, "ncu_traffic.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
