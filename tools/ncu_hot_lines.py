"""Top SASS instructions (and, when the report carries them, source lines) of one kernel in an ncu report by warp-stall
samples.  usage: python tools/ncu_hot_lines.py <report.ncu-rep> [kernel-regex] [N]   (runs `ncu -i ... --page source --csv`)"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else None; N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if kern: cmd += ["-k", "regex:" + kern]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next((r for r in rows if "Source" in r), None)
if hdr is None:
    print("no 'Source' header; first lines:\n" + raw[:4000]); sys.exit(0)
i0 = rows.index(hdr)
print("columns:", hdr)
def col(pat):
    for i, h in enumerate(hdr):
        if re.search(pat, h, re.I): return i
    return None
c_src = hdr.index("Source")
c_smp = col(r"sampling.*\(all") or col(r"sampling")
c_ins = col(r"instructions executed")
body = [r for r in rows[i0 + 1:] if len(r) == len(hdr)]
if c_smp is None:
    print("no sampling column; first rows:"); [print(r) for r in body[:40]]; sys.exit(0)
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
tot = sum(num(r[c_smp]) for r in body) or 1.0
print(f"{len(body)} rows, {tot:.0f} samples")
# program order with cumulative share helps to see which phase is hot
order = sorted(range(len(body)), key=lambda i: -num(body[i][c_smp]))[:N]
for i in sorted(order):
    r = body[i]
    print(f"row {i:5d} {100 * num(r[c_smp]) / tot:5.1f}%  inst {r[c_ins] if c_ins is not None else '-':>10}  {r[c_src].strip()[:140]}")
