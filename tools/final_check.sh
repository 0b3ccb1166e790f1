#!/bin/bash
# End-of-round GPU check (run under gpurun): parity tests, both bench arms, ncu captures of the final kernels.
# usage: bash tools/final_check.sh <tag>      -> files gpurun_out/*_<tag>.*
TAG=${1:-final}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.log
timeout 300 python bench.py > gpurun_out/bench_${TAG}_ours.json 2> gpurun_out/bench_${TAG}_ours.log
python - <<PY
import json
for a in ("reference", "ours"):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}_%s.json" % a)); r = d.get("roofline") or {}
        print(a, d["value"], d["e2e"]["value"], r.get("kernel"), r.get("frac"), r.get("pair_evals_upper_per_s"), d["clocks"],
              {k: v["ms"] for k, v in (d.get("kernels_ms") or {}).items()})
    except Exception as ex:
        print(a, "FAILED", ex)
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:composite -o gpurun_out/prof_composite_${TAG} \
    python tools/prof_frame.py --frames 1 > gpurun_out/prof_${TAG}.log 2>&1
tail -1 gpurun_out/prof_${TAG}.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_ours_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-ref-cuda --cpu-frames 0 --no-graph > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_ours_${TAG}.csv 2>/dev/null | sed -n 1,12p
