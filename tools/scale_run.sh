#!/bin/bash
# Multi-GPU evidence run (under `gpurun --gpus 8`): 2-GPU NCCL gradient-equivalence test, the headline bench on 8 GPUs with
# NCCL's own log of the all-reduce algorithm, and BASELINE config C4 (1 M surfels, 1024^2, 64 frames = 8 per GPU).
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING NCCL_DEBUG_FILE=gpurun_out/nccl_${TAG}_%h_%p.log timeout 500 $RUN bench.py --gpus 8 --steps 10 --warmup 3 \
    > gpurun_out/bench_${TAG}_hl_8gpu.json 2> gpurun_out/bench_${TAG}_hl_8gpu.log
ls gpurun_out/nccl_${TAG}_* 2>/dev/null | head -1 | xargs -I{} sh -c 'grep -E "NVLS|Ring|Tree|Algo|algo|busbw|AllReduce" {} | head -40' > gpurun_out/nccl_${TAG}_summary.txt
rm -f gpurun_out/nccl_${TAG}_*_*.log
timeout 600 $RUN bench.py --gpus 8 --steps 6 --warmup 3 --surfels 1000000 --res 1024 --frames-per-step 8 \
    > gpurun_out/bench_${TAG}_c4_8gpu.json 2> gpurun_out/bench_${TAG}_c4_8gpu.log
python - <<PY
import json
for n in ("hl", "c4"):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}_%s_8gpu.json" % n))
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["config"]["parallelism"], d["clocks"])
    except Exception as ex:
        print(n, "FAILED", ex)
PY
tail -3 gpurun_out/bench_${TAG}_hl_8gpu.log; tail -3 gpurun_out/bench_${TAG}_c4_8gpu.log; wc -l gpurun_out/nccl_${TAG}_summary.txt
