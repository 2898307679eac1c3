#!/bin/bash
# Multi-GPU evidence for profiles/: usage gpu_multi_evidence.sh N   (run under gpurun --gpus N)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
nproc > gpurun_out/r2_box_${N}gpu.txt
QPB_DIST_MODES=peer timeout 300 $TR --master-port 29601 scripts/dist_check.py 0.01 0 > gpurun_out/r2_dist_check_${N}gpu.log 2>&1; echo "dist_check rc=$?"
QPB200_TIMING=1 timeout 300 $TR --master-port 29602 bench.py --gpus $N --no-parity > gpurun_out/r2_bench_cfg5_${N}gpu.json 2> gpurun_out/r2_bench_cfg5_${N}gpu.err; echo "cfg5 rc=$?"
grep -E "peer rank 0\]" gpurun_out/r2_bench_cfg5_${N}gpu.err | tail -1 > gpurun_out/r2_dist${N}_phase_timers.txt
timeout 300 $TR --master-port 29603 bench.py --gpus $N --workload cfg3 > gpurun_out/r2_bench_cfg3_${N}gpu.json 2> gpurun_out/r2_bench_cfg3_${N}gpu.err; echo "cfg3 rc=$?"
timeout 200 $TR --master-port 29604 bench.py --gpus $N --workload cfg3shared > gpurun_out/r2_bench_cfg3shared_${N}gpu.json 2> gpurun_out/r2_bench_cfg3shared_${N}gpu.err; echo "cfg3shared rc=$?"
python - <<PY
import json
for w in ("cfg5", "cfg3", "cfg3shared"):
    try:
        d = json.load(open(f"gpurun_out/r2_bench_{w}_${N}gpu.json"))
        print(w, "value", round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 1))
    except Exception as e:
        print(w, "no line:", e)
PY
tail -2 gpurun_out/r2_dist_check_${N}gpu.log | cut -c1-300
cat gpurun_out/r2_dist${N}_phase_timers.txt
