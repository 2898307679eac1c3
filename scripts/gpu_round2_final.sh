#!/bin/bash
# Final evidence of round 2 on one B200: usage gpu_round2_final.sh <git-sha>
SHA=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 --durations=10 > gpurun_out/r2f_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2f_gpu_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2f_smoke.log
timeout 400 python bench.py > gpurun_out/r2f_bench_cfg5_1gpu.json 2> gpurun_out/r2f_bench_cfg5_1gpu.err; echo "bench exit $?"
timeout 200 python bench.py --workload cfg2 > gpurun_out/r2f_bench_cfg2_1gpu.json 2>/dev/null; echo "cfg2 exit $?"
timeout 200 python bench.py --workload cfg4 > gpurun_out/r2f_bench_cfg4_1gpu.json 2>/dev/null; echo "cfg4 exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches_bench_cfg5_gpu_time.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-parity > /dev/null 2>&1; echo "launchlist exit $?"
timeout 400 python scripts/ncu_traffic.py $SHA cfg5 > gpurun_out/ncu_traffic.log 2>&1; echo "traffic exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_shared_kernel -c 1 -o gpurun_out/r2f_dense_shared python bench.py --workload cfg3shared --batch 16384 --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1; echo "ncu shared exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gj_update_kernel --launch-skip 150 -c 2 -o gpurun_out/r2f_gj_update python scripts/profile_direct.py > gpurun_out/ncu_direct.log 2>&1; echo "ncu direct exit $?"
python - <<PY
import json
for w in ("cfg5", "cfg2", "cfg4"):
    try:
        d = json.load(open(f"gpurun_out/r2f_bench_{w}_1gpu.json"))
        print(w, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"].get("traffic"), "parity", (d.get("parity") or {}).get("pass"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(w, "no line", e)
PY
