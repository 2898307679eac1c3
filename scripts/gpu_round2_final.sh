#!/bin/bash
# Final evidence of round 2 on one B200 (sized for the ~13 GPU-minutes the round had left): usage gpu_round2_final.sh <git-sha>
# Order = value of the evidence if the box time runs out: tests, smoke, driver-format bench line, slow-test probe,
# ncu DRAM traffic of the timed kernel, ncu launch list.
SHA=$1
mkdir -p gpurun_out
SLOW='tests/test_gpu_direct.py::test_exact_solve_agrees_with_tight_pcg_and_reuses_the_factor'
timeout 420 python -m pytest tests -m gpu -x -q --timeout 200 --durations=8 --deselect "$SLOW" > gpurun_out/r2f_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2f_gpu_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2f_smoke.log
timeout 300 python bench.py > gpurun_out/r2f_bench_cfg5_1gpu.json 2> gpurun_out/r2f_bench_cfg5_1gpu.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2f_bench_cfg5_1gpu.json"))
    print("cfg5 value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "parity", (d.get("parity") or {}).get("pass"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d.get("clocks"))
except Exception as e:
    print("no bench line", e)
PY
timeout 110 python scripts/gpu_slow_test_probe.py > gpurun_out/r2f_slow_test_probe.log 2>&1; echo "probe exit $?"; tail -6 gpurun_out/r2f_slow_test_probe.log
timeout 200 python scripts/ncu_traffic.py $SHA cfg5 > gpurun_out/ncu_traffic.log 2>&1; echo "traffic exit $?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches_bench_cfg5_gpu_time.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-parity > /dev/null 2>&1; echo "launchlist exit $?"
