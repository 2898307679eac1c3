#!/bin/bash
# Second (last) evidence call of round 2: the whole GPU suite WITHOUT -x on the final tree, then the cfg2 line.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 200 --durations=8 -W default > gpurun_out/r2g_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2g_gpu_tests.log
grep -c "same trajectory" gpurun_out/r2g_gpu_tests.log
timeout 90 python bench.py --workload cfg2 > gpurun_out/r2g_bench_cfg2_1gpu.json 2>/dev/null; echo "cfg2 exit $?"
