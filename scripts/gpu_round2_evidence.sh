mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_box.txt; nproc >> gpurun_out/r2_box.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?"
timeout 600 python bench.py > gpurun_out/r2_bench_cfg5_1gpu.json 2> gpurun_out/r2_bench_cfg5_1gpu.err; echo "bench exit $?"
timeout 400 python bench.py --workload banded --no-cpu > gpurun_out/r2_bench_banded_1gpu.json 2> gpurun_out/r2_bench_banded.err; echo "banded exit $?"
timeout 400 python bench.py --workload cfg4 > gpurun_out/r2_bench_cfg4_1gpu.json 2> gpurun_out/r2_bench_cfg4.err; echo "cfg4 exit $?"
timeout 400 python bench.py --workload cfg2 > gpurun_out/r2_bench_cfg2_1gpu.json 2> gpurun_out/r2_bench_cfg2.err; echo "cfg2 exit $?"
timeout 600 python bench.py --workload cfg3 > gpurun_out/r2_bench_cfg3_1gpu.json 2> gpurun_out/r2_bench_cfg3.err; echo "cfg3 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench_cfg5_gpu_time.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-parity > gpurun_out/ncu_launch.log 2>&1; echo "launchlist exit $?"
timeout 900 python scripts/ncu_traffic.py 7508fd2 cfg5 banded > gpurun_out/ncu_traffic.log 2>&1; echo "traffic exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -o gpurun_out/r2_spmv_banded python scripts/profile_spmv.py banded > gpurun_out/ncu_full_banded.log 2>&1; echo "ncu banded exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -o gpurun_out/r2_spmv_cfg5 python scripts/profile_spmv.py cfg5 > gpurun_out/ncu_full_cfg5.log 2>&1; echo "ncu cfg5 exit $?"
tail -3 gpurun_out/r2_gpu_tests.log
