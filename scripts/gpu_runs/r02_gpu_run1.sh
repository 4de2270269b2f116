# round 2, run 1: all GPU parity tests (new goldens), default bench, cfg-4 sweep, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/r02_tests_gpu_1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_gpu_1.log
grep -E "margins|agreement|passed|failed|FAILED|Error|rc=" gpurun_out/r02_tests_gpu_1.log | tail -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_1.log
grep -v Warning gpurun_out/r02_bench_1.log | tail -2 | cut -c1-600
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg4_1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_cfg4_1.log
tail -2 gpurun_out/r02_bench_cfg4_1.log | cut -c1-1500
export WARM=1 STEPS=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_1.csv python scripts/quick_bench.py > gpurun_out/r02_ncu_launch_1.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_1.csv | head -40
