mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_gpu.log
grep -E "ensemble B=|passed|failed|Error|rc=" gpurun_out/tests_gpu.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
grep -v Warning gpurun_out/bench.log | tail -2 | cut -c1-300
grep -o '"e2e": {[^}]*}' gpurun_out/bench.log
export WARM=1 STEPS=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv python scripts/quick_bench.py > gpurun_out/ncu_launch.log 2>&1
