mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_gpu.log
grep -E "losses|passed|failed|FAILED|Error|rc=" gpurun_out/tests_gpu.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
grep -v Warning gpurun_out/bench.log | tail -2 | cut -c1-260
grep -o '"optimizer_step": {[^}]*}' gpurun_out/bench.log
grep -i "optimiser section failed" gpurun_out/bench.log
