mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/r02_tests_gpu_7.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_gpu_7.log
grep -E "margins|agreement|passed|failed|FAILED|Error|rc=|Roberta|grad rel-L2" gpurun_out/r02_tests_gpu_7.log | cut -c1-900 | tail -40
timeout 300 python scripts/attn_time.py 2>&1 | tee gpurun_out/r02_attn_time_7.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_7.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_7.log
grep -v Warning gpurun_out/r02_bench_7.log | tail -2 | cut -c1-300
