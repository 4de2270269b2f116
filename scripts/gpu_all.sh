mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 240 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/t1.log 2>&1; echo "rc=$?" >> gpurun_out/t1.log
timeout 900 python -m pytest tests/test_model_gpu.py -q -s --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/t3.log 2>&1; echo "rc=$?" >> gpurun_out/t3.log
PROFILE=1 timeout 600 python scripts/quick_bench.py > gpurun_out/qb.log 2>&1
tail -5 gpurun_out/t1.log; grep -E "^(ragged|cfg1)|passed|failed|Error" gpurun_out/t3.log | cut -c1-900; head -60 gpurun_out/qb.log | cut -c1-200
