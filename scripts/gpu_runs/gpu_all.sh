mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 240 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/t1.log 2>&1; echo "rc=$?" >> gpurun_out/t1.log
timeout 900 python -m pytest tests/test_model_gpu.py -q -s --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/t3.log 2>&1; echo "rc=$?" >> gpurun_out/t3.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
tail -3 gpurun_out/t1.log; grep -E "^(ragged|cfg1)|passed|failed|Error" gpurun_out/t3.log | cut -c1-330; grep -v Warning gpurun_out/bench.log | tail -3 | cut -c1-1800
