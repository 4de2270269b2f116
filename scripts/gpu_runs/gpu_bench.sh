mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -s --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/t3.log 2>&1; echo "rc=$?" >> gpurun_out/t3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
grep -E "^(ragged|cfg1)|passed|failed|Error" gpurun_out/t3.log | cut -c1-600; tail -3 gpurun_out/smoke.log; tail -4 gpurun_out/bench.log
