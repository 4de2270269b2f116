mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_gpu.log
grep -E "losses|passed|failed|FAILED|Error|rc=|Abstract|^E " gpurun_out/tests_gpu.log | tail -30
true
