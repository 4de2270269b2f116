# Re-entry check: GPU parity tests, smoke, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_gpu.log
grep -E "passed|failed|FAILED|Error|rc=|^E " gpurun_out/tests_gpu.log | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
grep -v Warning gpurun_out/bench.log | tail -2 | cut -c1-600
