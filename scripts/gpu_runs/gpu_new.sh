# new-feature check: selected GPU tests (pattern in $K), then a short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider -k "$K" -s > gpurun_out/tests_new.log 2>&1; echo "rc=$?" >> gpurun_out/tests_new.log
grep -E "passed|failed|FAILED|Error|rc=|^E |err" gpurun_out/tests_new.log | tail -25
