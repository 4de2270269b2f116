mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -s --timeout 800 -p no:cacheprovider -k model > gpurun_out/r02_tests_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_2gpu.log
grep -o "a parameter (shape[^)]*)[^)]*)" gpurun_out/r02_tests_2gpu.log | head -3; tail -3 gpurun_out/r02_tests_2gpu.log | cut -c1-600
