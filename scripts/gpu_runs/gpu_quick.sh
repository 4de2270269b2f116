# GEMM standalone check + perf, then model parity tests and the bench
mkdir -p gpurun_out
(timeout 120 build/test_gemm; echo rc=$?) > gpurun_out/gemm_check.log 2>&1
(timeout 120 build/test_gemm perf; echo rc=$?) > gpurun_out/gemm_perf_2cta.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
grep -c PASS gpurun_out/gemm_check.log; tail -12 gpurun_out/gemm_perf_2cta.log; tail -3 gpurun_out/tests_gpu.log; grep -v Warning gpurun_out/bench.log | tail -2 | cut -c1-400
