# parity tests, bench, launch list, ncu full on forward and backward GEMMs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
grep -v Warning gpurun_out/bench.log | tail -2 | cut -c1-300
export WARM=1 STEPS=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv python scripts/quick_bench.py > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 60 -c 6 -o gpurun_out/prof_gemm_fwd -f python scripts/quick_bench.py > gpurun_out/ncu_gemm_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 520 -c 10 -o gpurun_out/prof_gemm_bwd -f python scripts/quick_bench.py > gpurun_out/ncu_gemm_bwd.log 2>&1
ls -la gpurun_out | tail -8
