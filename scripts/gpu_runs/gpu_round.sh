# One GPU-box visit: parity tests, smoke, bench, ncu launch list, ncu full captures of the top kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
export WARM=1 STEPS=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv python scripts/quick_bench.py > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 12 -c 1 -o gpurun_out/prof_attn_bwd -f python scripts/quick_bench.py > gpurun_out/ncu_attn_bwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 12 -c 1 -o gpurun_out/prof_attn_fwd -f python scripts/quick_bench.py > gpurun_out/ncu_attn_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 60 -c 9 -o gpurun_out/prof_gemm -f python scripts/quick_bench.py > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/tests_gpu.log; tail -2 gpurun_out/smoke.log; grep -v Warning gpurun_out/bench.log | tail -2 | cut -c1-2500; ls -la gpurun_out | tail
