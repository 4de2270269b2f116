mkdir -p gpurun_out
export WARM=1 STEPS=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python scripts/quick_bench.py > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 12 -c 2 -o gpurun_out/prof_attn_bwd -f python scripts/quick_bench.py > gpurun_out/ncu_attn_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 12 -c 2 -o gpurun_out/prof_attn_fwd -f python scripts/quick_bench.py > gpurun_out/ncu_attn_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 60 -c 9 -o gpurun_out/prof_gemm -f python scripts/quick_bench.py > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ln_bwd_kernel -s 4 -c 1 -o gpurun_out/prof_lnbwd -f python scripts/quick_bench.py > gpurun_out/ncu_lnbwd.log 2>&1
ls -la gpurun_out/ | tail -12; tail -2 gpurun_out/ncu_launch.log
