mkdir -p gpurun_out
timeout 300 python scripts/attn_time.py 2>&1 | tee gpurun_out/r02_attn_time_12.log
export WARM=1 STEPS=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 60 -c 12 -o gpurun_out/r02_prof_gemm_fwd -f python scripts/quick_bench.py > gpurun_out/r02_ncu_gemm_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 520 -c 12 -o gpurun_out/r02_prof_gemm_bwd -f python scripts/quick_bench.py > gpurun_out/r02_ncu_gemm_bwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 12 -c 2 -o gpurun_out/r02_prof_attn_bwd -f python scripts/quick_bench.py > gpurun_out/r02_ncu_attn_bwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_persist -s 12 -c 2 -o gpurun_out/r02_prof_attn_fwd -f python scripts/quick_bench.py > gpurun_out/r02_ncu_attn_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:attn_fwd_persist_kernel<2>|attn_fwd_persist_kernelILi2" -s 0 -c 1 -o gpurun_out/r02_prof_attn_fwd2 -f python scripts/quick_bench.py > gpurun_out/r02_ncu_attn_fwd2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_ -s 8 -c 4 -o gpurun_out/r02_prof_ln -f python scripts/quick_bench.py > gpurun_out/r02_ncu_ln.log 2>&1
ls -la gpurun_out/*.ncu-rep
