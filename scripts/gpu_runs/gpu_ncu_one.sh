mkdir -p gpurun_out
export WARM=1 STEPS=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_persist -s 21 -c 1 -o gpurun_out/prof_attn_fwdp2 -f python scripts/quick_bench.py > gpurun_out/ncu_attn_fwdp2.log 2>&1
tail -2 gpurun_out/ncu_attn_fwdp2.log
