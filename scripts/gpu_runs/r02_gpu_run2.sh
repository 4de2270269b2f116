mkdir -p gpurun_out
timeout 300 python scripts/attn_bwd_trace.py > gpurun_out/r02_attn_bwd_trace.log 2>&1; echo "rc=$?" >> gpurun_out/r02_attn_bwd_trace.log
head -150 gpurun_out/r02_attn_bwd_trace.log
