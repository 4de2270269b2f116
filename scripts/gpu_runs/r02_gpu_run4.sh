mkdir -p gpurun_out
timeout 300 python scripts/attn_bwd_trace.py > gpurun_out/r02_attn_bwd_trace_4.log 2>&1; echo "rc=$?" >> gpurun_out/r02_attn_bwd_trace_4.log
cat gpurun_out/r02_attn_bwd_trace_4.log | cut -c1-100
