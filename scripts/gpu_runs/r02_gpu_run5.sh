mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" --timeout 600 -p no:cacheprovider 2>&1 | tail -2
timeout 300 python scripts/attn_time.py 2>&1 | tee gpurun_out/r02_attn_time_5.log
MODCR_ATTN_BWD_FUSED=0 timeout 300 python scripts/attn_time.py 2>&1 | tee gpurun_out/r02_attn_time_5_unfused.log
timeout 300 python scripts/attn_bwd_trace.py > gpurun_out/r02_attn_bwd_trace_5.log 2>&1; echo "rc=$?" >> gpurun_out/r02_attn_bwd_trace_5.log
grep -A95 "CTA 21" gpurun_out/r02_attn_bwd_trace_5.log | cut -c1-100
