mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention or gemm" --timeout 600 -p no:cacheprovider 2>&1 | tail -2
for i in 1 2; do
timeout 300 python scripts/attn_time.py 2>&1 | tee -a gpurun_out/r02_attn_time_6.log
MODCR_ATTN_BWD_FUSED=0 timeout 300 python scripts/attn_time.py 2>&1 | tee -a gpurun_out/r02_attn_time_6_unfused.log
done
timeout 300 python scripts/attn_bwd_trace.py > gpurun_out/r02_attn_bwd_trace_6.log 2>&1; echo "rc=$?" >> gpurun_out/r02_attn_bwd_trace_6.log
grep -A95 "CTA 21" gpurun_out/r02_attn_bwd_trace_6.log | cut -c1-100
