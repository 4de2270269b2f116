mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" --timeout 600 -p no:cacheprovider > gpurun_out/r02_tests_attn_3.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_attn_3.log
tail -5 gpurun_out/r02_tests_attn_3.log
timeout 300 python scripts/attn_bwd_trace.py > gpurun_out/r02_attn_bwd_trace_3.log 2>&1; echo "rc=$?" >> gpurun_out/r02_attn_bwd_trace_3.log
head -70 gpurun_out/r02_attn_bwd_trace_3.log
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg4_3.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_cfg4_3.log'):
    if l.startswith('{'):
        for r in json.loads(l)['sweep']: print({k: round(v,1) for k,v in r.items()})
PY
STEPS=5 WARM=3 timeout 600 python scripts/quick_bench.py 2>&1 | tail -1
