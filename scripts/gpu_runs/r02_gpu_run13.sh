mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/r02_tests_gpu_13.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_gpu_13.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/r02_tests_gpu_13.log | cut -c1-300 | tail -8
timeout 300 python scripts/attn_time.py 2>&1 | tee gpurun_out/r02_attn_time_13.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 12 -c 12 -o gpurun_out/r02_prof_gemm_layer -f python bench.py --gemm-layer-once > gpurun_out/r02_ncu_gemm_layer.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_13.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_13.log
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_13.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['scoring']['value'], d['cpu_baseline'])
        for k in d['roofline']['kernels']: print('  ', k['case'], round(k['us'],1), round(k['frac'],3))
PY
