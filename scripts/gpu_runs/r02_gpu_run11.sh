mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/r02_tests_gpu_11.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_gpu_11.log
grep -E "passed|failed|FAILED|Error|rc=|gtToken" gpurun_out/r02_tests_gpu_11.log | cut -c1-500 | tail -20
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-optimizer --no-cpu-baseline > gpurun_out/r02_bench_11.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_11.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['scoring']['value'])
        for k in d['roofline']['kernels']: print('  ', k['case'], round(k['us'],1), round(k['frac'],3))
PY
export WARM=1 STEPS=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_11.csv python scripts/quick_bench.py > gpurun_out/r02_ncu_launch_11.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_11.csv 2>/dev/null | head -32
