mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-kernel-roofline --no-cpu-baseline --no-optimizer > gpurun_out/r02_bench_clocks_check.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_clocks_check.log
python - <<PY
import json
for l in open('gpurun_out/r02_bench_clocks_check.log'):
    if l.startswith('{'):
        d=json.loads(l); print('p=0', round(d['value'],1), round(d['ms_per_step'],2), d['clocks'])
PY
tail -2 gpurun_out/r02_bench_clocks_check.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --dropout --no-kernel-roofline --no-cpu-baseline --no-optimizer > gpurun_out/r02_bench_dropout.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_dropout.log
python - <<PY
import json
for l in open('gpurun_out/r02_bench_dropout.log'):
    if l.startswith('{'):
        d=json.loads(l); print('dropout', round(d['value'],1), round(d['ms_per_step'],2), d['clocks'], d['config'].get('dropout'))
PY
tail -2 gpurun_out/r02_bench_dropout.log | cut -c1-300
