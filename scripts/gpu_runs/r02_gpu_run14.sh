mkdir -p gpurun_out
for wl in cfg3 cfg5; do
timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-optimizer --no-cpu-baseline --no-kernel-roofline > gpurun_out/r02_bench_${wl}_n1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_${wl}_n1.log
grep -v Warning gpurun_out/r02_bench_${wl}_n1.log | tail -2 | cut -c1-260
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_14.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_14.log
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_14.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['dominant_kernel']['achieved'])
        for k in d['roofline']['kernels']: print('  ', k['case'], round(k['us'],1), round(k['frac'],3))
PY
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg4.log 2>&1
tail -1 gpurun_out/r02_bench_cfg4.log | cut -c1-300
