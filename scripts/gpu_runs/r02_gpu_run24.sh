# 1 GPU, final tree: full GPU test suite, smoke, default bench (with the CPU baseline), reference arm, ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_tests_gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_gpu_final.log
tail -3 gpurun_out/r02_tests_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_smoke_final.log; tail -2 gpurun_out/r02_smoke_final.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r02_bench_final_cfg2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_final_cfg2.log
python - <<PY
import json
for l in open('gpurun_out/r02_bench_final_cfg2.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N=1', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline'])
        for k in d['roofline']['kernels']: print('  ', k['case'], round(k['us'],1), round(k['frac'],3), k.get('bound'), round(k.get('frac_of_floor',0),3))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_reference_arm.log; grep '^{' gpurun_out/r02_bench_reference_arm.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-graph --no-kernel-roofline --no-optimizer --no-cpu-baseline > gpurun_out/r02_ncu_launch_final.log 2>&1; echo "ncu rc=$?"
