mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 900 --timeout-method=thread -p no:cacheprovider > gpurun_out/r02_tests_gpu_10.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_gpu_10.log
grep -E "margins|agreement|passed|failed|FAILED|Error|rc=|worst ours" gpurun_out/r02_tests_gpu_10.log | cut -c1-400 | tail -30
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-optimizer --no-cpu-baseline > gpurun_out/r02_bench_n1_10.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_n1_10.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'])
        for k in d['roofline']['kernels']: print('  ', k['case'], round(k['us'],1), round(k['frac'],3))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-kernel-roofline --no-optimizer > gpurun_out/r02_bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_n2.log
grep -v Warning gpurun_out/r02_bench_n2.log | tail -2 | cut -c1-300
