mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" --timeout 600 -p no:cacheprovider 2>&1 | tail -2
for wde in 1 0; do
MODCR_GEMM_WIDE=$wde timeout 600 python bench.py --steps 10 --warmup 3 --no-optimizer --no-cpu-baseline > gpurun_out/r02_bench_wide$wde.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/r02_bench_wide$wde.log'):
    if l.startswith('{'):
        d=json.loads(l); print('WIDE=$wde N=1', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1))
        for k in d['roofline']['kernels'][:4]: print('  ', k['case'], round(k['us'],1), round(k['frac'],3))
PY
done
# compute-sanitizer on the smoke forward+backward: memcheck and synccheck
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -4 gpurun_out/r02_sanitizer_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_synccheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitizer_synccheck.log
tail -4 gpurun_out/r02_sanitizer_synccheck.log | cut -c1-200
