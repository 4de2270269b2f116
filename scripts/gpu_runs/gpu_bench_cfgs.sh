mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
grep -v Warning gpurun_out/bench.log | tail -2 | cut -c1-6000
timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.log 2>&1; echo "rc=$?" >> gpurun_out/bench_cfg3.log
grep -v Warning gpurun_out/bench_cfg3.log | tail -2 | cut -c1-700
timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.log 2>&1; echo "rc=$?" >> gpurun_out/bench_cfg5.log
grep -v Warning gpurun_out/bench_cfg5.log | tail -2 | cut -c1-700
