# 8 GPUs: weak-scaling bench of the three model workloads (per-layer gradient hand-off, bf16 buckets)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for wl in cfg2 cfg3 cfg5; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload $wl --steps 10 --warmup 3 --no-kernel-roofline --no-optimizer > gpurun_out/r02_bench_${wl}_n8.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_${wl}_n8.log
grep -v Warning gpurun_out/r02_bench_${wl}_n8.log | tail -2 | cut -c1-260
done
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-kernel-roofline --no-optimizer --no-cpu-baseline > gpurun_out/r02_bench_cfg2_n1_8gpubox.log 2>&1
grep -v Warning gpurun_out/r02_bench_cfg2_n1_8gpubox.log | tail -1 | cut -c1-200
