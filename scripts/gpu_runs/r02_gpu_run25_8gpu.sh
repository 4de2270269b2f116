# 8 GPUs, final tree (128 MB buckets by default): cfg-2 / cfg-3 / cfg-5 weak scaling + single-GPU lines on the same box
mkdir -p gpurun_out
run() {  # name, flags
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-kernel-roofline --no-optimizer $2 > gpurun_out/r02_bench_final_$1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_final_$1.log
echo "variant $1"; grep '^{' gpurun_out/r02_bench_final_$1.log | cut -c1-200; tail -1 gpurun_out/r02_bench_final_$1.log
}
run cfg2_n8 ""
run cfg2_n8_bucket256 "--bucket-mb 256"
run cfg3_n8 "--workload cfg3"
run cfg5_n8 "--workload cfg5"
for wl in cfg2 cfg3 cfg5; do
timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-kernel-roofline --no-optimizer --no-cpu-baseline > gpurun_out/r02_bench_final_${wl}_n1_8gpubox.log 2>&1
echo "N=1 $wl"; grep '^{' gpurun_out/r02_bench_final_${wl}_n1_8gpubox.log | cut -c1-200
done
