# 2 GPUs: NCCL reducer tests (incl. windows) + bench with / without the attention-backward comm windows
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -x --timeout 800 -p no:cacheprovider -s > gpurun_out/r02_tests_2gpu_b.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_2gpu_b.log
tail -5 gpurun_out/r02_tests_2gpu_b.log | cut -c1-600
for v in "--comm-windows" ""; do
name=$(echo "x$v" | tr -d ' -')
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-roofline --no-optimizer $v > gpurun_out/r02_bench_cfg2_n2_$name.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_cfg2_n2_$name.log
echo "variant [$v]"; grep '^{' gpurun_out/r02_bench_cfg2_n2_$name.log | cut -c1-200; tail -1 gpurun_out/r02_bench_cfg2_n2_$name.log
done
