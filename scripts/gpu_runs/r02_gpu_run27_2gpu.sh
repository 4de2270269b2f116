# 2 GPUs: NCCL protocol of the gradient all-reduce (default picks RING_LL for these bucket sizes) + the NCCL reducer tests
mkdir -p gpurun_out
run() {  # name, env, flags
timeout 600 env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-roofline --no-optimizer $3 > gpurun_out/r02_bench_cfg2_n2_$1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_cfg2_n2_$1.log
echo "variant $1"; grep '^{' gpurun_out/r02_bench_cfg2_n2_$1.log | cut -c1-200; tail -1 gpurun_out/r02_bench_cfg2_n2_$1.log
}
run proto_simple NCCL_PROTO=Simple ""
run proto_ll128 NCCL_PROTO=LL128 ""
run proto_default X=1 ""
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -x --timeout 800 -p no:cacheprovider -s > gpurun_out/r02_tests_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_2gpu.log
tail -4 gpurun_out/r02_tests_2gpu.log | cut -c1-400
