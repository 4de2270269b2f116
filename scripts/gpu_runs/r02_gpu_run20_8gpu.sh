# 8 GPUs: gradient all-reduce issued beside the attention-backward kernels (--comm-windows) and / or on high-priority
# streams (MODCR_COMM_PRIORITY), against the per-layer hand-off that issues a collective whenever a bucket fills
mkdir -p gpurun_out
run() {  # name, env, flags
timeout 600 env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-kernel-roofline --no-optimizer $3 > gpurun_out/r02_bench_cfg2_n8_$1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_cfg2_n8_$1.log
echo "variant $1"; grep '^{' gpurun_out/r02_bench_cfg2_n8_$1.log | cut -c1-200; tail -1 gpurun_out/r02_bench_cfg2_n8_$1.log
}
run windows_prio MODCR_COMM_PRIORITY=1 --comm-windows
run fill_prio MODCR_COMM_PRIORITY=1 ""
run windows_noprio MODCR_COMM_PRIORITY=0 --comm-windows
run fill_noprio MODCR_COMM_PRIORITY=0 ""
