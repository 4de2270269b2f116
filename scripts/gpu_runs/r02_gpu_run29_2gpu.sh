# 2 GPUs: even larger gradient buckets (512 MB; 1024 MB = one all-reduce of everything at the end of backward)
mkdir -p gpurun_out
run() {  # name, flags
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-roofline --no-optimizer $2 > gpurun_out/r02_bench_cfg2_n2_$1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_cfg2_n2_$1.log
echo "variant $1"; grep '^{' gpurun_out/r02_bench_cfg2_n2_$1.log | cut -c1-200; tail -1 gpurun_out/r02_bench_cfg2_n2_$1.log
}
run bucket1024 "--bucket-mb 1024"
run bucket512 "--bucket-mb 512"
run bucket128 ""
