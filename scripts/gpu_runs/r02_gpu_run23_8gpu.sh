# 8 GPUs: the data-parallel overhead follows the time the collectives' kernels are resident (profiles/r02_comm_overlap_*.txt),
# so: fewer, larger collectives / more CTAs per collective
mkdir -p gpurun_out
run() {  # name, env, flags
timeout 600 env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-kernel-roofline --no-optimizer $3 > gpurun_out/r02_bench_cfg2_n8_$1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_cfg2_n8_$1.log
echo "variant $1"; grep '^{' gpurun_out/r02_bench_cfg2_n8_$1.log | cut -c1-200; tail -1 gpurun_out/r02_bench_cfg2_n8_$1.log
}
run bucket64 X=1 "--bucket-mb 64"
run bucket128 X=1 "--bucket-mb 128"
run bucket25_minctas32 NCCL_MIN_CTAS=32 ""
run bucket25 X=1 ""
