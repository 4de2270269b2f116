# 2 GPUs: what the data-parallel overhead consists of -- no exchange / gather+scatter only / full / fewer NCCL CTAs / fp32
mkdir -p gpurun_out
run() {  # name, env, flags
timeout 600 env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-roofline --no-optimizer $3 > gpurun_out/r02_bench_cfg2_n2_$1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_cfg2_n2_$1.log
echo "variant $1"; grep '^{' gpurun_out/r02_bench_cfg2_n2_$1.log | cut -c1-200; tail -1 gpurun_out/r02_bench_cfg2_n2_$1.log
}
run noallreduce X=1 --no-allreduce
run copiesonly MODCR_DEBUG_SKIP_COLLECTIVE=1 ""
run full X=1 ""
run maxctas2 NCCL_MAX_CTAS=2 ""
run fp32 X=1 --fp32-allreduce
timeout 600 python bench.py --steps 20 --warmup 5 --no-kernel-roofline --no-optimizer --no-cpu-baseline > gpurun_out/r02_bench_cfg2_n1_2gpubox.log 2>&1
echo "N=1 same box"; grep '^{' gpurun_out/r02_bench_cfg2_n1_2gpubox.log | cut -c1-200
