# 4 GPUs: the driver's N=4 scaling point with default flags
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_final_cfg2_n4.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_final_cfg2_n4.log
grep '^{' gpurun_out/r02_bench_final_cfg2_n4.log | cut -c1-220; tail -1 gpurun_out/r02_bench_final_cfg2_n4.log
