# 8 GPUs: SM budget left to the NCCL CTAs while gradient buckets overlap backward (MODCR_COMM_SMS)
mkdir -p gpurun_out
for c in 8 16 4; do
MODCR_COMM_SMS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-kernel-roofline --no-optimizer > gpurun_out/r02_bench_cfg2_n8_comm$c.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_cfg2_n8_comm$c.log
echo "COMM_SMS=$c"; grep -v Warning gpurun_out/r02_bench_cfg2_n8_comm$c.log | tail -2 | cut -c1-200
done
