# 8 GPUs: where the N=8 overhead comes from -- no exchange at all / hooks only / per-layer hand-off
mkdir -p gpurun_out
for v in "--no-allreduce" "--no-handoff" ""; do
name=$(echo "x$v" | tr -d ' -')
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-kernel-roofline --no-optimizer $v > gpurun_out/r02_bench_cfg2_n8_$name.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_cfg2_n8_$name.log
echo "variant [$v]"; grep -v Warning gpurun_out/r02_bench_cfg2_n8_$name.log | grep '^{' | cut -c1-200
done
