# 2 GPUs: CUPTI timeline of the data-parallel step (fill / windows) and of the single-GPU step on the same box
mkdir -p gpurun_out
run() {  # name, env, flags
timeout 600 env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-kernel-roofline --no-optimizer --trace-out gpurun_out/r02_trace_n2_$1 $3 > gpurun_out/r02_bench_trace_n2_$1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_trace_n2_$1.log
echo "variant $1"; grep '^{' gpurun_out/r02_bench_trace_n2_$1.log | cut -c1-200; tail -1 gpurun_out/r02_bench_trace_n2_$1.log
}
run fill X=1 ""
run windows X=1 --comm-windows
timeout 600 python bench.py --steps 10 --warmup 3 --no-kernel-roofline --no-optimizer --no-cpu-baseline --trace-out gpurun_out/r02_trace_n1 > gpurun_out/r02_bench_trace_n1.log 2>&1
grep '^{' gpurun_out/r02_bench_trace_n1.log | cut -c1-200
ls -la gpurun_out/*.json | head; gzip -f gpurun_out/r02_trace_*.json; ls -la gpurun_out/*.gz
