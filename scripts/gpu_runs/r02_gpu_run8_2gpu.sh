# 2 GPUs: NCCL reducer tests (fp32 / bf16 buckets, per-layer hand-off on the real model) and the N=2 bench
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -s --timeout 800 -p no:cacheprovider > gpurun_out/r02_tests_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_2gpu.log
tail -12 gpurun_out/r02_tests_2gpu.log | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-kernel-roofline --no-optimizer > gpurun_out/r02_bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_n2.log
grep -v Warning gpurun_out/r02_bench_n2.log | tail -3 | cut -c1-400
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-kernel-roofline --no-optimizer --no-cpu-baseline > gpurun_out/r02_bench_n1_samebox.log 2>&1
grep -v Warning gpurun_out/r02_bench_n1_samebox.log | tail -1 | cut -c1-200
