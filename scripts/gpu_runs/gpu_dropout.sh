# dropout-mode bench (training regime of the reference: hidden 0.3 / attention 0.1) + its ncu launch list
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --dropout --no-cpu-baseline --no-kernel-roofline > gpurun_out/bench_dropout.log 2>&1; echo "rc=$?" >> gpurun_out/bench_dropout.log
grep -v Warning gpurun_out/bench_dropout.log | tail -2 | cut -c1-260
bash scripts/gpu_runs/gpu_launches_dropout.sh
python scripts/summarize_launches.py gpurun_out/launches_dropout.csv > gpurun_out/launches_dropout_summary.txt 2>&1
head -40 gpurun_out/launches_dropout_summary.txt
true
