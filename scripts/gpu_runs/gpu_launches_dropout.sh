mkdir -p gpurun_out
export WARM=1 STEPS=1 P_HIDDEN=0.3 P_ATTN=0.1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_dropout.csv python scripts/quick_bench.py > gpurun_out/ncu_launch_dropout.log 2>&1
tail -2 gpurun_out/ncu_launch_dropout.log
