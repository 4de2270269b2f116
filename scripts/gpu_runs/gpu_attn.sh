mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/t_attn.log 2>&1; echo "rc=$?" >> gpurun_out/t_attn.log
tail -5 gpurun_out/t_attn.log
if grep -q "rc=0" gpurun_out/t_attn.log; then
timeout 900 python -m pytest tests/test_model_gpu.py -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/t_model.log 2>&1; echo "rc=$?" >> gpurun_out/t_model.log
tail -3 gpurun_out/t_model.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
grep -v Warning gpurun_out/bench.log | tail -2 | cut -c1-300
export WARM=1 STEPS=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv python scripts/quick_bench.py > gpurun_out/ncu_launch.log 2>&1

fi
