mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "not prefix_attention" --timeout 240 -p no:cacheprovider > gpurun_out/t1.log 2>&1
echo "rc1=$?" >> gpurun_out/t1.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "prefix_attention" --timeout 240 --timeout-method=thread -p no:cacheprovider > gpurun_out/t2.log 2>&1
echo "rc2=$?" >> gpurun_out/t2.log
tail -30 gpurun_out/t1.log; tail -40 gpurun_out/t2.log
