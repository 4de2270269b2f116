mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -s --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/t3.log 2>&1
echo "rc3=$?" >> gpurun_out/t3.log
tail -60 gpurun_out/t3.log
