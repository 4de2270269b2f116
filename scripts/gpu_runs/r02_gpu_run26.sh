mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_tests_gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_gpu_final.log
tail -4 gpurun_out/r02_tests_gpu_final.log | cut -c1-300
grep -n "3-D mask grad" gpurun_out/r02_tests_gpu_final.log | head
