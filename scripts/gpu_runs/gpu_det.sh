for i in 1 2 3; do timeout 300 python -m pytest tests/test_model_gpu.py -q -x -k full_size -p no:cacheprovider 2>&1 | grep -E "passed|failed|test_model_gpu.py:[0-9]+: Assert" ; done
