timeout 600 python -m pytest tests/test_gpu_recall.py tests/test_store.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
bash tools/experiments/r2_exp59.sh 2>&1 | grep "\[extra\] recall"
