# round 2, call 16: store hook tests, exact search after the hippo_rescore signature change
set -u
timeout 600 python -m pytest tests/test_gpu_recall.py tests/test_gpu_search.py tests/test_gpu_fullsize.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -12
