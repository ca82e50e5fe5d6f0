set -u
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_segmentation.py -m gpu -q --tb=short --timeout 500 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -15
timeout 400 python bench.py --steps 3 2>&1 >/dev/null | grep extra
