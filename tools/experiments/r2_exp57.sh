# batch of streams with one grow-only arena for SSIM values + audio pyramids (no allocator churn)
set -u
timeout 600 python -m pytest tests/test_gpu_segmentation.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
for i in 1 2 3; do BATCH=32 HOSTTIME=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "stages|pipeline|host issue" | sed 's/\[seg_only\] //' | tr '\n' ';'; echo; done
