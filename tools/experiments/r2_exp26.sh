set -x
timeout 900 python -m pytest tests/test_gpu_segmentation.py -m gpu -x -q --tb=short -p no:cacheprovider 2>&1 | tail -4
CHUNKS=0 TIMELINE=0 BATCH=32 python tools/seg_only.py 2>&1 | tail -20
HIPPO_FRAMES_UNFUSED=1 CHUNKS=0 python tools/seg_only.py 2>&1 | grep seg_only
CUDA_LAUNCH_BLOCKING=1 CHUNKS=0 python tools/seg_only.py 2>&1 | grep "overlapped"
