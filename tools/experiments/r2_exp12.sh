# round 2, call 12: SSIM chunks static vs persistent, fine tail bands
set -u
timeout 300 python -m pytest tests/test_gpu_segmentation.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
for cfg in "1 28" "1 56" "0 28" "0 56"; do set -- $cfg
  echo "persistent=$1 fine=$2"
  HIPPO_SSIM_PERSISTENT=$1 HIPPO_SSIM_FINE=$2 CHUNKS=444,888 timeout 300 python tools/seg_only.py 2>&1 | grep -E "overlapped"
done
HIPPO_SSIM_PERSISTENT=1 CHUNKS=444 TIMELINE=444 BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "pattern|batch" | head -40
