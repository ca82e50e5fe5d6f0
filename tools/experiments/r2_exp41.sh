# SSIM with 7 columns per lane on a 7-in-8 gray layout (one warp per 224-pixel row) against 4 columns per lane
set -u
mkdir -p gpurun_out
for cpl in 7 4; do
  echo "== cpl $cpl"; HIPPO_SSIM_CPL=$cpl BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | tail -9
done
echo "== tests, cpl 7 forced"; HIPPO_SSIM_CPL=7 timeout 900 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py tests/test_gpu_fullsize.py -m gpu -q -x --tb=short -p no:cacheprovider -k "not consol and not search and not bank" 2>&1 | tail -8
echo "== tests, default"; timeout 900 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py tests/test_gpu_fullsize.py -m gpu -q -x --tb=short -p no:cacheprovider -k "not consol and not search and not bank" 2>&1 | tail -5
