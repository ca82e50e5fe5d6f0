# chain released by a certain audio hit (no wait for pairs the audio boundary overrides) + cpl 7 default at 224 px
set -u
for cpl in 7 4; do
  echo "== cpl $cpl"
  for i in 1 2; do HIPPO_SSIM_CPL=$cpl TIMELINE=444 BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "pattern\]|overlapped|stages, 3|pipeline, 2|digest" | tr '\n' ';'; echo; done
done
echo "== tests (default layout)"; timeout 900 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py tests/test_gpu_fullsize.py -m gpu -q -x --tb=short -p no:cacheprovider -k "not consol and not search and not bank" 2>&1 | tail -5
echo "== tests (cpl 4 forced)"; HIPPO_SSIM_CPL=4 timeout 900 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
echo "== tests (cpl 7 forced)"; HIPPO_SSIM_CPL=7 timeout 900 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
