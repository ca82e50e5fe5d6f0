# validation of the cpl 7 layout + audio-released chain + L2 window: memcheck over the frame-pair tests (both layouts),
# whole GPU suite, seg_only
set -u
mkdir -p gpurun_out
for cpl in 7 4; do
  echo "== memcheck cpl $cpl"; HIPPO_SSIM_CPL=$cpl timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py -m gpu -q -x -p no:cacheprovider -k "frame or ssim or pair or prefilter or dedup or saved" 2>&1 | tail -4
done
echo "== whole suite"; timeout 1200 python -m pytest tests -m gpu -q --tb=short --timeout 600 --timeout-method=thread -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -12 gpurun_out/pytest_gpu.log
BATCH=32 TIMELINE=444 timeout 300 python tools/seg_only.py 2>&1 | tail -12
