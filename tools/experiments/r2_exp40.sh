# SSIM row step with running pointers + dp2a products (213 -> 173 instructions): A/B against the previous kernel
# (tools/probes/libhippo_old_ssim.so = HEAD's frames.cu), bit-identity by digest, then the segmentation tests.
set -u
mkdir -p gpurun_out
cp hippomm_b200/libhippo_b200.so /tmp/new.so
echo "== new"; TIMELINE=444 BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | tail -12
cp tools/probes/libhippo_old_ssim.so hippomm_b200/libhippo_b200.so
echo "== old"; BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | tail -9
cp /tmp/new.so hippomm_b200/libhippo_b200.so
echo "== new again"; timeout 300 python tools/seg_only.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py tests/test_gpu_fullsize.py -m gpu -q -x --tb=short -p no:cacheprovider -k "not consol and not search and not bank" 2>&1 | tail -5
