# cpl 7: CTAs per SM (5 at 96 registers / 6 at 80) x band height
set -u
mkdir -p gpurun_out
for ctas in 5 6; do for band in 56 28 19; do
  echo "== cpl 7 ctas $ctas band $band"; HIPPO_SSIM7_CTAS=$ctas HIPPO_SSIM_BAND7=$band BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "overlapped|stages, 3|pipeline, 2|digest"
done; done
