# does a thin persistent gray grid (1 - 2 CTAs per SM) hide under the SSIM kernel of the previous stream of a batch?
set -u
for occ in 0 1 2 3; do
  echo "== gray CTAs per SM cap $occ"; HIPPO_GRAY_OCC=$occ BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "stages|pipeline, 2|frame pairs" | sed 's/\[seg_only\] //'
done
