# with the L2 window: cpl 7 band height against the chain's tail; cpl 4 for reference
set -u
for cfg in "7 56" "7 28" "7 19" "7 14" "4 56"; do set -- $cfg
  echo "== cpl $1 band $2"
  for i in 1 2; do HIPPO_SSIM_CPL=$1 HIPPO_SSIM_BAND7=$2 TIMELINE=444 BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "follow end|pairs end|overlapped|stages, 3|pipeline, 2|frame pairs" | sed 's/\[seg_only\] //; s/\[pattern\] *//; s/segments 160 digest [0-9a-f]* //' | tr '\n' ';'; echo; done
done
