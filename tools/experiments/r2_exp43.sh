# timelines of the single-stream pipeline, cpl 4 / 7 (band 56 / 28)
set -u
for cfg in "4 56" "7 56" "7 28"; do set -- $cfg
  echo "== cpl $1 band $2"
  for i in 1 2 3; do HIPPO_SSIM_CPL=$1 HIPPO_SSIM_BAND7=$2 TIMELINE=444 timeout 300 python tools/seg_only.py 2>&1 | grep -E "pattern\]|overlapped" | tr '\n' ';'; echo; done
done
