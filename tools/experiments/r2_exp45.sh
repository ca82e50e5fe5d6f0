# persisting-L2 window over the 512-sample sums for the pyramid kernels and the chain
set -u
for pin in 1 0; do
  echo "== L2 pin $pin"
  for i in 1 2; do HIPPO_PATTERN_L2PIN=$pin TIMELINE=444 timeout 300 python tools/seg_only.py 2>&1 | grep -E "pattern\]|overlapped" | tr '\n' ';'; echo; done
done
