# round 2, call 13: SSIM kernel at 4 CTAs per SM (64 registers, no spill after the SsimArgs refactor)
set -u
TAIL=1 CHUNKS=444 BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "seg_only"
