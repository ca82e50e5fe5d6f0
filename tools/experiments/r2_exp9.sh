# round 2, call 9: is the slow 888 / 1200 chunk timing reproducible?
set -u
CHUNKS=888,1200,888,444,600,1776,888 timeout 300 python tools/seg_only.py 2>&1 | grep -E "seg_only" | head -20
