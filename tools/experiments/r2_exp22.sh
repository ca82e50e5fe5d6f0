set -x
TIMELINE=444 python tools/seg_only.py 2>&1 | tail -60
