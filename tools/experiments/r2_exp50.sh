set -u
for i in 1 2; do BATCH=32 TIMELINE=444 timeout 300 python tools/seg_only.py 2>&1 | grep -E "follow end|pairs end|overlapped|stages, 3|pipeline, 2|frame pairs" | sed 's/\[seg_only\] //; s/\[pattern\] *//; s/segments 160 digest [0-9a-f]* //' | tr '\n' ';'; echo; done
