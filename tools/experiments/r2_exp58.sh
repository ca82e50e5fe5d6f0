# batch of streams: gray conversion of video i + 1 issued BEFORE the SSIM kernel of video i, as a thin grid
set -u
timeout 600 python -m pytest tests/test_gpu_segmentation.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
for i in 1 2; do GRAYCAP=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "pipelined issue" | sed 's/\[seg_only\] batch of 32, pipelined issue, //' | tr '\n' ';'; echo; done
