# gray kernels: contiguous run of items per warp (no 64-bit division per item, min / max reduced once per frame)
set -u
for i in 1 2; do BATCH=32 TIMELINE=444 timeout 300 python tools/seg_only.py 2>&1 | grep -E "follow end|pairs end|overlapped|stages, 3|frame pairs" | sed 's/\[seg_only\] //; s/\[pattern\] *//; s/segments 160 digest /digest /' | tr '\n' ';'; echo; done
HIPPO_SSIM_CPL=4 timeout 300 python tools/seg_only.py 2>&1 | grep -E "overlapped|frame pairs" | tr '\n' ';'; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gray_minmax" -c 30 --csv --log-file gpurun_out/launches_gray.csv python tools/seg_only.py > /dev/null 2>&1
grep gray_minmax gpurun_out/launches_gray.csv | awk -F'","' '{print $NF}' | sort | uniq -c | sort -rn | head -3
timeout 600 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py tests/test_gpu_fullsize.py -m gpu -q -x --tb=short -p no:cacheprovider -k "not consol and not search and not bank" 2>&1 | tail -3
