# gray kernels: next item's loads under the conversion, grid = resident CTAs
set -u
for cpl in 7 4; do
  echo "== cpl $cpl"; HIPPO_SSIM_CPL=$cpl BATCH=32 timeout 300 python tools/seg_only.py 2>&1 | grep -E "overlapped|stages, 3|pipeline, 2|digest"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gray_minmax" -c 40 --csv --log-file gpurun_out/launches_gray.csv python tools/seg_only.py > /dev/null 2>&1
grep gray_minmax gpurun_out/launches_gray.csv | awk -F'","' '{print $NF}' | sort | uniq -c | sort -rn | head -4
timeout 600 python -m pytest tests/test_gpu_segmentation.py tests/test_gpu_prefilter.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
