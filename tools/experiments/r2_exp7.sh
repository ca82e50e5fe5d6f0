# round 2, call 7: timeline of the overlapped segmentation pipeline (timing events per chunk / chain launch)
set -u
mkdir -p gpurun_out
CHUNKS=600 TIMELINE=600 timeout 300 python tools/seg_only.py 2>&1 | grep -E "seg_only|pattern" | head -80
KREGEX='regex:ssim|gray_minmax|audio_energy|segment_kernel|minmax_init|pattern_init'
CHUNKS=600 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 120 --csv --log-file gpurun_out/r2_launches_seg_pipeline.csv python tools/seg_only.py > gpurun_out/r2_ncu_seg_pipeline.log 2>&1; echo "ncu rc $?"
