# full validation of the follow-mode segmentation + ncu evidence for it
bash tools/run_all_gpu.sh
KREGEX='regex:ssim_pair|gray_minmax|audio_energy|segment_kernel|minmax_init|pattern_'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file gpurun_out/r2_launches_seg_follow.csv env CHUNKS=0 python tools/seg_only.py > gpurun_out/r2_ncu_seg_follow.log 2>&1; echo "ncu seg rc $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:segment_kernel -c 3 -o gpurun_out/r2_segment_follow env CHUNKS=0 python tools/seg_only.py > gpurun_out/r2_ncu_segment_follow.log 2>&1; echo "ncu full rc $?"
