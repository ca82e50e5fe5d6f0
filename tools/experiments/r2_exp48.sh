# r2_v7 validation pass + ncu --set full of the 7-column SSIM kernel and its gray kernel (tools/seg_only.py as driver)
set -u
mkdir -p gpurun_out
bash tools/run_all_gpu.sh
for kn in ssim_pair7_kernel gray_minmax_vec7_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$kn" -s 3 -c 1 -f -o gpurun_out/r2_v7_$kn python tools/seg_only.py > gpurun_out/ncu_$kn.log 2>&1; echo "$kn rc $?"
  ncu -i gpurun_out/r2_v7_$kn.ncu-rep --page details > gpurun_out/r2_v7_${kn}_ncu_details.txt 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:ssim_pair|gray_minmax|audio_energy|segment_kernel|ssim_finalize|minmax_init|pattern_" -c 200 --csv --log-file gpurun_out/r2_v7_launches_seg.csv python tools/seg_only.py > gpurun_out/ncu_seg.log 2>&1; echo "launch list rc $?"
