# ncu --set full captures of the final SSIM and boundary kernels (one launch each) + launch list of one segmentation pass
set -u
mkdir -p gpurun_out
for kn in ssim_pair_kernel segment_kernel gray_minmax_vec_kernel; do
  timeout 150 ncu --set full --clock-control none --import-source on -k "regex:$kn" -s 3 -c 1 -f -o gpurun_out/${kn}_v7 python tools/seg_only.py > gpurun_out/ncu_$kn.log 2>&1; echo "$kn rc $?"
done
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:ssim|gray_minmax|audio_energy|segment_kernel|minmax_init" -c 40 --csv --log-file gpurun_out/launches_seg.csv python tools/seg_only.py > gpurun_out/ncu_seg_list.log 2>&1; echo "list rc $?"
