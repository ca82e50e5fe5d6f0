# ncu --set full captures (one launch each) of the secondary kernels, driven by tools/extras_only.py
set -u
mkdir -p gpurun_out
for kn in ssim_pair7_kernel gray_minmax_vec7_kernel greedy_scan_kernel segment_kernel cons_advance_kernel "sim_tc_kernel<1>"; do
  out=$(echo $kn | tr -d '<>')
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$kn" -s 5 -c 1 -f -o gpurun_out/$out python tools/extras_only.py 200000 > gpurun_out/ncu_$out.log 2>&1; echo "$kn rc $?"
done
KREGEX="regex:sim_tc_kernel|greedy_scan_kernel|recheck_kernel|cons_advance|cons_finish|ssim_pair|gray_minmax|audio_energy|segment_kernel|ssim_finalize|minmax_init|bank_build"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file gpurun_out/launches_extras.csv python tools/extras_only.py 200000 > gpurun_out/ncu_extras.log 2>&1; echo "launch list rc $?"
