# ncu --set full captures (one launch each) of the secondary kernels, driven by tools/extras_only.py
set -u
mkdir -p gpurun_out
for kn in ssim_pair_kernel gray_minmax_kernel greedy_scan_kernel segment_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kn -s 3 -c 1 -f -o gpurun_out/$kn python tools/extras_only.py 200000 > gpurun_out/ncu_$kn.log 2>&1; echo "$kn rc $?"
done
