# round 2, call 1 (no code change since r1_v7): what holds the tensor pipe (HIPPO_TC_DEBUG=64 counters at the full
# config), --set full captures of the three bandwidth kernels at 10M rows, segmentation stage split as the baseline
set -u
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - T0 ))s"; }
timeout 300 python tools/prof_counters.py > gpurun_out/r2_tc_counters.log 2>&1; echo "counters rc $?"; tail -3 gpurun_out/r2_tc_counters.log; lap counters
timeout 300 python tools/r2_ncu_bw.py > gpurun_out/r2_bw_plain.log 2>&1; echo "bw plain rc $?"; tail -1 gpurun_out/r2_bw_plain.log; lap bw_plain
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:bank_build_kernel|topk_single_kernel|topk_few_kernel" -f -o gpurun_out/r2_bw_kernels python tools/r2_ncu_bw.py > gpurun_out/r2_bw_ncu.log 2>&1; echo "bw ncu rc $?"; tail -2 gpurun_out/r2_bw_ncu.log; lap bw_ncu
HIPPO_SEG_DEBUG=1 timeout 200 python tools/seg_only.py > gpurun_out/r2_seg_baseline.log 2>&1; echo "seg rc $?"; tail -3 gpurun_out/r2_seg_baseline.log; lap seg
