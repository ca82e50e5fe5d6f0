# ncu --set full capture of the batched-search kernel at the full 10M-row config (one launch, after warm-up)
set -u
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sim_tc_kernel -s 3 -c 1 -f -o gpurun_out/sim_tc_topk_v7 python bench.py --steps 1 --no-extra > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"; tail -3 gpurun_out/ncu_full.log
