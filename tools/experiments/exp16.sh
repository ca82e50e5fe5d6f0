# 2-GPU validation of the sharded path over real NCCL + DRAM traffic with/without the slow path
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc $?"; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc $?"; cut -c1-300 gpurun_out/bench_ref_n2.json
for dbg in 0 8; do
HIPPO_TC_DEBUG=$dbg timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:sim_tc_kernel -s 3 -c 2 --csv --log-file gpurun_out/dram_dbg$dbg.csv python bench.py --steps 2 --no-extra > /dev/null 2>&1; echo "ncu dbg$dbg rc $?"; grep -E "dram__bytes|duration|hit_rate" gpurun_out/dram_dbg$dbg.csv | cut -d, -f 5,13- | head -8
done
