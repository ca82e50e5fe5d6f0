# the driver's command line at N GPUs, wall-clocked; sharded parity test first
N=${1:-4}
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -3
s=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "rc $? wall $(( $(date +%s) - s )) s"
tail -c 600 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]))
x=d["extra"]; print(x["sharded_single_query"]["latency_ms"], x["sharded_single_query"]["roofline"]["frac"])
c=x["config5_weak"]; print(c["queries_per_s"], c["ms_per_step"], c["single_query"]["latency_ms"])
PY
