# round 2, call 5 (2 GPUs): fused sharded calls (batched: merge inside the exchange kernel; single query: one launch),
# sharded test, bench at N=2
set -u
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - T0 ))s"; }
timeout 300 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_search.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -8; lap tests
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; echo "bench n2 rc $?"; tail -6 gpurun_out/r2_bench_n2b.err; lap bench2
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n2b.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]))
e=d["extra"]
print(e["per_rank_local_pass_ms"]); print(e["sharded_single_query"]["latency_ms"], e["sharded_single_query"]["roofline"]["frac"])
c=e["config5_weak"]; print(c["ms_per_step"], c["queries_per_s"], c["single_query"]["latency_ms"])
PY
