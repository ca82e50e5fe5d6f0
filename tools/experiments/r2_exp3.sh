# round 2, call 3 (2 GPUs): sharded-search test, bench at N=2 with the config-5 block and the per-rank parity gate
set -u
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - T0 ))s"; }
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -4; lap sharded_test
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc $?"; tail -12 gpurun_out/r2_bench_n2.err; lap bench2
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]))
print(json.dumps(d.get("extra",{}), indent=1)[:3000])
PY
