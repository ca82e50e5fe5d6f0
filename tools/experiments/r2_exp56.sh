# 8-GPU driver-style bench of the final code (strong scaling of the 10M-row bank + config 5: 80M rows)
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench rc $?"
grep -E "gate" gpurun_out/bench_n8.err | wc -l
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), json.dumps(d.get("extra",{}).get("config5_weak",{}))[:400])
print(json.dumps(d.get("extra",{}).get("sharded_single_query",{}))[:300])
PY
