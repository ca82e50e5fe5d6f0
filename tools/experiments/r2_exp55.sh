# 2-GPU validation of the final code: sharded parity test + the driver-style bench command
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc $?"
grep -E "gate|config5|extra" gpurun_out/bench_n2.err | tail -12
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), json.dumps(d.get("extra",{}).get("config5_weak",{}))[:500])
PY
