# Multi-GPU validation + measurement (run under `gpurun --gpus N` from the repo root): bash tools/run_multi_gpu.sh N [rows_per_gpu_for_config5]
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -4
run() { # name, extra env, extra args
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/bench_${name}.json 2> gpurun_out/bench_${name}.err
  echo "$name rc $?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${name}.json").read().strip().splitlines()[-1])
    print("  ", d["n_gpus"], "GPUs", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), d["config"].get("parallelism","")[:90], d.get("extra",{}).get("exchange"))
except Exception as e:
    print("   parse failed", e); print(open("gpurun_out/bench_${name}.err").read()[-1500:])
PY
}
run strong${N}_p2p HIPPO_EXCHANGE=p2p
run strong${N}_nccl HIPPO_EXCHANGE=nccl --no-extra
run weak${N}_p2p HIPPO_EXCHANGE=p2p --no-extra --bank-rows $((10000000 * N))
