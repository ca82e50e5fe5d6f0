# e2e with two steps in flight at 8 GPUs (strong scaling of the 10M-row bank, no config-5 block)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-extra > gpurun_out/bench_n8_e2e.json 2> gpurun_out/bench_n8_e2e.err; echo "rc $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n8_e2e.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3))"
