# e2e with two steps in flight (host waits for the previous step's results): 1 GPU sanity
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra 2>gpurun_out/b1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2))"
tail -2 gpurun_out/b1.err
