set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_segmentation.py -m gpu -q --tb=short --timeout 500 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -30
for ex in p2p nccl; do
HIPPO_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_$ex.json 2> gpurun_out/bench_n2_$ex.err; echo "n2 $ex rc $?"; tail -3 gpurun_out/bench_n2_$ex.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2_$ex.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['exchange'])"
done
