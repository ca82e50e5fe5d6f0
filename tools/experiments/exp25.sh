# 2-GPU validation of the sharded search: parity test, both bench arms under torchrun
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc $?"; tail -4 gpurun_out/bench_n2.err; cut -c1-400 gpurun_out/bench_n2.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc $?"; cut -c1-200 gpurun_out/bench_ref_n2.json
