# shared score pool for a single query block (129 .. 256 queries): on (HIPPO_TC_POOL_BLOCKS=1) against off (default 2)
for PB in 2 1; do
HIPPO_TC_POOL_BLOCKS=$PB SWEEP=129,192,256 timeout 200 python tools/batch_sweep.py 2>/dev/null | python -c "
import json,sys
print('pool from $PB blocks:', '  '.join(f\"{r['queries']}q {r['ms']:.3f}ms\" for r in json.load(sys.stdin)[1:]))"
done
HIPPO_TC_POOL_BLOCKS=1 timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q --tb=short -p no:cacheprovider -k "ragged or few_queries or lattice or config1" 2>&1 | tail -3
