# same-box A/B of the score pool for 129 .. 256 queries (old rule: HIPPO_TC_POOL_BLOCKS=2), twice; then the search tests
for rep in 1 2; do for PB in 2 1; do
HIPPO_TC_POOL_BLOCKS=$PB SWEEP=129,160,192,224,256 timeout 200 python tools/batch_sweep.py 2>/dev/null | python -c "
import json,sys
print('rep $rep pool blocks >= $PB:', '  '.join(f\"{r['queries']}q {r['ms']:.3f}ms\" for r in json.load(sys.stdin)[1:]))"
done; done
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q --tb=short -p no:cacheprovider 2>&1 | tail -3
