# 65 .. 128 queries through the small-batch kernel (N = 128): parity tests, then the sweep with and without it
timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -x -q --tb=short -p no:cacheprovider -k "few_queries or ragged" 2>&1 | tail -5
SWEEP=64,65,96,128,129 python tools/batch_sweep.py 2>/dev/null | python -c "
import json,sys
for r in json.load(sys.stdin): print('N128', r['queries'], r['path'], round(r['ms'],3), 'ms', round(r['bank_GBps']), 'GB/s')"
HIPPO_SMALL_BATCH=64 SWEEP=65,96,128 python tools/batch_sweep.py 2>/dev/null | python -c "
import json,sys
for r in json.load(sys.stdin): print('old ', r['queries'], r['path'], round(r['ms'],3), 'ms', round(r['bank_GBps']), 'GB/s')"
