# 2-CTA small-batch kernel (65 .. 128 queries): parity first (under timeouts), then the sweep against the 1-CTA tile
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q --tb=short -p no:cacheprovider -k "few_queries or ragged" 2>&1 | tail -5
SWEEP=64,65,96,128 timeout 200 python tools/batch_sweep.py 2>/dev/null | python -c "
import json,sys
for r in json.load(sys.stdin): print('pair', r['queries'], r['path'], round(r['ms'],3), 'ms', round(r['bank_GBps']), 'GB/s')"
HIPPO_SMALL_PAIR=0 SWEEP=65,96,128 timeout 200 python tools/batch_sweep.py 2>/dev/null | python -c "
import json,sys
for r in json.load(sys.stdin): print('1cta', r['queries'], r['path'], round(r['ms'],3), 'ms', round(r['bank_GBps']), 'GB/s')"
bash tools/experiments/r2_exp34.sh
