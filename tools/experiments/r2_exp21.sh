# round 2, call 21: score-pool update once per chunk instead of once per hit: clustered / Gaussian banks, search tests
set -u
timeout 300 python -m pytest tests/test_gpu_search.py tests/test_gpu_sharded.py -m gpu -q --tb=short --timeout 200 -p no:cacheprovider -x 2>&1 | tail -4
timeout 300 python tools/other_banks.py 2>&1 | grep -E "extra|ms_per_step|mma_wait|slow_chunks\"" 
timeout 200 python bench.py --steps 10 --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lattice', d['ms_per_step'], d['roofline']['frac'])"
