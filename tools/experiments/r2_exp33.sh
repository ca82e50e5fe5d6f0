for Q in 256 512 4096; do echo "== $Q queries"; NQ=$Q timeout 120 python tools/prof_counters.py 2>&1 | tail -2; done
echo "== 256 queries, TMEM loads but no epilogue math (EXTRA_DEBUG=2)"; EXTRA_DEBUG=2 NQ=256 timeout 120 python tools/prof_counters.py 2>&1 | tail -1
echo "== 256 queries, never the slow path (EXTRA_DEBUG=8)"; EXTRA_DEBUG=8 NQ=256 timeout 120 python tools/prof_counters.py 2>&1 | tail -1
echo "== 256 queries, no norm staging (EXTRA_DEBUG=4)"; EXTRA_DEBUG=4 NQ=256 timeout 120 python tools/prof_counters.py 2>&1 | tail -1
