# asymmetric A / B rings of the tcgen05 kernel: parity, then the batch sweep per ring split; consolidation band sweep
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_consolidation.py -m gpu -x -q --tb=short -p no:cacheprovider 2>&1 | tail -4
for A in 6 5 4 3 2; do
HIPPO_TC_ASTAGES=$A SWEEP=129,256,512,1024,2048,4096 python tools/batch_sweep.py 2>/dev/null | python -c "
import json,sys
print('A ring $A:', '  '.join(f\"{r['queries']}q {r['ms']:.3f}ms\" for r in json.load(sys.stdin)[1:]))"
done
python tools/cons_band_sweep.py 2>&1 | tail -24
