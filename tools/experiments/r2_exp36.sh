# global score pool in the small-batch kernel; pool from 129 queries in the 256-query tiles: parity, then the sweep
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_fullsize.py tests/test_gpu_recall.py -m gpu -x -q --tb=short -p no:cacheprovider -k "not consolidation and not stream_hour" 2>&1 | tail -5
SWEEP=2,3,8,16,32,33,64,65,96,128,129,192,256,512 timeout 300 python tools/batch_sweep.py 2>/dev/null > gpurun_out/r2_batch_sweep3.json
python - <<PY
import json
for r in json.load(open("gpurun_out/r2_batch_sweep3.json")):
    print(r["queries"], r["path"], round(r["ms"],3), "ms", round(r["bank_GBps"]), "GB/s", round(r["tflops"],1), "TF")
PY
