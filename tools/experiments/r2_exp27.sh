SWEEP=64,65,96,128,192,256,384,512,768,1024,2048,4096 python tools/batch_sweep.py > gpurun_out/r2_batch_sweep2.json 2> gpurun_out/r2_batch_sweep2.err
python - <<PY
import json
for r in json.load(open("gpurun_out/r2_batch_sweep2.json")):
    print(r["queries"], r["path"], round(r["ms"],3), "ms", round(r["tflops"],1), "TF", round(r["bank_GBps"]), "GB/s")
PY
