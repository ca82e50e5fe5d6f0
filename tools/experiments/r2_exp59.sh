python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
out = bench.run_next_rows(dev, bench.load_peaks())
print(json.dumps(out, indent=1))
PY
