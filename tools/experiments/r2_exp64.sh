# mid-size batches against the flow-control window (cohorts of 2 / 4 pairs at 512 / 1,024 queries)
for w in 8 0 2 16 32; do
  echo "== HIPPO_TC_WINDOW=$w"; HIPPO_TC_WINDOW=$w SWEEP=256,512,1024,2048 timeout 300 python tools/batch_sweep.py 2>&1 >/dev/null | grep -E "'queries': (256|512|1024|2048)" | sed "s/'path'.*'ms': /ms /; s/, 'queries_per_s.*tflops': / TF /; s/}//"
done
