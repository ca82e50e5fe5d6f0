# consolidation: grids of the chain's kernels (they run on the ~16 SMs the triangle kernel leaves free)
set -u
for cfg in "0 0" "64 0" "32 0" "0 128" "0 64" "64 128" "64 64" "32 64"; do set -- $cfg
  echo "== compact grid $1 recheck(R) grid $2"; HIPPO_CONS_COMPACT_GRID=$1 HIPPO_CONS_RECHECK_GRID=$2 BANDS=8192 timeout 300 python tools/cons_band_sweep.py 2>&1 | grep -E "gamma 0.9" 
done
