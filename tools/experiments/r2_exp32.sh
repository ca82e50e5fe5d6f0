# what bounds the 256- and 512-query steps: ncu --set full of one sim_tc launch each (ring split 6+6 and 4+8)
for Q in 256 512; do for A in 6 4; do
HIPPO_TC_ASTAGES=$A SWEEP=$Q timeout 600 ncu --set full --clock-control none -k regex:sim_tc_kernel -s 4 -c 1 -o gpurun_out/r2_simtc_q${Q}_a$A -f python tools/batch_sweep.py > gpurun_out/r2_simtc_q${Q}_a$A.log 2>&1; echo "q$Q a$A rc $?"
done; done
