set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo bench rc=$?
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --blocks 0 --e2e-steps 1"
$CMD > gpurun_out/plain_r2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_l_r2.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:pxm_ring_fft3_kernel -s 12 -c 2 -o gpurun_out/prof_fft3_r2 -f $CMD > gpurun_out/ncu_fft3_r2.log 2>&1; echo full rc=$?
python scripts/refresh_traffic.py > gpurun_out/refresh_traffic.log 2>&1; echo traffic rc=$?; tail -5 gpurun_out/refresh_traffic.log
ncu --set full --clock-control none --import-source on -k regex:pxm_legendre_kernel -s 61 -c 3 -o gpurun_out/prof_leg_r2 -f $CMD > gpurun_out/ncu_leg_r2.log 2>&1; echo legfull rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_myula_update_pair -s 6 -c 1 -o gpurun_out/prof_upd_r2 -f $CMD > gpurun_out/ncu_upd_r2.log 2>&1; echo updfull rc=$?
