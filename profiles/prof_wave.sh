CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra"
true
$CMD > gpurun_out/plain_w1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/launches_r1g.csv $CMD > gpurun_out/ncu_w1.log 2>&1
$CMD > gpurun_out/plain_w2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_wave_kernel -s 700 -c 3 -o gpurun_out/wave_r1g $CMD > gpurun_out/ncu_w2.log 2>&1
tail -n 2 gpurun_out/ncu_w1.log | cut -c1-200; tail -n 2 gpurun_out/ncu_w2.log | cut -c1-200
$CMD > gpurun_out/plain_w3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mlp_pair2_kernel -s 700 -c 2 -o gpurun_out/pair2_r1g $CMD > gpurun_out/ncu_w3.log 2>&1
tail -n 2 gpurun_out/ncu_w3.log | cut -c1-200
