CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra"
true
$CMD > gpurun_out/plain_r1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 400 --csv --log-file gpurun_out/launches_r1e.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain_r1b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3000 -c 3 -o gpurun_out/step_r1e $CMD > gpurun_out/ncu2.log 2>&1
$CMD > gpurun_out/plain_r1c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 3000 -c 2 -o gpurun_out/pair_r1e $CMD > gpurun_out/ncu3.log 2>&1
tail -n 3 gpurun_out/ncu1.log; tail -n 3 gpurun_out/ncu2.log; tail -n 3 gpurun_out/ncu3.log
