CMD="python bench.py --games 65536 --steps 1 --warmup 2 --no-cpu-baseline --no-extra"
$CMD > gpurun_out/plain_g8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 2000 -c 2 -o gpurun_out/step_g8 $CMD > gpurun_out/ncu_g8.log 2>&1
tail -n 2 gpurun_out/ncu_g8.log
