set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 10 --warmup 3 --quick --no-cpu"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1m.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:planning_ -s 8 -c 2 -o gpurun_out/prof_r1m $CMD > gpurun_out/ncu2.log 2>&1
CMD2="python bench.py --workload pushing --steps 6 --warmup 3 --quick --no-cpu"
timeout 300 $CMD2 > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pushing_step -s 4 -c 1 -o gpurun_out/prof_push_r1m $CMD2 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out
