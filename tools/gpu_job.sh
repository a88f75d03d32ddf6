set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_final_planning4.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_reference.log 2>&1
timeout 600 python bench.py --workload pushing --steps 50 --warmup 5 > gpurun_out/bench_final_pushing.log 2>&1
timeout 600 python bench.py --workload planning8box --steps 20 --warmup 3 > gpurun_out/bench_final_p8.log 2>&1
CMD="python bench.py --steps 10 --warmup 3 --quick --no-cpu"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:planning_ -s 8 -c 2 -o gpurun_out/prof_final $CMD > gpurun_out/ncu2.log 2>&1
