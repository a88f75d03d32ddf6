set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 6 --warmup 3 --quick --no-cpu --no-extra --repeats 1"
timeout 300 $CMD > gpurun_out/plain4.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:planning_ -s 8 -c 2 -f -o gpurun_out/prof_planning4 $CMD > gpurun_out/ncu_full4.log 2>&1
true
