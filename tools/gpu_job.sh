# Standard GPU verification run of a round (under gpurun, one B200):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_job.sh'
# parity tests, smoke, the default bench (headline + extra_workloads) + the reference arm, then the ncu launch list and one
# full capture per workload (each ncu command only after the same command exited 0 without ncu).  Outputs land in
# gpurun_out/; tools/ncu_summary.py and tools/attribute_lines.py turn the captures into the summaries under profiles/.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.log 2>&1
timeout 600 python bench.py --workload pushing --steps 50 --warmup 5 --quick > gpurun_out/bench_pushing.log 2>&1
timeout 600 python bench.py --workload planning8box --steps 20 --warmup 3 --quick > gpurun_out/bench_planning8box.log 2>&1
python tools/pcie_bw.py > gpurun_out/pcie_bw.log 2>&1
CMD="python bench.py --steps 10 --warmup 3 --quick --no-cpu --no-extra --repeats 1"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:planning_ -s 8 -c 2 -f -o gpurun_out/prof_planning4 $CMD > gpurun_out/ncu_full4.log 2>&1
CMD="python bench.py --workload planning8box --steps 4 --warmup 3 --quick --no-cpu --repeats 1"
timeout 300 $CMD > gpurun_out/plain8.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:planning_ -s 8 -c 2 -f -o gpurun_out/prof_planning8box $CMD > gpurun_out/ncu_full8.log 2>&1
CMD="python bench.py --workload pushing --steps 30 --warmup 30 --quick --no-cpu --repeats 1"
timeout 300 $CMD > gpurun_out/plainp.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pushing_ -s 80 -c 2 -f -o gpurun_out/prof_pushing $CMD > gpurun_out/ncu_fullp.log 2>&1
true
