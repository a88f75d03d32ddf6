set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pushing.py tests/test_reference_trajectories.py -m gpu -q -x -k "pushing" > gpurun_out/pytest_push.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_push.log
timeout 300 python bench.py --workload pushing --steps 50 --warmup 5 --no-cpu --quick > gpurun_out/bench_push_b200.log 2>&1
CMD="python bench.py --workload pushing --steps 4 --warmup 3 --quick --no-cpu --repeats 1"
timeout 300 $CMD > gpurun_out/plainp.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pushing_ -s 6 -c 2 -f -o gpurun_out/prof_pushing $CMD > gpurun_out/ncu_fullp.log 2>&1
true
