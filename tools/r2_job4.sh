set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pushing.py tests/test_reference_trajectories.py tests/test_gpu_parity.py -m gpu -q -x -k "pushing" > gpurun_out/pytest_push.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_push.log
timeout 300 python bench.py --workload pushing --steps 50 --warmup 5 --no-cpu --quick > gpurun_out/bench_push_b200.log 2>&1
timeout 300 python bench.py --workload pushing --steps 20 --warmup 5 --no-cpu --quick --num-envs 1048576 > gpurun_out/bench_push_b200_1M.log 2>&1
true
