set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pushing.py -m gpu -q -x > gpurun_out/pytest_push.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_push.log
timeout 600 python bench.py --workload pushing --steps 50 --warmup 5 > gpurun_out/bench_push_r1f.log 2>&1
timeout 600 python bench.py --workload planning8box --steps 20 --warmup 3 > gpurun_out/bench_p8_r1f.log 2>&1
ls -la gpurun_out
