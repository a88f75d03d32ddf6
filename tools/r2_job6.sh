set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/bench_p4.log 2>&1
timeout 300 python bench.py --workload planning8box --steps 20 --warmup 3 --no-cpu --quick > gpurun_out/bench_p8_b200.log 2>&1
true
