set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m "gpu and not slow" -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload planning4 --steps 20 --warmup 5 --no-cpu --quick --no-extra > gpurun_out/bench_p4.log 2>&1
true
