set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
GPR_HOST_IO=dma timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu --quick > gpurun_out/bench_dma_n1.log 2>&1
