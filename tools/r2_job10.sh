set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host or goal_on_change or obstacles_through" > gpurun_out/pytest_host.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_host.log
GPR_HOST_IO=dma timeout 300 python bench.py --steps 20 --warmup 5 --quick --no-cpu --no-extra --repeats 3 > gpurun_out/bench_p4_dma_compact.log 2>&1
GPR_HOST_IO=dma GPR_HOST_COMPACT=0 timeout 300 python bench.py --steps 20 --warmup 5 --quick --no-cpu --no-extra --repeats 3 > gpurun_out/bench_p4_dma_dense.log 2>&1
true
