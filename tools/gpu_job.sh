set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_r1l.log 2>&1
timeout 600 python bench.py --workload planning8box --steps 20 --warmup 3 --no-cpu --quick > gpurun_out/bench_p8_r1l.log 2>&1
GPR_B200_LIB=$PWD/gymnasium-planar-robotics_b200/csrc/libgpr_x4.so timeout 600 python bench.py --workload planning8box --steps 20 --warmup 3 --no-cpu --quick > gpurun_out/bench_p8_r1l_x4.log 2>&1
