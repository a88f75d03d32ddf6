set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
L=gymnasium-planar-robotics_b200/csrc
for v in b200 x3 x4; do
  GPR_B200_LIB=$PWD/$L/libgpr_$v.so timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_var_$v.log 2>&1
  GPR_B200_LIB=$PWD/$L/libgpr_$v.so timeout 600 python bench.py --workload planning8box --steps 10 --warmup 3 --no-cpu --quick > gpurun_out/bench_var_p8_$v.log 2>&1
done
ls -la gpurun_out
