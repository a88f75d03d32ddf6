set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pushing.py tests/test_reference_trajectories.py tests/test_gpu_parity.py -m gpu -q -x -k "pushing or lockstep or ordered or distribution" > gpurun_out/pytest_push.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_push.log
C=gymnasium-planar-robotics_b200/csrc
for v in b200 xpc4 xpc8; do
GPR_B200_LIB=$PWD/$C/libgpr_$v.so timeout 300 python bench.py --workload pushing --steps 50 --warmup 5 --no-cpu --quick > gpurun_out/bench_push_$v.log 2>&1
GPR_B200_LIB=$PWD/$C/libgpr_$v.so timeout 300 python bench.py --workload pushing --steps 20 --warmup 5 --no-cpu --quick --num-envs 1048576 > gpurun_out/bench_push_${v}_1M.log 2>&1
done
true
