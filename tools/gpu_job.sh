set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --workload pushing --steps 50 --warmup 5 > gpurun_out/bench_push_r1p.log 2>&1
timeout 600 python bench.py --workload planning8box --steps 20 --warmup 3 > gpurun_out/bench_p8_r1p.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r1p.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r1p.log 2>&1
