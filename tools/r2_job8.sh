set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_debug_bounds.py -m gpu -q -x > gpurun_out/pytest_dbg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dbg.log
true
