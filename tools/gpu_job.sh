set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "desired_goal_on_change or step_host" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
