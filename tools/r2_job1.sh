# round-2 GPU call 1: parity tests (incl. the reference-trajectory replays), baseline benches, ncu of the box / pushing kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
python -c "import mujoco" > gpurun_out/mujoco_probe.log 2>&1; python -c "import gymnasium" >> gpurun_out/mujoco_probe.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_planning4.log 2>&1
timeout 600 python bench.py --workload pushing --steps 50 --warmup 5 --no-cpu --quick > gpurun_out/bench_pushing.log 2>&1
timeout 600 python bench.py --workload planning8box --steps 20 --warmup 3 --no-cpu --quick > gpurun_out/bench_planning8box.log 2>&1
CMD="python bench.py --workload planning8box --steps 4 --warmup 3 --quick --no-cpu"
timeout 300 $CMD > gpurun_out/plain8.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:planning_ -s 8 -c 2 -f -o gpurun_out/prof_planning8box $CMD > gpurun_out/ncu_full8.log 2>&1
CMD="python bench.py --workload pushing --steps 4 --warmup 3 --quick --no-cpu"
timeout 300 $CMD > gpurun_out/plainp.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pushing_step -s 3 -c 1 -f -o gpurun_out/prof_pushing $CMD > gpurun_out/ncu_fullp.log 2>&1
true
