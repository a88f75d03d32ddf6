set -x
mkdir -p gpurun_out
C=gymnasium-planar-robotics_b200/csrc
export GPR_B200_LIB=$PWD/$C/libgpr_b200.so
CMD="python bench.py --workload pushing --steps 4 --warmup 3 --quick --no-cpu"
timeout 300 $CMD > gpurun_out/plainp.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pushing_contact -s 3 -c 1 -f -o gpurun_out/prof_push_contact $CMD > gpurun_out/ncu_fullp.log 2>&1
true
