set -x
mkdir -p gpurun_out
C=gymnasium-planar-robotics_b200/csrc
for v in b200 xar4 xar5 b200; do
GPR_B200_LIB=$PWD/$C/libgpr_$v.so timeout 300 python bench.py --steps 20 --warmup 5 --quick --no-cpu --no-extra >> gpurun_out/bench_ar2_$v.log 2>&1
done
true
