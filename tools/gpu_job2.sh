set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 240 env GPR_HOST_IO=${GPR_HOST_IO:-zc} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n$N.log
