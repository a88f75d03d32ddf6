set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 --quick --repeats 3 > gpurun_out/bench8b_planning4_compact.log 2>&1
GPR_HOST_COMPACT=0 timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 --quick --repeats 3 > gpurun_out/bench8b_planning4_dense.log 2>&1
true
