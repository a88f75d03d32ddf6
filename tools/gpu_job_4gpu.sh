# 2- and 4-GPU calls of the round (gpurun --gpus 4): end-to-end per host route (GPR_HOST_IO=dma = compact copy engine, zerocopy).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531"
GPR_HOST_IO=dma timeout 300 $TR bench.py --gpus 4 --steps 20 --warmup 5 --quick --repeats 3 > gpurun_out/bench4_planning4_compact.log 2>&1
GPR_HOST_IO=zerocopy timeout 300 $TR bench.py --gpus 4 --steps 20 --warmup 5 --quick --repeats 3 > gpurun_out/bench4_planning4_zc.log 2>&1
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532"
GPR_HOST_IO=dma timeout 300 $TR2 bench.py --gpus 2 --steps 20 --warmup 5 --quick --repeats 3 > gpurun_out/bench2_planning4_compact.log 2>&1
GPR_HOST_IO=zerocopy timeout 300 $TR2 bench.py --gpus 2 --steps 20 --warmup 5 --quick --repeats 3 > gpurun_out/bench2_planning4_zc.log 2>&1
true
