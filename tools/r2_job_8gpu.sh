# 8-GPU call: headline scaling incl. end-to-end (auto = DMA route vs forced zero-copy), the host-ingest ceiling with 8 ranks
# copying at once, and the 8M-env sweep points of BASELINE configs[4] (1,048,576 envs per GPU)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR tools/pcie_bw.py > gpurun_out/pcie_bw_8ranks.log 2>&1
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 --quick > gpurun_out/bench8_planning4_auto.log 2>&1
GPR_HOST_IO=zerocopy timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 --quick > gpurun_out/bench8_planning4_zc.log 2>&1
timeout 400 $TR bench.py --gpus 8 --num-envs 1048576 --steps 10 --warmup 3 --quick --repeats 3 > gpurun_out/bench8_planning4_8M.log 2>&1
timeout 400 $TR bench.py --gpus 8 --workload pushing --num-envs 1048576 --steps 10 --warmup 3 --quick --repeats 3 > gpurun_out/bench8_pushing_8M.log 2>&1
timeout 400 $TR bench.py --gpus 8 --workload pushing --steps 20 --warmup 5 --quick --repeats 3 > gpurun_out/bench8_pushing.log 2>&1
true
