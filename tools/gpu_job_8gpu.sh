# 8-GPU call (gpurun --gpus 8; charged 8x): headline scaling incl. end-to-end on the automatic host route (compact copy engine
# at 8 ranks), pushing at 65,536 envs per GPU and the 8 M-env pushing sweep point of BASELINE configs[4] (1,048,576 envs per GPU).
# Earlier runs of the round also took: tools/pcie_bw.py under the same torchrun line (host-ingest ceiling with 8 ranks copying
# at once), GPR_HOST_IO=zerocopy and GPR_HOST_COMPACT=0 variants of the first line, and --num-envs 1048576 for planning4
# (profiles/r2_bench_multi_gpu.txt, profiles/r2_pcie_bandwidth.txt).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 --quick --repeats 3 > gpurun_out/bench8_planning4_auto.log 2>&1
timeout 400 $TR bench.py --gpus 8 --workload pushing --steps 20 --warmup 5 --quick --repeats 3 > gpurun_out/bench8_pushing.log 2>&1
timeout 400 $TR bench.py --gpus 8 --workload pushing --num-envs 1048576 --steps 10 --warmup 3 --quick --repeats 3 > gpurun_out/bench8_pushing_8M.log 2>&1
true
