set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_n2.log
