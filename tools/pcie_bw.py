"""Host<->device bandwidth with page-locked memory — the physical bound of the end-to-end leg.

    python tools/pcie_bw.py                                  one GPU: copy engine, 2 / 16 / 128 MB
    torchrun --nproc-per-node 8 tools/pcie_bw.py             8 ranks AT ONCE on one host: the host-ingest ceiling that bounds
                                                             the 8-GPU end-to-end number (every rank copies concurrently;
                                                             reported per rank and in total)
Also measured: SM-issued zero-copy stores into pinned host memory (what gpr_step_host's default route uses).
"""
import json
import os

import torch

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
if world > 1:
    import torch.distributed as dist

    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
res = {}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for mb in (2, 8, 16, 128):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, (src, dst) in (('d2h', (d, h)), ('h2d', (h, d))):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        e1.record()
        barrier()
        res[f'{name}_{mb}MB_GBps'] = round(20 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
if world > 1:
    t = torch.tensor([res[k] for k in sorted(res)], dtype=torch.float64, device=dev)
    tot = t.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    mn = t.clone()
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({'ranks': world, 'total_GBps': dict(zip(sorted(res), [round(x, 1) for x in tot.tolist()])),
                          'slowest_rank_GBps': dict(zip(sorted(res), [round(x, 1) for x in mn.tolist()]))}))
    dist.destroy_process_group()
else:
    print(json.dumps(res))
