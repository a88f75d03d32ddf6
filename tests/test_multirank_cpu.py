"""world_size-2 `gloo` test of the multi-rank host logic: env indices shard across ranks with NO traffic
on the step path, results are independent of the split because the RNG is keyed by the GLOBAL env index, and the only
collective — the episode-statistics all-reduce — sums correctly.  The ranks run the CPU oracle on their shard (the CUDA
path is covered by the same test on one GPU, tests/test_gpu_parity.py::test_sharding_does_not_change_results)."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

TOTAL, STEPS, WORLD = 37, 12, 2  # odd total: ragged shards


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _kw():
    return dict(layout_tiles=np.ones((3, 3)), num_movers=3, std_noise=1e-5, seed=5, max_episode_steps=6)


def _run(env, actions):
    env.reset(seed=5)
    obs, rew, eps = [], [], 0
    ret, length = 0.0, 0.0
    for a in actions:
        env.step(a)
        obs.append(env.observation.copy())
        rew.append(env.reward.copy())
        done = (env.terminated | env.truncated).astype(bool)
        eps += int(done.sum())
    return np.stack(obs), np.stack(rew), eps


def _worker(rank, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    import gpr_oracle as oracle
    import gymnasium_planar_robotics_b200 as gpr

    base, count = gpr.shard_range(TOTAL, rank, WORLD)
    cfg, _ = gpr.planning_config(num_envs=count, env_index_base=base, **_kw())
    rng = np.random.default_rng(0)  # every rank draws the same global action tensor and slices its shard
    actions = [rng.uniform(-10, 10, (TOTAL, 6)).astype(np.float32) for _ in range(STEPS)]
    obs, rew, eps = _run(oracle.OracleEnv(cfg), [a[base:base + count] for a in actions])
    # the only collective: 6 episode counters
    counters = torch.tensor([eps, float(rew.sum()), 0.0, 0.0, 0.0, 0.0], dtype=torch.float64)
    gpr.all_reduce_stats(counters)
    # gather the shards for the comparison (test plumbing, not part of the product path)
    parts = [None] * WORLD
    dist.all_gather_object(parts, (base, count, obs, rew))
    if rank == 0:
        q.put((counters.tolist(), parts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank_and_stats_reduce():
    import gpr_oracle as oracle
    import gymnasium_planar_robotics_b200 as gpr

    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    counters, parts = q.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-rank run over all envs
    cfg, _ = gpr.planning_config(num_envs=TOTAL, **_kw())
    rng = np.random.default_rng(0)
    actions = [rng.uniform(-10, 10, (TOTAL, 6)).astype(np.float32) for _ in range(STEPS)]
    obs, rew, eps = _run(oracle.OracleEnv(cfg), actions)
    parts.sort(key=lambda x: x[0])
    assert [p[0] for p in parts] == [0, 19] and [p[1] for p in parts] == [19, 18]  # ragged split, contiguous, complete
    assert np.array_equal(obs, np.concatenate([p[2] for p in parts], axis=1))
    assert np.array_equal(rew, np.concatenate([p[3] for p in parts], axis=1))
    assert counters[0] == eps and eps > 0 and np.isclose(counters[1], rew.sum())
    assert gpr.stats_dict(counters)['episodes'] == eps
