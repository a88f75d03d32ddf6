"""Generate tests/golden/reference_reset_samples.npz: starts and goals drawn by the reference's UNMODIFIED
``BenchmarkPlanningEnv._reset_callback`` (planning:355-418; PCG64 through ``np_random``) for BASELINE configs[1]
(4 movers, 3x3 tiles) — the sample the on-device rejection sampler is compared with in distribution
(tests/test_gpu_parity.py::test_sampled_starts_follow_the_reference_distribution).  Run in the build container:

    python tests/golden/make_reset_samples.py
"""

from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
N_MOVERS, PER_PROC, PROCS = 4, 512, 8


def work(seed):
    import ref_harness

    env = ref_harness.make_planning_env(layout_tiles=np.ones((3, 3)), num_movers=N_MOVERS, std_noise=0.0)
    env.np_random = np.random.default_rng(seed)
    starts, goals = np.zeros((PER_PROC, N_MOVERS, 2)), np.zeros((PER_PROC, N_MOVERS, 2))
    for i in range(PER_PROC):
        env._reset_callback()
        for m, name in enumerate(env.mover_names):
            starts[i, m] = env.get_mover_qpos(name)[:2]
        goals[i] = env.goals
    return starts, goals


if __name__ == '__main__':
    with mp.Pool(PROCS) as pool:
        res = pool.map(work, [20261018 + k for k in range(PROCS)])
    starts = np.concatenate([r[0] for r in res])
    goals = np.concatenate([r[1] for r in res])
    np.savez_compressed(os.path.join(HERE, 'reference_reset_samples.npz'), start=starts, goal=goals)
    print('wrote', starts.shape, goals.shape)
