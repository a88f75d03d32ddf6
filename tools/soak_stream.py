"""Soak test of the streamed auto-reset (DESIGN.md §5 "Launch structure"): the same seeds and actions through an env whose
auto-reset kernel overlaps the step kernel (production) and one whose kernels are serialised (per-kernel timing on),
for many steps at full batch sizes; every output and the final state must be identical.  Run on a B200:
    python tools/soak_stream.py > gpurun_out/soak.log
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnasium_planar_robotics_b200 as gpr  # noqa: E402

DEV = 'cuda:0'
CASES = [
    # (num_envs, steps, kwargs)
    (65536, 400, dict(layout_tiles=np.ones((3, 3)), num_movers=4)),
    (65536, 200, dict(layout_tiles=np.ones((3, 3)), num_movers=4, autoreset_mode='next_step')),
    (1048576, 60, dict(layout_tiles=np.ones((3, 3)), num_movers=4)),
    (65536 + 13, 150, dict(layout_tiles=np.ones((3, 3)), num_movers=2, learn_jerk=True)),
    (262144, 40, dict(layout_tiles=np.ones((5, 5)), num_movers=8, learn_jerk=True, collision_params={'shape': 'box', 'size': np.array([0.08, 0.08])})),
    (65536, 150, dict(layout_tiles=np.ones((4, 4)), num_movers=3, obstacles=[[0.36, 0.36, 0.05], [0.7, 0.2, 0.03]], collision_params={'shape': 'circle', 'size': 0.08})),
]
ok = True
for B, steps, kw in CASES:
    t0 = time.time()
    a = gpr.BenchmarkPlanningVecEnv(B, device=DEV, seed=7, **kw)
    b = gpr.BenchmarkPlanningVecEnv(B, device=DEV, seed=7, **kw)
    b.core.kernel_times(True)  # serial launches
    a.reset(seed=7)
    b.reset(seed=7)
    lim = a.j_max if a.learn_jerk else a.a_max
    gen = torch.Generator(device=DEV).manual_seed(1)
    bad = None
    for t in range(steps):
        act = (torch.rand((B, a.core.action_dim), device=DEV, generator=gen) * 2 - 1) * lim
        oa, ra, ta, tra, ia = a.step(act)
        ob, rb, tb, trb, ib = b.step(act)
        same = all(torch.equal(oa[k], ob[k]) for k in oa) and torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(tra, trb) \
            and all(torch.equal(ia[k], ib[k]) for k in ('is_success', 'mover_collision', 'wall_collision'))
        if not same:
            bad = t
            break
        if t % 50 == 49:
            b.core.kernel_times(True)  # (drop the accumulated events)
    sa, sb = a.get_state(), b.get_state()
    state_same = all(torch.equal(sa[k], sb[k]) for k in sa)
    fails = (a.core.reset_failures(), b.core.reset_failures())
    eps = a.episode_stats()['episodes']
    good = bad is None and state_same and fails == (0, 0)
    ok = ok and good
    print(f"{'OK ' if good else 'BAD'} envs={B} steps={steps} kw={ {k: (v if not isinstance(v, np.ndarray) else v.shape) for k, v in kw.items()} } "
          f"first_mismatch={bad} state_same={state_same} reset_failures={fails} episodes={eps:.0f} {time.time() - t0:.1f}s", flush=True)
    a.close()
    b.close()
print('SOAK', 'PASSED' if ok else 'FAILED')
sys.exit(0 if ok else 1)
