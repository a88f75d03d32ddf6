"""Time the reference's UNMODIFIED NumPy share of the step path — build container only.

    python tools/reference_numpy_share.py        -> profiles/reference_numpy_share.json

The MuJoCo-backed reference cannot run here (no wheel offline), but everything it executes in Python/NumPy per control cycle
can, unmodified, through tests/ref_harness.py: `ensure_max_dyn_val` (planning:610-645) per mover, `qpos_is_valid`
(basic:459-788) and `check_mover_collision` (basic:355-424) per cycle, 40 cycles per env-step.  That share is an UPPER
bound on the real reference's env-steps/s/core (it excludes mj_step, name lookups, noise draws, observation assembly).
Also timed: the reference's whole unmodified `step()` on the closed-form MuJoCo stand-in (tests/mujoco_standin.py) — every
line of the reference's Python, with mj_step replaced by a few NumPy operations (labelled as such: not MuJoCo).
`bench.py` copies the committed JSON into its line (`reference_numpy_share`), since the GPU box has no reference mount.
"""

from __future__ import annotations

import json
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'tests'), ROOT]
import ref_harness  # noqa: E402

CASES = {
    'configs0_planning_n2_circle': dict(layout_tiles=np.ones((3, 3)), num_movers=2),
    'configs1_planning_n4_circle': dict(layout_tiles=np.ones((3, 3)), num_movers=4),
    'configs3_planning_n8_box_jerk': dict(layout_tiles=np.ones((5, 5)), num_movers=8, learn_jerk=True,
                                          collision_params={'shape': 'box', 'size': np.array([0.08, 0.08])}),
}


def numpy_share(kw, cycles=400):
    env = ref_harness.make_planning_env(std_noise=0.0, **kw)
    N = env.num_movers
    rng = np.random.default_rng(0)
    qpos = np.zeros((N, 7))
    qpos[:, 3] = 1.0
    lo, hi = env.min_xy_pos, env.max_xy_pos
    qpos[:, :2] = rng.uniform(lo, hi, (N, 2))
    vel, acc = rng.normal(0, 0.5, (N, 2)), rng.normal(0, 3.0, (N, 2))
    act = rng.uniform(-10, 10, (N, 2))
    t0 = time.perf_counter()
    for _ in range(cycles):
        for m in range(N):
            if env.learn_jerk:
                a_tmp, j = env.ensure_max_dyn_val(acc[m], env.a_max, act[m])
                env.ensure_max_dyn_val(vel[m], env.v_max, a_tmp)
            else:
                env.ensure_max_dyn_val(vel[m], env.v_max, act[m])
        env.qpos_is_valid(qpos, env.c_size, add_safety_offset=False)
        env.check_mover_collision(env.mover_names, env.c_size, add_safety_offset=False, mover_qpos=qpos)
    per_cycle = (time.perf_counter() - t0) / cycles
    return per_cycle


def full_step_on_standin(kw, steps=30):
    env = ref_harness.make_planning_env(**kw)  # reference-default std_noise
    env.np_random = np.random.default_rng(0)
    env.reset()
    lim = env.j_max if env.learn_jerk else env.a_max
    rng = np.random.default_rng(1)
    n, cycles_run, t = 0, 0, 0.0
    for _ in range(steps):
        a = rng.uniform(-lim, lim, 2 * env.num_movers) * 0.05  # gentle actions: episodes that run all 40 cycles
        t0 = time.perf_counter()
        _, _, term, _, _ = env.step(a)
        t += time.perf_counter() - t0
        n += 1
        if term:
            env.reset()
    return t / n


if __name__ == '__main__':
    out = {'where': 'build container (reference mounted read-only at /root/reference), one core', 'python': platform.python_version(),
           'numpy': np.__version__, 'cpu': platform.processor() or platform.machine(), 'num_cycles': 40, 'cases': {}}
    for name, kw in CASES.items():
        pc = numpy_share(kw)
        fs = full_step_on_standin(kw)
        out['cases'][name] = {
            'numpy_share_us_per_cycle': 1e6 * pc, 'numpy_share_env_steps_per_s_per_core_upper_bound': 1.0 / (40 * pc),
            'full_step_on_closed_form_standin_ms': 1e3 * fs, 'full_step_on_closed_form_standin_env_steps_per_s_per_core': 1.0 / fs,
        }
        print(name, out['cases'][name])
    out['note'] = ('numpy_share: unmodified ensure_max_dyn_val + qpos_is_valid + check_mover_collision per cycle x 40 — an upper bound on the '
                   'MuJoCo-backed reference (mj_step, name lookups, noise, observation excluded).  full_step_on_closed_form_standin: the whole '
                   'unmodified step() with mj_step replaced by the closed-form stand-in (NOT MuJoCo).')
    with open(os.path.join(ROOT, 'profiles', 'reference_numpy_share.json'), 'w') as f:
        json.dump(out, f, indent=1)
