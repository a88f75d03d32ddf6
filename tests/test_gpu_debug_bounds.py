"""Memory-safety substitute for the closed compute-sanitizer: the GPR_DEBUG_BOUNDS build of the library
(csrc/Makefile target `debug`, built by __graft_entry__.build()) range-checks every index the kernels derive for global
memory — lane (env, mover) indices, work-list slots / entries / claims of the streamed auto-reset and of the pushing contact
queue, output rows, shared staging indices — into device counters; the host additionally verifies after every step that
the work list was handed back empty and that consumers claimed exactly what was published.  The checks run over RAGGED
sizes (B = 1, 31, 65,537 ...; 1-32 movers; partial CTAs and warps), all auto-reset modes, both envs and the host path.

The debug library is loaded in a subprocess (GPR_B200_LIB is read at import), so the main test process keeps the release
build.  The soak of the streamed auto-reset (tools/soak_stream.py: overlapped vs serialised launches, 1,000+ steps at full
batch sizes) is the same kind of evidence and runs as a slow-marked test.
"""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, 'gymnasium-planar-robotics_b200', 'csrc', 'libgpr_b200_dbg.so')

SCRIPT = r'''
import sys, numpy as np, torch
sys.path[:0] = [%r]
import gymnasium_planar_robotics_b200 as gpr
assert gpr._lib.load().gpr_debug_build() == 1, 'not the bounds-checking build'
DEV = 'cuda:0'
box = {'shape': 'box', 'size': np.array([0.05, 0.04])}
cases = []
for B in (1, 31, 4097, 65537):
    for N, layout, cp in ((1, (3, 3), None), (3, (3, 3), None), (4, (3, 3), None), (8, (5, 5), box), (12, (6, 6), {'shape': 'circle', 'size': 0.06}),
                          (32, (12, 12), {'shape': 'circle', 'size': 0.05})):
        if B == 65537 and N > 8:
            continue
        for mode in ('same_step', 'next_step', 'off'):
            cases.append(('planning', B, N, layout, cp, mode))
    for mode in ('same_step', 'next_step'):
        cases.append(('pushing', B, 1, None, None, mode))
bad = []
for kind, B, N, layout, cp, mode in cases:
    if kind == 'planning':
        env = gpr.BenchmarkPlanningVecEnv(B, np.ones(layout), N, device=DEV, collision_params=cp, learn_jerk=(N %% 2 == 0), autoreset_mode=mode,
                                          max_episode_steps=7, seed=B + N, obstacles=([[0.3, 0.3, 0.02]] if (cp is None and N == 3) else None))
    else:
        env = gpr.BenchmarkPushingVecEnv(B, device=DEV, autoreset_mode=mode, max_episode_steps=7, seed=B)
    env.reset(seed=B + N)
    lim = env.j_max if env.learn_jerk else env.a_max
    g = torch.Generator(device=DEV).manual_seed(B)
    for t in range(12):
        a = (torch.rand((B, env.core.action_dim), device=DEV, generator=g) * 2 - 1) * lim
        if t %% 3 == 2:
            env.step_host(a.cpu().numpy())
        else:
            env.step(a)
        if mode == 'off' and t %% 4 == 3:
            env.reset(options={'mask': torch.rand(B, device=DEV, generator=g) < 0.5})
        err = env.core.debug_errors()
        if any(err):
            bad.append((kind, B, N, mode, t, err))
            break
    env.close()
print('cases', len(cases), 'violations', bad)
sys.exit(1 if bad else 0)
'''


def _run(code, timeout):
    env = dict(os.environ, GPR_B200_LIB=DBG)
    return subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=timeout, env=env)


def test_bounds_checking_build_finds_no_violation_over_ragged_sizes():
    assert os.path.exists(DBG), 'libgpr_b200_dbg.so missing: run __graft_entry__.build() (make -C csrc debug)'
    out = _run(SCRIPT % ROOT, 1500)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert 'violations []' in out.stdout


def test_release_build_reports_its_kind_and_clean_work_lists():
    import numpy as np
    import torch

    import gymnasium_planar_robotics_b200 as gpr

    assert gpr._lib.load().gpr_debug_build() == 0
    env = gpr.BenchmarkPlanningVecEnv(20011, np.ones((3, 3)), 4, device='cuda:0', seed=3)
    env.reset(seed=3)
    for _ in range(5):
        env.step(torch.zeros((20011, 8), device='cuda:0'))
        assert env.core.debug_errors() == [0] * 8  # ([6], [7]: the host-side work-list invariants hold in every build)
    env.close()


@pytest.mark.slow
def test_soak_streamed_autoreset_overlapped_equals_serialised():
    """tools/soak_stream.py: 1,000 steps at 65,536 - 1,048,576 envs, overlapped vs serialised launches, identical outputs."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'soak_stream.py')], capture_output=True, text=True, timeout=3000)
    assert out.returncode == 0 and 'SOAK PASSED' in out.stdout, out.stdout[-3000:]
