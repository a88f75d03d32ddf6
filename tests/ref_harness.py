"""Test-only harness: import the UNMODIFIED reference package from its read-only mount.

The reference (``/root/reference``) needs ``mujoco``, ``gymnasium``, ``pettingzoo`` and ``matplotlib``, none of which are
installed here.  Every pure-NumPy function on the step path (``qpos_is_valid``, ``check_mover_collision``,
``geometry_2D_utils.*``, ``ensure_max_dyn_val``, ``compute_reward`` ...) takes its numeric inputs as arguments, so it runs
unmodified once those four imports are stubbed in ``sys.modules`` (recipe: SURVEY.md appendix A).

This module is used ONLY by ``tests/`` and by ``tests/golden/make_golden.py`` (the script that generated the committed
fixtures).  It never runs on the GPU box (``/root/reference`` does not exist there): ``available()`` is False there and the
tests that need it skip.
"""

from __future__ import annotations

import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

REFERENCE_ROOT = os.environ.get('GPR_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'gymnasium_planar_robotics'))


_installed = False


def install() -> None:
    """Stub the four missing third-party modules and put the reference on ``sys.path`` (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mj = MagicMock(name='mujoco')
    sys.modules['mujoco'] = mj
    sys.modules['mujoco.viewer'] = mj.viewer

    class _Env:
        def reset(self, seed=None, options=None):
            if seed is not None:
                self.np_random = np.random.default_rng(seed)

    class _Box:
        def __init__(self, low, high, shape=None, dtype=None):
            self.low, self.high, self.shape = low, high, shape

        def contains(self, x):
            return True

    class _Dict(dict):
        pass

    stub(
        'gymnasium',
        Env=_Env,
        logger=types.SimpleNamespace(warn=lambda *a, **k: None),
        spaces=types.SimpleNamespace(Box=_Box, Dict=_Dict),
    )
    stub('gymnasium.envs')
    stub('gymnasium.envs.registration', register=lambda **k: None)
    stub('gymnasium.envs.mujoco')

    class _R:
        def __init__(self, *a, **k):
            pass

    stub('gymnasium.envs.mujoco.mujoco_rendering', MujocoRenderer=_R, BaseRender=_R, OffScreenViewer=_R, WindowViewer=_R)
    stub('pettingzoo', ParallelEnv=type('ParallelEnv', (), {}))
    mpl = MagicMock()
    sys.modules.update({'matplotlib': mpl, 'matplotlib.pyplot': mpl.pyplot, 'matplotlib.patches': mpl.patches})
    sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True

    from gymnasium_planar_robotics.envs import basic_envs
    from gymnasium_planar_robotics.utils import mujoco_utils

    # the only MuJoCo-name-dependent bits on the construction path
    mujoco_utils.get_mujoco_type_names = lambda model, obj_type, name_pattern='': []
    basic_envs.BasicPlanarRoboticsEnv._check_mujoco_name_order = lambda self: None
    _installed = True


def make_planning_env(**kwargs):
    """Construct the reference's BenchmarkPlanningEnv (MuJoCo mocked). Only argument-taking NumPy methods are meaningful."""
    install()
    from gymnasium_planar_robotics.envs.planning.benchmark_planning_env import BenchmarkPlanningEnv

    kwargs.setdefault('show_2D_plot', False)
    kwargs.setdefault('render_mode', None)
    env = BenchmarkPlanningEnv(**kwargs)
    env.cycle_time = 0.001
    return env


def geometry():
    install()
    from gymnasium_planar_robotics.utils import geometry_2D_utils

    return geometry_2D_utils
