"""Test-only harness: import the UNMODIFIED reference package from its read-only mount.

The reference (``/root/reference``) needs ``mujoco``, ``gymnasium``, ``pettingzoo`` and ``matplotlib``, none of which are
installed here.  Every pure-NumPy function on the step path (``qpos_is_valid``, ``check_mover_collision``,
``geometry_2D_utils.*``, ``ensure_max_dyn_val``, ``compute_reward`` ...) takes its numeric inputs as arguments, so it runs
unmodified once those four imports are stubbed in ``sys.modules`` (recipe: SURVEY.md appendix A).  ``mujoco`` itself is replaced by
the closed-form stand-in of ``tests/mujoco_standin.py``, so the reference's ``reset()`` / ``step()`` run unmodified too.

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

    # `mujoco`: the closed-form stand-in (tests/mujoco_standin.py) — parses the XML the reference generates and integrates
    # the movers with the recurrence the reference's own tests assert against real MuJoCo
    import mujoco_standin

    mj = mujoco_standin.as_module()
    sys.modules['mujoco'] = mj
    sys.modules['mujoco.viewer'] = mj.viewer

    class _Env:
        def reset(self, seed=None, options=None):
            if seed is not None:
                self.np_random = np.random.default_rng(seed)

    class _Box:
        def __init__(self, low, high, shape=None, dtype=None):
            shape = tuple(shape) if shape is not None else np.asarray(low).shape
            self.shape, self.dtype = shape, np.dtype(dtype or np.float64)
            self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), shape).copy()

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

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

    import gymnasium_planar_robotics.envs.basic_envs  # noqa: F401  (nothing of the reference is patched)

    _installed = True


def make_planning_env(**kwargs):
    """Construct the reference's unmodified BenchmarkPlanningEnv on the closed-form MuJoCo stand-in."""
    install()
    from gymnasium_planar_robotics.envs.planning.benchmark_planning_env import BenchmarkPlanningEnv

    kwargs.setdefault('show_2D_plot', False)
    kwargs.setdefault('render_mode', None)
    return BenchmarkPlanningEnv(**kwargs)


def make_pushing_env(**kwargs):
    """The reference's BenchmarkPushingEnv on the stand-in (contact-free trajectories only: the stand-in raises
    ``mujoco.ContactError`` when the mover reaches the object)."""
    install()
    from gymnasium_planar_robotics.envs.manipulation.benchmark_pushing_env import BenchmarkPushingEnv

    kwargs.setdefault('render_mode', None)
    return BenchmarkPushingEnv(**kwargs)


def geometry():
    install()
    from gymnasium_planar_robotics.utils import geometry_2D_utils

    return geometry_2D_utils
