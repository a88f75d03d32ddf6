"""B200-native batched simulator for the step path of Gymnasium-Planar-Robotics (GymPR).

Drop-in for ``BenchmarkPlanningEnv-v0`` (Gymnasium form, vectorised form and a PettingZoo-parallel view) and
``BenchmarkPushingEnv-v0``: same env IDs, constructor kwargs and observation / action / info layout as
``gymnasium_planar_robotics`` v1.1.0a2; the physics, collision checks, observation, reward, termination and auto-reset of
``step()`` run as hand-written CUDA kernels (sm_100a) behind the C ABI declared in ``include/gpr.h``.

Importing the package does not need a GPU; constructing an env does (there is no CPU fallback).
"""

from . import _config, _lib  # noqa: F401
from ._config import planning_config, pushing_config  # noqa: F401
from ._lib import GprError  # noqa: F401

__version__ = '0.1.0'

_ENV_NAMES = (
    'BatchedCore',
    'BenchmarkPlanningVecEnv',
    'BenchmarkPushingVecEnv',
    'BenchmarkPlanningEnv',
    'BenchmarkPushingEnv',
    'BenchmarkPlanningParallelEnv',
    'register_gymnasium_envs',
    'shard_range',
    'all_reduce_stats',
    'stats_dict',
)


def __getattr__(name):  # torch is imported only when an env class is first touched
    if name in _ENV_NAMES or name == 'envs':
        import importlib

        envs = importlib.import_module(__name__ + '.envs')
        return envs if name == 'envs' else getattr(envs, name)
    raise AttributeError(name)
