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
    'shard_range',
    'all_reduce_stats',
    'stats_dict',
)


def register_gymnasium_envs() -> bool:
    """Register 'BenchmarkPlanningEnv-v0' / 'BenchmarkPushingEnv-v0' (max_episode_steps=50) exactly like the reference's
    ``__init__.py:21-41`` and, where gymnasium >= 1.0 offers it, the batched classes as ``vector_entry_point``.  Entry
    points are strings, so nothing heavy is imported.  Returns False when gymnasium is not installed."""
    try:
        from gymnasium.envs.registration import register, registry
    except Exception:
        return False
    for env_id, single, vec in (
        ('BenchmarkPlanningEnv-v0', 'BenchmarkPlanningEnv', 'BenchmarkPlanningVecEnv'),
        ('BenchmarkPushingEnv-v0', 'BenchmarkPushingEnv', 'BenchmarkPushingVecEnv'),
    ):
        if env_id in registry:
            continue
        try:
            register(id=env_id, entry_point=f'{__name__}.envs:{single}', vector_entry_point=f'{__name__}.envs:{vec}', max_episode_steps=50)
        except TypeError:  # gymnasium < 1.0
            register(id=env_id, entry_point=f'{__name__}.envs:{single}', max_episode_steps=50)
    return True


register_gymnasium_envs()  # the reference registers its ids on import (__init__.py:41); a no-op without gymnasium


def bind_to_gpu_numa(device_index: int) -> list[int] | None:
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``device_index`` (one process per GPU:
    page-locked host buffers are then first-touched on the GPU's own NUMA node, so the zero-copy result traffic of
    ``step_host`` does not cross the socket interconnect).  Call it before creating envs.  Returns the core list, or None
    when NVML / affinity control is unavailable (nothing is changed then)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def __getattr__(name):  # torch is imported only when an env class is first touched
    if name in _ENV_NAMES or name == 'envs':
        import importlib

        envs = importlib.import_module(__name__ + '.envs')
        return envs if name == 'envs' else getattr(envs, name)
    raise AttributeError(name)
