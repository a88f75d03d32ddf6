"""B200-native batched simulator for the step path of Gymnasium-Planar-Robotics (GymPR).

Drop-in for ``BenchmarkPlanningEnv-v0`` (Gymnasium form, vectorised form and a PettingZoo-parallel view) and
``BenchmarkPushingEnv-v0``: same env IDs, constructor kwargs and observation / action / info layout as
``gymnasium_planar_robotics`` v1.1.0a2; the physics, collision checks, observation, reward, termination and auto-reset of
``step()`` run as hand-written CUDA kernels (sm_100a) behind the C ABI declared in ``include/gpr.h``.
"""

from . import _config  # noqa: F401
from ._config import planning_config, pushing_config  # noqa: F401

__version__ = '0.1.0'
