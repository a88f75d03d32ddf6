"""ctypes binding of the C ABI in include/gpr.h (``csrc/libgpr_b200.so``).

There is no fallback of any kind: if the CUDA library is missing or does not match the header this module raises, and so
does every env constructor.
"""

from __future__ import annotations

import ctypes
import os

from ._config import GPR_ABI_VERSION, GprConfig, GprOutputs, GprState

_HERE = os.path.dirname(os.path.abspath(__file__))
# GPR_B200_LIB: path of an alternative build of the same library (kernel tuning experiments); never a fallback
LIB_PATH = os.environ.get('GPR_B200_LIB') or os.path.join(_HERE, 'csrc', 'libgpr_b200.so')

# every symbol include/gpr.h declares (tests check the .so exports all of them)
EXPORTED_SYMBOLS = (
    'gpr_config_bytes',
    'gpr_abi_version',
    'gpr_last_error',
    'gpr_create',
    'gpr_destroy',
    'gpr_obs_dim',
    'gpr_goal_dim',
    'gpr_action_dim',
    'gpr_reset',
    'gpr_step',
    'gpr_step_host',
    'gpr_reset_host',
    'gpr_get_state',
    'gpr_set_state',
    'gpr_get_seed',
    'gpr_set_seed',
    'gpr_compute_reward',
    'gpr_compute_reward_f64',
    'gpr_episode_stats',
    'gpr_reset_failures',
    'gpr_kernel_times',
    'gpr_launch_count',
    'gpr_invalidate_outputs',
    'gpr_debug_build',
    'gpr_debug_errors',
)

_lib = None


class GprError(RuntimeError):
    pass


def load():
    """Load the CUDA library (once). Raises if it is absent or its ABI differs from this binding."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GprError(
            f'{LIB_PATH} not found. Build it with `python -c "import __graft_entry__ as g; g.build()"` or '
            f'`make -C {os.path.dirname(LIB_PATH)}`. There is no CPU fallback.'
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
    lib.gpr_config_bytes.restype = ctypes.c_uint32
    lib.gpr_abi_version.restype = ctypes.c_uint32
    lib.gpr_last_error.restype = ctypes.c_char_p
    lib.gpr_create.argtypes = [ctypes.POINTER(GprConfig), i32, ctypes.POINTER(vp)]
    lib.gpr_destroy.argtypes = [vp]
    lib.gpr_destroy.restype = None
    for name in ('gpr_obs_dim', 'gpr_goal_dim', 'gpr_action_dim'):
        getattr(lib, name).argtypes = [vp]
    lib.gpr_reset.argtypes = [vp, vp, i32, u64, vp, vp, vp, ctypes.POINTER(GprOutputs), vp]
    lib.gpr_step.argtypes = [vp, vp, ctypes.POINTER(GprOutputs), vp]
    lib.gpr_step_host.argtypes = [vp, vp, ctypes.POINTER(GprOutputs)]
    lib.gpr_reset_host.argtypes = [vp, i32, u64, ctypes.POINTER(GprOutputs)]
    lib.gpr_get_state.argtypes = [vp, ctypes.POINTER(GprState), vp]
    lib.gpr_set_state.argtypes = [vp, ctypes.POINTER(GprState), vp]
    lib.gpr_get_seed.argtypes = [vp, ctypes.POINTER(u64)]
    lib.gpr_set_seed.argtypes = [vp, u64]
    lib.gpr_compute_reward.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.gpr_compute_reward_f64.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.gpr_episode_stats.argtypes = [vp, vp, i32, vp]
    lib.gpr_reset_failures.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32)]
    lib.gpr_kernel_times.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double)]
    lib.gpr_launch_count.argtypes = [vp]
    lib.gpr_launch_count.restype = u64
    lib.gpr_invalidate_outputs.argtypes = [vp]
    lib.gpr_debug_errors.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32)]
    if lib.gpr_abi_version() != GPR_ABI_VERSION or lib.gpr_config_bytes() != ctypes.sizeof(GprConfig):
        raise GprError(
            f'ABI mismatch: library ABI {lib.gpr_abi_version()} / gpr_config {lib.gpr_config_bytes()} bytes, '
            f'binding ABI {GPR_ABI_VERSION} / {ctypes.sizeof(GprConfig)} bytes. Rebuild csrc/.'
        )
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().gpr_last_error()
        raise GprError(f'gpr error {rc}: {msg.decode() if msg else "?"}')
