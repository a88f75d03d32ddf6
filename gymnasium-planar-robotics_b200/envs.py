"""Host-side mirror of the reference's env classes for the step path.

Classes
    BatchedCore                 thin owner of one ``gpr_handle`` + persistent torch I/O tensors (device memory plumbing)
    BenchmarkPlanningVecEnv     ``num_envs`` BenchmarkPlanningEnv-v0 instances stepped by one kernel (torch CUDA I/O)
    BenchmarkPushingVecEnv      same for BenchmarkPushingEnv-v0
    BenchmarkPlanningEnv        the reference's single-env Gymnasium API (NumPy float64 in/out) on top of num_envs=1
    BenchmarkPushingEnv
    BenchmarkPlanningParallelEnv PettingZoo-parallel re-keying of the planning state: agents 'mover_0', ... (the reference
                                ships only a non-constructible base class, basic_envs.py:1629-1693)

Constructor kwargs, spaces, observation/action/info layout follow planning/benchmark_planning_env.py:165-259 and
manipulation/benchmark_pushing_env.py:154-247.  torch is used for device memory and streams only.
"""

from __future__ import annotations

import ctypes
import os
from collections import OrderedDict
from typing import Any

import numpy as np
import torch

from . import _lib
from ._config import (
    AUTORESET_NEXT_STEP,
    AUTORESET_SAME_STEP,
    ENV_PLANNING,
    GprConfig,
    GprOutputs,
    GprState,
    planning_config,
    pushing_config,
)

_OUT_FIELDS = [name for name, _ in GprOutputs._fields_]


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def shard_range(total_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """Environments shard by index: rank r owns [base, base+count). No data ever crosses ranks on the step path."""
    per = total_envs // world_size
    rem = total_envs % world_size
    count = per + (1 if rank < rem else 0)
    base = rank * per + min(rank, rem)
    return base, count


def all_reduce_stats(counters: torch.Tensor) -> torch.Tensor:
    """Sum the 6 episode counters over ranks — the ONLY collective of the framework (NCCL for CUDA tensors, gloo for CPU
    tensors in the host-logic tests); a no-op without an initialised process group.  Never on the step path."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def stats_dict(v: list[float]) -> dict[str, float]:
    n = max(v[0], 1.0)
    return {
        'episodes': v[0],
        'mean_return': v[1] / n,
        'mean_length': v[2] / n,
        'success_rate': v[3] / n,
        'mover_collision_rate': v[4] / n,
        'wall_collision_rate': v[5] / n,
    }


class BatchedCore:
    """One ``gpr_handle`` on one device plus the caller-owned I/O tensors the C ABI writes into."""

    def __init__(self, cfg: GprConfig, derived: dict[str, Any], device: torch.device | str | int | None = None):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.GprError('no CUDA device visible: the step path runs on the GPU only (no CPU fallback)')
        device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if device.type != 'cuda':
            raise _lib.GprError(f'device must be a CUDA device, got {device}')
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = device
        self.cfg = cfg
        self.derived = derived
        self.lib = lib
        self.num_envs = int(cfg.num_envs)
        self.num_movers = int(cfg.num_movers)
        self.kind = int(cfg.env_kind)
        handle = ctypes.c_void_p()
        _lib.check(lib.gpr_create(ctypes.byref(cfg), device.index, ctypes.byref(handle)))
        self.handle = handle
        self.obs_dim = lib.gpr_obs_dim(handle)
        self.goal_dim = lib.gpr_goal_dim(handle)
        self.action_dim = lib.gpr_action_dim(handle)
        B = self.num_envs
        f32, u8 = torch.float32, torch.uint8
        self.f64_outputs = bool(int(cfg.output_flags) & 2)  # GPR_OUT_FLOAT64
        fo = torch.float64 if self.f64_outputs else torch.float32

        def z(shape, dtype):
            return torch.zeros(shape, dtype=dtype, device=device)

        self.buf = {
            'observation': z((B, self.obs_dim), fo),
            'achieved_goal': z((B, self.goal_dim), fo),
            'desired_goal': z((B, self.goal_dim), fo),
            'reward': z((B,), f32),
            'terminated': z((B,), u8),
            'truncated': z((B,), u8),
            'is_success': z((B,), u8),
            'mover_collision': z((B,), u8),
            'wall_collision': z((B,), u8),
        }
        if int(getattr(cfg, 'num_obstacles', 0)) > 0:  # (the reference's info has no such key: only present with obstacles)
            self.buf['other_collision'] = z((B,), u8)
        if int(cfg.autoreset_mode) == AUTORESET_SAME_STEP:
            self.buf['final_observation'] = z((B, self.obs_dim), fo)
            self.buf['final_achieved_goal'] = z((B, self.goal_dim), fo)
            self.buf['final_desired_goal'] = z((B, self.goal_dim), fo)
        self._out = GprOutputs()
        for name in _OUT_FIELDS:
            setattr(self._out, name, _ptr(self.buf.get(name)))
        self._action = z((B, self.action_dim), f32)
        self._host: dict[str, np.ndarray] | None = None
        self._pinned_actions: list[torch.Tensor] = []

    # ------------------------------------------------------------------------------------------------------------ util
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self) -> None:
        if getattr(self, 'handle', None) is not None and self.handle.value:
            self.lib.gpr_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dev_f64(self, x, shape) -> torch.Tensor | None:
        if x is None:
            return None
        t = torch.as_tensor(x, dtype=torch.float64, device=self.device).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f'expected shape {tuple(shape)}, got {tuple(t.shape)}')
        return t

    # ----------------------------------------------------------------------------------------------------------- reset
    def reset(self, seed: int | None = None, mask=None, start_pos=None, goal_pos=None, object_pos=None) -> None:
        """basic_envs.py:1770-1833 for the masked envs (all if ``mask`` is None); results land in ``self.buf``."""
        B, N = self.num_envs, self.num_movers
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            if tuple(m.shape) != (B,):
                raise ValueError('mask must have shape (num_envs,)')
        st = self._dev_f64(start_pos, (B, N, 2))
        gl = self._dev_f64(goal_pos, (B, self.goal_dim // 2, 2))
        ob = self._dev_f64(object_pos, (B, 2)) if object_pos is not None else None
        _lib.check(
            self.lib.gpr_reset(
                self.handle, _ptr(m), int(seed is not None), int(seed or 0) & (2**64 - 1), _ptr(st), _ptr(gl), _ptr(ob),
                ctypes.byref(self._out), self._stream(),
            )
        )

    # ------------------------------------------------------------------------------------------------------------ step
    def step(self, action: torch.Tensor) -> None:
        """basic_envs.py:1835-1950 for every env; ``action``: float32 CUDA tensor (num_envs, action_dim)."""
        if not isinstance(action, torch.Tensor):
            action = torch.as_tensor(action)
        if action.device != self.device or action.dtype != torch.float32 or not action.is_contiguous():
            self._action.copy_(action.reshape(self.num_envs, self.action_dim), non_blocking=True)
            action = self._action
        if tuple(action.shape) != (self.num_envs, self.action_dim):
            raise ValueError(f'action dim != action_space dim: expected {(self.num_envs, self.action_dim)}, got {tuple(action.shape)}')
        _lib.check(self.lib.gpr_step(self.handle, action.data_ptr(), ctypes.byref(self._out), self._stream()))

    def pinned_action_buffer(self) -> np.ndarray:
        """A page-locked float32 (num_envs, action_dim) array to fill with actions: ``step_host`` reads it in place."""
        t = torch.zeros((self.num_envs, self.action_dim), dtype=torch.float32, pin_memory=True)
        self._pinned_actions.append(t)  # torch owns the memory, NumPy views it
        return t.numpy()

    def step_host(self, action: np.ndarray) -> dict[str, np.ndarray]:
        """The same step called with HOST buffers (``gpr_step_host``): NumPy action in, NumPy results out.  The result
        arrays are page-locked, so the kernels write them in place over PCIe (no staging copy); an action array obtained
        from ``pinned_action_buffer`` is likewise read in place, any other one is staged through pinned memory."""
        if action.dtype == np.float32 and action.flags.c_contiguous and action.size == self.num_envs * self.action_dim:
            a = action  # (the common case: no NumPy calls on the per-step path)
        else:
            a = np.ascontiguousarray(action, dtype=np.float32).reshape(self.num_envs, self.action_dim)
        if self._host is None:
            # page-locked result arrays (torch owns the memory, NumPy views it)
            self._host_pin = {k: torch.zeros(tuple(v.shape), dtype=v.dtype, pin_memory=True) for k, v in self.buf.items()}
            self._host = {k: t.numpy() for k, t in self._host_pin.items()}
            # terminal observations may come back as a LIST (gpr_outputs.final_index; compact transport with several ranks
            # on one host): room for the index and the count
            self._final_index = np.zeros(self.num_envs, dtype=np.int32)
            self._final_count = np.full(1, 0xFFFFFFFF, dtype=np.uint32)  # (UINT32_MAX = dense rows; the library sets it per call)
            self._host_out = GprOutputs()
            for name in _OUT_FIELDS:
                setattr(self._host_out, name, self._host[name].ctypes.data if name in self._host else None)
            if 'final_observation' in self._host:
                self._host_out.final_index = self._final_index.ctypes.data
                self._host_out.final_count = self._final_count.ctypes.data
        _lib.check(self.lib.gpr_step_host(self.handle, a.ctypes.data, ctypes.byref(self._host_out)))
        return self._host

    @property
    def host_final_count(self) -> int | None:
        """After ``step_host``: None when the terminal observations were delivered densely (row = env), else the number of
        list rows (row s of the final_* arrays belongs to env ``_final_index[s]``)."""
        if getattr(self, '_final_count', None) is None or self._host is None or 'final_observation' not in self._host:
            return None
        c = int(self._final_count[0])
        return None if c == 0xFFFFFFFF else c

    # ----------------------------------------------------------------------------------------------------------- state
    def _state_tensors(self) -> dict[str, torch.Tensor]:
        B, N = self.num_envs, self.num_movers
        f64 = torch.float64

        def z(shape, dtype=f64):
            return torch.zeros(shape, dtype=dtype, device=self.device)

        st = {
            'pos': z((B, N, 2)),
            'vel': z((B, N, 2)),
            'acc': z((B, N, 2)),
            'goal': z((B, self.goal_dim // 2, 2)),
            'elapsed_steps': z((B,), torch.int32),
            'rng_counter': z((B,), torch.int32),  # uint32 bit pattern
            'needs_reset': z((B,), torch.uint8),
            'episode_return': z((B,), torch.float32),
        }
        if self.kind != ENV_PLANNING:
            st.update(act=z((B, 2)), mover_rot=z((B, 3)), object_pos=z((B, 4)), object_vel=z((B, 3)),
                      contact_warm=z((B, 13), torch.float32))
        return st

    def get_state(self) -> dict[str, torch.Tensor]:
        """Float64 SoA state (MjData.qpos/qvel/qacc/act equivalents) — parity tests and checkpointing."""
        st = self._state_tensors()
        s = GprState()
        for k, t in st.items():
            setattr(s, k, t.data_ptr())
        _lib.check(self.lib.gpr_get_state(self.handle, ctypes.byref(s), self._stream()))
        return st

    def set_state(self, state: dict[str, Any]) -> None:
        ref = self._state_tensors()
        s = GprState()
        keep = []
        for k, v in state.items():
            if k not in ref:
                raise KeyError(k)
            t = torch.as_tensor(v, device=self.device).to(ref[k].dtype).contiguous()
            if tuple(t.shape) != tuple(ref[k].shape):
                raise ValueError(f'{k}: expected {tuple(ref[k].shape)}, got {tuple(t.shape)}')
            keep.append(t)
            setattr(s, k, t.data_ptr())
        _lib.check(self.lib.gpr_set_state(self.handle, ctypes.byref(s), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # `keep` must outlive the copies

    # ---------------------------------------------------------------------------------------------------------- extras
    def compute_reward(self, achieved_goal, desired_goal, mover_collision=None, wall_collision=None):
        """Batched compute_reward + compute_terminated on device (HER relabelling, planning:459-534 / pushing:457-527)."""
        # float64 goals (the single-env classes; GPR_OUT_FLOAT64) are compared in float64: the step's own decision
        ag = torch.as_tensor(achieved_goal, device=self.device)
        f64 = ag.dtype == torch.float64
        gt = torch.float64 if f64 else torch.float32
        ag = ag.to(gt).reshape(-1, self.goal_dim).contiguous()
        dg = torch.as_tensor(desired_goal, device=self.device).to(gt).reshape(-1, self.goal_dim).contiguous()
        b = ag.shape[0]
        mc = None if mover_collision is None else torch.as_tensor(mover_collision, device=self.device).to(torch.uint8).contiguous()
        wc = None if wall_collision is None else torch.as_tensor(wall_collision, device=self.device).to(torch.uint8).contiguous()
        r = torch.empty(b, dtype=torch.float32, device=self.device)
        t = torch.empty(b, dtype=torch.uint8, device=self.device)
        fn = self.lib.gpr_compute_reward_f64 if f64 else self.lib.gpr_compute_reward
        _lib.check(fn(self.handle, b, ag.data_ptr(), dg.data_ptr(), _ptr(mc), _ptr(wc), r.data_ptr(), t.data_ptr(), self._stream()))
        return r, t.bool()

    def episode_stats(self, reset: bool = True, all_reduce: bool = False) -> dict[str, float]:
        """On-device episode counters; ``all_reduce=True`` sums them over ranks with NCCL (the only collective there is)."""
        out = torch.zeros(6, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.gpr_episode_stats(self.handle, out.data_ptr(), int(reset), self._stream()))
        if all_reduce:
            all_reduce_stats(out)
        return stats_dict(out.tolist())

    def kernel_times(self, enable: bool = True) -> dict[str, float]:
        """Per-kernel device time of the ``step`` calls since the last query (``gpr_kernel_times``): CUDA events recorded
        around each kernel on the launching stream.  The first call just switches the recording on."""
        out = (ctypes.c_double * 3)()
        _lib.check(self.lib.gpr_kernel_times(self.handle, int(enable), out))
        n = max(out[2], 1.0)
        return {'steps': out[2], 'step_kernel_ms': out[0] / n, 'autoreset_kernel_ms': out[1] / n}

    def debug_errors(self) -> list[int]:
        """``gpr_debug_errors``: the 8 counters of the bounds-checking build (all 0 in a release build except the two
        host-side work-list invariants [6], [7])."""
        c = (ctypes.c_uint32 * 8)()
        _lib.check(self.lib.gpr_debug_errors(self.handle, c))
        return [int(x) for x in c]

    def invalidate_outputs(self) -> None:
        _lib.check(self.lib.gpr_invalidate_outputs(self.handle))

    def reset_failures(self) -> int:
        c = ctypes.c_uint32()
        _lib.check(self.lib.gpr_reset_failures(self.handle, ctypes.byref(c)))
        return int(c.value)

    @property
    def launch_count(self) -> int:
        return int(self.lib.gpr_launch_count(self.handle))


# ----------------------------------------------------------------------------------------------------------------------
# spaces (gymnasium is optional: imported lazily, never required)
# ----------------------------------------------------------------------------------------------------------------------
class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.asarray(low).shape
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self, rng: np.random.Generator | None = None):
        rng = np.random.default_rng() if rng is None else rng
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return rng.uniform(lo, hi).astype(self.dtype)


def _spaces(derived, action_lim, num_goal_bodies):
    try:  # real gymnasium spaces when the package is importable
        import gymnasium as gym

        Box, Dict = gym.spaces.Box, gym.spaces.Dict
    except Exception:
        Box, Dict = _Box, dict
    low_goals = np.zeros((num_goal_bodies * 2,))
    high_goals = np.array(list(derived['high_goals']) * num_goal_bodies)
    obs = Dict(
        {
            'observation': Box(low=-np.inf, high=np.inf, shape=(derived['obs_dim'],), dtype=np.float64),
            'achieved_goal': Box(low=low_goals, high=high_goals, dtype=np.float64),
            'desired_goal': Box(low=low_goals, high=high_goals, dtype=np.float64),
        }
    )
    act = Box(low=-action_lim, high=action_lim, shape=(derived['action_dim'],), dtype=np.float64)
    return obs, act


# ----------------------------------------------------------------------------------------------------------------------
# vector envs
# ----------------------------------------------------------------------------------------------------------------------
class _VecEnvBase:
    """Gymnasium-VectorEnv-shaped API over torch CUDA tensors.  Returned tensors are views of persistent buffers that the
    next ``step``/``reset`` overwrites (clone them to keep a transition)."""

    metadata = {'render_modes': [], 'autoreset_mode': 'same_step'}
    _config_fn = None

    def __init__(self, num_envs: int, device=None, rank: int | None = None, world_size: int | None = None,
                 total_envs: int | None = None, **kwargs):
        # multi-GPU: env indices shard across ranks; the global index keys the RNG so results do not depend on the split
        if total_envs is not None:
            rank = int(os.environ.get('RANK', 0)) if rank is None else rank
            world_size = int(os.environ.get('WORLD_SIZE', 1)) if world_size is None else world_size
            base, num_envs = shard_range(total_envs, rank, world_size)
            kwargs['env_index_base'] = base
        cfg, derived = type(self)._config_fn(num_envs=num_envs, **kwargs)
        self.core = BatchedCore(cfg, derived, device)
        self.cfg, self.derived = cfg, derived
        self.num_envs = num_envs
        self.num_movers = int(cfg.num_movers)
        self.device = self.core.device
        self.learn_jerk = bool(cfg.learn_jerk)
        self.v_max, self.a_max, self.j_max = cfg.v_max, cfg.a_max, cfg.j_max
        self.threshold_pos = cfg.threshold_pos
        self.num_cycles = int(cfg.num_cycles)
        self.cycle_time = cfg.cycle_time
        lim = cfg.j_max if cfg.learn_jerk else cfg.a_max
        self.single_observation_space, self.single_action_space = _spaces(derived, lim, derived['goal_dim'] // 2)
        self.observation_space, self.action_space = self.single_observation_space, self.single_action_space
        self.metadata = dict(self.metadata)
        self.metadata['autoreset_mode'] = {0: 'disabled', 1: 'same_step', 2: 'next_step'}[int(cfg.autoreset_mode)]
        self.closed = False
        self._host_ret = None

    # -- views
    def _obs(self) -> dict[str, torch.Tensor]:
        b = self.core.buf
        return {'observation': b['observation'], 'achieved_goal': b['achieved_goal'], 'desired_goal': b['desired_goal']}

    def _info(self, with_final: bool) -> dict[str, torch.Tensor]:
        b = self.core.buf
        info = {
            'is_success': b['is_success'].bool(),
            'mover_collision': b['mover_collision'].bool(),
            'wall_collision': b['wall_collision'].bool(),
        }
        if 'other_collision' in b:
            info['other_collision'] = b['other_collision'].bool()
        if with_final and 'final_observation' in b:
            # SAME_STEP autoreset: rows are valid where terminated | truncated
            info['final_obs'] = {
                'observation': b['final_observation'],
                'achieved_goal': b['final_achieved_goal'],
                'desired_goal': b['final_desired_goal'],
            }
        return info

    # -- API
    def reset(self, seed: int | None = None, options: dict[str, Any] | None = None):
        """reset(seed, options) -> (obs, info).  options: 'mask' (bool[num_envs]), 'mover_start_xy_pos' (B,N,2),
        'mover_goal_xy_pos' / 'object_goal_xy_pos', 'object_start_xy_pos' inject positions instead of sampling."""
        options = {} if options is None else options
        goal = options.get('mover_goal_xy_pos', options.get('object_goal_xy_pos'))
        if goal is not None:
            goal = torch.as_tensor(goal, dtype=torch.float64).reshape(self.num_envs, -1, 2)
        self.core.reset(seed=seed, mask=options.get('mask'), start_pos=options.get('mover_start_xy_pos'), goal_pos=goal,
                        object_pos=options.get('object_start_xy_pos'))
        return self._obs(), self._info(False)

    def step(self, action: torch.Tensor):
        self.core.step(action)
        b = self.core.buf
        return self._obs(), b['reward'], b['terminated'].bool(), b['truncated'].bool(), self._info(True)

    def step_host(self, action: np.ndarray):
        """step() for callers that live on the host (NumPy in / NumPy out through ``gpr_step_host``).

        ``info['final_obs']`` holds the terminal observations of the envs that finished in this step (SAME_STEP auto-reset).
        Normally its arrays are dense — row = env, valid where terminated | truncated.  When several ranks share one host
        (compact transport) they come as a list instead, like gymnasium's vector envs report them only for
        the sub-envs that finished: ``info['final_obs']['index']`` (int32 env indices) is then present and row s of the other
        arrays belongs to env ``index[s]``; ``final_obs_dense(info)`` converts either form to the dense one."""
        h = self.core.step_host(action)
        n_list = self.core.host_final_count
        if n_list is not None:
            obs = {'observation': h['observation'], 'achieved_goal': h['achieved_goal'], 'desired_goal': h['desired_goal']}
            info = {k: h[k].view(np.bool_) for k in ('is_success', 'mover_collision', 'wall_collision', 'other_collision') if k in h}
            info['final_obs'] = {k: h['final_' + k][:n_list] for k in ('observation', 'achieved_goal', 'desired_goal')}
            info['final_obs']['index'] = self.core._final_index[:n_list]
            return obs, h['reward'], h['terminated'].view(np.bool_), h['truncated'].view(np.bool_), info
        if self._host_ret is None:
            # the result arrays are persistent (rewritten in place every step), so the returned structure is built once;
            # the flag arrays hold 0/1 bytes and are viewed as bool instead of converted
            obs = {'observation': h['observation'], 'achieved_goal': h['achieved_goal'], 'desired_goal': h['desired_goal']}
            info = {k: h[k].view(np.bool_) for k in ('is_success', 'mover_collision', 'wall_collision', 'other_collision') if k in h}
            if 'final_observation' in h:
                info['final_obs'] = {k: h['final_' + k] for k in ('observation', 'achieved_goal', 'desired_goal')}
            self._host_ret = (obs, h['reward'], h['terminated'].view(np.bool_), h['truncated'].view(np.bool_), info)
        return self._host_ret

    def final_obs_dense(self, info) -> dict[str, np.ndarray]:
        """``info['final_obs']`` of ``step_host`` in the dense form (row = env; rows of other envs zero) whichever form it has."""
        fo = info['final_obs']
        if 'index' not in fo:
            return fo
        out = {}
        for k in ('observation', 'achieved_goal', 'desired_goal'):
            d = np.zeros((self.num_envs,) + fo[k].shape[1:], dtype=fo[k].dtype)
            d[fo['index']] = fo[k]
            out[k] = d
        return out

    def compute_reward(self, achieved_goal, desired_goal, info=None):
        mc, wc = self._split_info(info)
        return self.core.compute_reward(achieved_goal, desired_goal, mc, wc)[0]

    def compute_terminated(self, achieved_goal, desired_goal, info=None):
        mc, wc = self._split_info(info)
        return self.core.compute_reward(achieved_goal, desired_goal, mc, wc)[1]

    def compute_truncated(self, achieved_goal, desired_goal, info=None):
        ag = torch.as_tensor(achieved_goal)
        return torch.zeros(ag.reshape(-1, self.core.goal_dim).shape[0], dtype=torch.bool, device=self.device)

    @staticmethod
    def _split_info(info):
        if info is None:
            return None, None
        if isinstance(info, dict):
            mc = info.get('mover_collision')
            if info.get('other_collision') is not None:  # an obstacle hit counts like a mover collision in the reward
                oc = info['other_collision']
                mc = oc if mc is None else (torch.as_tensor(mc).bool() | torch.as_tensor(oc).bool().to(torch.as_tensor(mc).device))
            return mc, info.get('wall_collision')
        # array of per-transition dicts, as the reference accepts (planning:666-688)
        mc = np.array([bool(i['mover_collision']) or bool(i.get('other_collision', False)) for i in info], dtype=np.uint8)
        wc = np.array([bool(i['wall_collision']) for i in info], dtype=np.uint8)
        return mc, wc

    def get_state(self):
        return self.core.get_state()

    def set_state(self, state):
        self.core.set_state(state)

    def state_dict(self):
        """Checkpoint: the SoA state tensors, the per-env RNG event counters and the RNG key in use.
        ``load_state_dict`` on an env built with the same kwargs continues the run bit for bit."""
        st = {k: v.cpu() for k, v in self.core.get_state().items()}
        seed = ctypes.c_uint64()
        _lib.check(self.core.lib.gpr_get_seed(self.core.handle, ctypes.byref(seed)))
        st['seed'] = int(seed.value)
        return st

    def load_state_dict(self, state):
        state = dict(state)
        seed = state.pop('seed')
        self.core.set_state(state)
        _lib.check(self.core.lib.gpr_set_seed(self.core.handle, ctypes.c_uint64(int(seed))))

    def debug_view(self, env_index: int = 0, ppm: float = 400.0) -> np.ndarray:
        """(H, W, 3) uint8 picture of ONE env — tiles, movers, collision shapes, goals, velocities (``debug_view.py``; the
        stand-in for the reference's ``Matplotlib2DViewer``, utils/rendering.py:283-507, off the step path)."""
        from . import debug_view

        return debug_view.view_of_env(self, env_index, ppm)

    def episode_stats(self, reset: bool = True, all_reduce: bool = False):
        return self.core.episode_stats(reset, all_reduce)

    def close(self):
        if not self.closed:
            self.core.close()
            self.closed = True


class BenchmarkPlanningVecEnv(_VecEnvBase):
    """``num_envs`` copies of BenchmarkPlanningEnv-v0 (kwargs of planning/benchmark_planning_env.py:165-185, plus
    ``cycle_time``, ``max_episode_steps``, ``autoreset_mode``, ``seed``)."""

    _config_fn = staticmethod(planning_config)

    def __init__(self, num_envs: int, layout_tiles: np.ndarray, num_movers: int, device=None, **kwargs):
        super().__init__(num_envs, device=device, layout_tiles=layout_tiles, num_movers=num_movers, **kwargs)
        self.goals = None


class BenchmarkPushingVecEnv(_VecEnvBase):
    """``num_envs`` copies of BenchmarkPushingEnv-v0 (kwargs of manipulation/benchmark_pushing_env.py:154-169)."""

    _config_fn = staticmethod(pushing_config)

    def __init__(self, num_envs: int, device=None, **kwargs):
        super().__init__(num_envs, device=device, **kwargs)


# ----------------------------------------------------------------------------------------------------------------------
# single-env Gymnasium form (the reference's exact call signatures, NumPy float64)
# ----------------------------------------------------------------------------------------------------------------------
class _SingleEnvBase:
    metadata = {'render_modes': []}
    _vec_cls = None

    def __init__(self, **kwargs):
        kwargs.setdefault('autoreset_mode', 'off')  # like the reference: the caller (or gymnasium's wrappers) resets
        kwargs.setdefault('max_episode_steps', 0)   # TimeLimit(50) is added by gymnasium at registration (__init__.py:28)
        kwargs.setdefault('float64_outputs', True)  # the reference's spaces are float64: observations / goals unrounded
        self._vec = type(self)._vec_cls(num_envs=1, **kwargs)
        v = self._vec
        self.observation_space, self.action_space = v.single_observation_space, v.single_action_space
        self.num_movers, self.learn_jerk = v.num_movers, v.learn_jerk
        self.v_max, self.a_max, self.j_max = v.v_max, v.a_max, v.j_max
        self.threshold_pos, self.num_cycles, self.cycle_time = v.threshold_pos, v.num_cycles, v.cycle_time
        self.render_mode = None
        self.np_random = np.random.default_rng()

    def _np_obs(self, obs):
        return OrderedDict((k, obs[k][0].double().cpu().numpy()) for k in ('observation', 'achieved_goal', 'desired_goal'))

    @staticmethod
    def _np_info(info):
        return {k: bool(info[k][0]) for k in ('is_success', 'mover_collision', 'wall_collision', 'other_collision') if k in info}

    def reset(self, seed: int | None = None, options: dict[str, Any] | None = None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        opts = {}
        if options:
            for k in ('mover_start_xy_pos', 'mover_goal_xy_pos', 'object_goal_xy_pos', 'object_start_xy_pos'):
                if k in options:
                    opts[k] = np.asarray(options[k], dtype=np.float64)[None]
        obs, info = self._vec.reset(seed=seed, options=opts)
        return self._np_obs(obs), self._np_info(info)

    def step(self, action):
        action = np.asarray(action, dtype=np.float64)
        assert action.shape == self.action_space.shape, 'action dim != action_space dim'
        obs, r, term, trunc, info = self._vec.step(torch.as_tensor(action[None], dtype=torch.float32))
        return self._np_obs(obs), float(r[0]), bool(term[0]), bool(trunc[0]), self._np_info(info)

    def compute_reward(self, achieved_goal, desired_goal, info=None):
        r = self._vec.compute_reward(np.asarray(achieved_goal), np.asarray(desired_goal), info).double().cpu().numpy()
        return r if np.asarray(achieved_goal).ndim > 1 and r.shape[0] > 1 else float(r[0])

    def compute_terminated(self, achieved_goal, desired_goal, info=None):
        t = self._vec.compute_terminated(np.asarray(achieved_goal), np.asarray(desired_goal), info).cpu().numpy()
        return t if np.asarray(achieved_goal).ndim > 1 and t.shape[0] > 1 else bool(t[0])

    def compute_truncated(self, achieved_goal, desired_goal, info=None):
        ag = np.asarray(achieved_goal)
        batch = ag.shape[0] if ag.ndim > 1 else 1
        return np.array([False] * batch) if batch > 1 else False

    def render(self):
        return None

    def debug_view(self, ppm: float = 400.0) -> np.ndarray:
        return self._vec.debug_view(0, ppm)

    def close(self):
        self._vec.close()

    @property
    def unwrapped(self):
        return self


class BenchmarkPlanningEnv(_SingleEnvBase):
    """Drop-in for ``gymnasium_planar_robotics.BenchmarkPlanningEnv`` (render_mode must be None)."""

    _vec_cls = BenchmarkPlanningVecEnv

    def __init__(self, layout_tiles: np.ndarray, num_movers: int, show_2D_plot: bool = False, **kwargs):
        kwargs.setdefault('render_mode', None)
        super().__init__(layout_tiles=layout_tiles, num_movers=num_movers, show_2D_plot=show_2D_plot, **kwargs)


class BenchmarkPushingEnv(_SingleEnvBase):
    """Drop-in for ``gymnasium_planar_robotics.BenchmarkPushingEnv`` (render_mode must be None)."""

    _vec_cls = BenchmarkPushingVecEnv

    def __init__(self, **kwargs):
        kwargs.setdefault('render_mode', None)
        super().__init__(**kwargs)


# ----------------------------------------------------------------------------------------------------------------------
# PettingZoo-parallel view
# ----------------------------------------------------------------------------------------------------------------------
class BenchmarkPlanningParallelEnv:
    """PettingZoo ``ParallelEnv``-shaped view of the batched planning env: one agent per mover.

    ``agents = possible_agents = ['mover_0', ...]`` (naming: basic_envs.py:878, 1692-1693).  Every per-agent tensor is a
    strided VIEW into the same output buffers the kernel writes, shaped (num_envs, ...); the team reward and the
    termination flags are shared by all agents (the single-agent env is the oracle: same movers, same state).
    """

    metadata = {'name': 'BenchmarkPlanningParallelEnv-v0', 'render_modes': []}

    def __init__(self, num_envs: int, layout_tiles: np.ndarray, num_movers: int, device=None, **kwargs):
        self._vec = BenchmarkPlanningVecEnv(num_envs, layout_tiles, num_movers, device=device, **kwargs)
        self.num_envs = num_envs
        self.possible_agents = [f'mover_{i}' for i in range(num_movers)]
        self.agents = list(self.possible_agents)
        self._action = torch.zeros((num_envs, num_movers, 2), dtype=torch.float32, device=self._vec.device)
        lim = self._vec.j_max if self._vec.learn_jerk else self._vec.a_max
        J = int(self._vec.learn_jerk)
        self._obs_space = {'observation': _Box(-np.inf, np.inf, (2 * (1 + J),)), 'achieved_goal': _Box(0, np.inf, (2,)),
                           'desired_goal': _Box(0, np.inf, (2,))}
        self._act_space = _Box(-lim, lim, (2,))

    def observation_space(self, agent):
        return self._obs_space

    def action_space(self, agent):
        return self._act_space

    def _views(self):
        v = self._vec
        B, N, J = v.num_envs, v.num_movers, int(v.learn_jerk)
        b = v.core.buf
        o = b['observation'].view(B, 1 + J, N, 2)
        ag = b['achieved_goal'].view(B, N, 2)
        dg = b['desired_goal'].view(B, N, 2)
        return {
            a: {'observation': o[:, :, i, :].reshape(B, 2 * (1 + J)) if J else o[:, 0, i, :], 'achieved_goal': ag[:, i], 'desired_goal': dg[:, i]}
            for i, a in enumerate(self.possible_agents)
        }

    def reset(self, seed: int | None = None, options: dict | None = None):
        _, info = self._vec.reset(seed=seed, options=options)
        self.agents = list(self.possible_agents)
        return self._views(), {a: info for a in self.agents}

    def step(self, actions: dict[str, torch.Tensor]):
        for i, a in enumerate(self.possible_agents):
            self._action[:, i, :].copy_(torch.as_tensor(actions[a]), non_blocking=True)
        _, r, term, trunc, info = self._vec.step(self._action.view(self.num_envs, -1))
        ag = self.agents
        return self._views(), {a: r for a in ag}, {a: term for a in ag}, {a: trunc for a in ag}, {a: info for a in ag}

    def state(self):
        return self._vec.get_state()

    def debug_view(self, env_index: int = 0, ppm: float = 400.0) -> np.ndarray:
        return self._vec.debug_view(env_index, ppm)

    def close(self):
        self._vec.close()
