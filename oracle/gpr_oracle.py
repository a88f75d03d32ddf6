"""ctypes wrapper around oracle/_ref/libgpr_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this
module.  The product package never does: it fails loudly when its CUDA extension is missing instead of falling back.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_ref', 'libgpr_oracle.so')
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement (oracle/Makefile). Building the checker is not using it."""
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ('gpr_oracle.c', '../include/gpr.h', '../include/gpr_rng.h', '../include/gpr_push_physics.h')
    ):
        subprocess.run(['make', '-C', _HERE, '-s'] + (['-B'] if force else []), check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.gpro_assert_trips.restype = ctypes.c_int64
        _lib.gpro_config_bytes.restype = ctypes.c_uint32
    return _lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


_D = ctypes.c_double
_F = ctypes.c_float
_U8 = ctypes.c_uint8


class OState(ctypes.Structure):
    _fields_ = [
        ('pos', ctypes.c_void_p),
        ('vel', ctypes.c_void_p),
        ('acc', ctypes.c_void_p),
        ('goal', ctypes.c_void_p),
        ('elapsed_steps', ctypes.c_void_p),
        ('rng_counter', ctypes.c_void_p),
        ('needs_reset', ctypes.c_void_p),
        ('act', ctypes.c_void_p),
        ('mover_rot', ctypes.c_void_p),
        ('object_pos', ctypes.c_void_p),
        ('object_vel', ctypes.c_void_p),
        ('contact_warm', ctypes.c_void_p),
    ]


class OOutputs(ctypes.Structure):
    _fields_ = [
        ('observation', ctypes.c_void_p),
        ('achieved_goal', ctypes.c_void_p),
        ('desired_goal', ctypes.c_void_p),
        ('reward', ctypes.c_void_p),
        ('terminated', ctypes.c_void_p),
        ('truncated', ctypes.c_void_p),
        ('is_success', ctypes.c_void_p),
        ('mover_collision', ctypes.c_void_p),
        ('wall_collision', ctypes.c_void_p),
        ('final_observation', ctypes.c_void_p),
        ('final_achieved_goal', ctypes.c_void_p),
        ('final_desired_goal', ctypes.c_void_p),
        ('reset_failed', ctypes.c_void_p),
        ('other_collision', ctypes.c_void_p),
    ]


# ---- function-level wrappers (one call = the reference's function on the same arguments) ----------------------------
def ensure_max_dyn_val(cur, max_value, deriv, dt):
    cur = np.ascontiguousarray(cur, dtype=np.float64)
    deriv = np.ascontiguousarray(deriv, dtype=np.float64)
    nv, nd = np.zeros(2), np.zeros(2)
    lib().gpro_ensure_max_dyn_val(_p(cur, _D), _D(max_value), _p(deriv, _D), _D(dt), _p(nv, _D), _p(nd, _D))
    return nv, nd


def segments_intersect(p1, p2, q1, q2) -> bool:
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (p1, p2, q1, q2)]
    return bool(lib().gpro_segments_intersect(*[_p(x, _D) for x in a]))


def rect_vertices(qpos, size):
    qpos = np.ascontiguousarray(qpos, dtype=np.float64)
    size = np.ascontiguousarray(size, dtype=np.float64)
    vx, vy = np.zeros(4), np.zeros(4)
    lib().gpro_rect_vertices(_p(qpos, _D), _p(size, _D), _p(vx, _D), _p(vy, _D))
    return np.stack([vx, vy])


def rectangles_intersect(q1, q2, s1, s2) -> bool:
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (q1, q2, s1, s2)]
    return bool(lib().gpro_rectangles_intersect(*[_p(x, _D) for x in a]))


def qpos_is_valid(cfg, qpos, csize) -> np.ndarray:
    qpos = np.ascontiguousarray(qpos, dtype=np.float64)
    n = qpos.shape[0]
    csize = np.ascontiguousarray(np.broadcast_to(np.asarray(csize, dtype=np.float64).reshape(n, -1), (n, 2)))
    out = np.zeros(n, dtype=np.int32)
    lib().gpro_qpos_is_valid(ctypes.byref(cfg), n, _p(qpos, _D), _p(csize, _D), _p(out, ctypes.c_int32))
    return out


def check_mover_collision(cfg, qpos, csize) -> bool:
    qpos = np.ascontiguousarray(qpos, dtype=np.float64)
    n = qpos.shape[0]
    csize = np.ascontiguousarray(np.broadcast_to(np.asarray(csize, dtype=np.float64).reshape(n, -1), (n, 2)))
    return bool(lib().gpro_check_mover_collision(ctypes.byref(cfg), n, _p(qpos, _D), _p(csize, _D)))


def check_obstacle_collision(cfg, qpos, csize) -> bool:
    """gpro_check_obstacle_collision: any mover of `qpos` (n, 7) touching a static obstacle of `cfg`."""
    qpos = np.ascontiguousarray(qpos, dtype=np.float64)
    n = qpos.shape[0]
    cs = np.asarray(csize, dtype=np.float64)
    cs = cs.reshape(1, -1) if cs.size <= 2 else cs.reshape(n, -1)  # one size for all movers, or one row per mover
    csize = np.ascontiguousarray(np.broadcast_to(cs, (n, 2)))
    return bool(lib().gpro_check_obstacle_collision(ctypes.byref(cfg), n, _p(qpos, _D), _p(csize, _D)))


def compute_reward(cfg, achieved, desired, mover_collision=None, wall_collision=None):
    achieved = np.ascontiguousarray(achieved, dtype=np.float32)
    desired = np.ascontiguousarray(desired, dtype=np.float32)
    b = achieved.shape[0]
    mc = None if mover_collision is None else np.ascontiguousarray(mover_collision, dtype=np.uint8)
    wc = None if wall_collision is None else np.ascontiguousarray(wall_collision, dtype=np.uint8)
    r = np.zeros(b, dtype=np.float32)
    t = np.zeros(b, dtype=np.uint8)
    lib().gpro_compute_reward(ctypes.byref(cfg), b, _p(achieved, _F), _p(desired, _F), _p(mc, _U8), _p(wc, _U8), _p(r, _F), _p(t, _U8))
    return r, t.astype(bool)


def philox(c, k, rounds=10):
    out = np.zeros(4, dtype=np.uint32)
    lib().gpro_philox_r(*[ctypes.c_uint32(int(x)) for x in c], *[ctypes.c_uint32(int(x)) for x in k], ctypes.c_int(rounds), _p(out, ctypes.c_uint32))
    return out


def normals(seed, env_global, event, stream, lane0, count4):
    out = np.zeros(4 * count4, dtype=np.float32)
    lib().gpro_normals(ctypes.c_uint64(seed), ctypes.c_uint32(env_global), ctypes.c_uint32(event), ctypes.c_uint32(stream),
                       ctypes.c_uint32(lane0), count4, _p(out, _F))
    return out


def max_threads() -> int:
    return int(lib().gpro_max_threads())


# ---- batched environment (same SoA layout as the CUDA handle) --------------------------------------------------------
class OracleEnv:
    """Float64 CPU restatement of the batched env. ``cfg`` is the same ``GprConfig`` the CUDA library gets."""

    def __init__(self, cfg, nthreads: int = 1):
        self.cfg = cfg
        self.kind = int(cfg.env_kind)
        self.B = int(cfg.num_envs)
        self.N = int(cfg.num_movers)
        J = int(cfg.learn_jerk)
        if self.kind == 0:
            self.obs_dim, self.goal_dim, self.action_dim = 2 * self.N * (1 + J), 2 * self.N, 2 * self.N
        else:
            self.obs_dim, self.goal_dim, self.action_dim = 2 * (2 + J), 2, 2
        self.seed = int(cfg.seed)
        self.nthreads = nthreads
        B, N = self.B, self.N
        self.pos = np.zeros((B, N, 2))
        self.vel = np.zeros((B, N, 2))
        self.acc = np.zeros((B, N, 2))
        self.goal = np.zeros((B, self.goal_dim // 2, 2))
        self.elapsed_steps = np.zeros(B, dtype=np.int32)
        self.rng_counter = np.zeros(B, dtype=np.uint32)
        self.needs_reset = np.zeros(B, dtype=np.uint8)
        self.act = np.zeros((B, 2))
        self.mover_rot = np.zeros((B, 3))
        self.object_pos = np.zeros((B, 4))
        self.object_vel = np.zeros((B, 3))
        self.contact_warm = np.zeros((B, 13), dtype=np.float32)
        self._alloc_out()

    def _alloc_out(self):
        B = self.B
        self.observation = np.zeros((B, self.obs_dim))
        self.achieved_goal = np.zeros((B, self.goal_dim))
        self.desired_goal = np.zeros((B, self.goal_dim))
        self.reward = np.zeros(B)
        self.terminated = np.zeros(B, dtype=np.uint8)
        self.truncated = np.zeros(B, dtype=np.uint8)
        self.is_success = np.zeros(B, dtype=np.uint8)
        self.mover_collision = np.zeros(B, dtype=np.uint8)
        self.wall_collision = np.zeros(B, dtype=np.uint8)
        self.final_observation = np.zeros((B, self.obs_dim))
        self.final_achieved_goal = np.zeros((B, self.goal_dim))
        self.final_desired_goal = np.zeros((B, self.goal_dim))
        self.reset_failed = np.zeros(B, dtype=np.uint8)
        self.other_collision = np.zeros(B, dtype=np.uint8)

    def _state(self) -> OState:
        s = OState()
        for name, _ in OState._fields_:
            setattr(s, name, getattr(self, name).ctypes.data)
        return s

    def _out(self) -> OOutputs:
        o = OOutputs()
        for name, _ in OOutputs._fields_:
            setattr(o, name, getattr(self, name).ctypes.data)
        return o

    def reset(self, seed=None, mask=None, inject_start=None, inject_goal=None, inject_object=None):
        if seed is not None:
            self.seed = int(seed)
            self.rng_counter[:] = 0
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        st = None if inject_start is None else np.ascontiguousarray(inject_start, dtype=np.float64)
        gl = None if inject_goal is None else np.ascontiguousarray(inject_goal, dtype=np.float64)
        s, o = self._state(), self._out()
        if self.kind == 0:
            lib().gpro_planning_reset(ctypes.byref(self.cfg), ctypes.c_uint64(self.seed), ctypes.byref(s), _p(m, _U8),
                                      _p(st, _D), _p(gl, _D), ctypes.byref(o), self.nthreads)
        else:
            ob = None if inject_object is None else np.ascontiguousarray(inject_object, dtype=np.float64)
            lib().gpro_pushing_reset(ctypes.byref(self.cfg), ctypes.c_uint64(self.seed), ctypes.byref(s), _p(m, _U8),
                                     _p(st, _D), _p(gl, _D), _p(ob, _D), ctypes.byref(o), self.nthreads)
        return self.observation, self.achieved_goal, self.desired_goal

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float32)
        assert a.shape == (self.B, self.action_dim)
        s, o = self._state(), self._out()
        fn = lib().gpro_planning_step if self.kind == 0 else lib().gpro_pushing_step
        fn(ctypes.byref(self.cfg), ctypes.c_uint64(self.seed), ctypes.byref(s), _p(a, _F), ctypes.byref(o), self.nthreads)
        return self.observation, self.reward, self.terminated, self.truncated
