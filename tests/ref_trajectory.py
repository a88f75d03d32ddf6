"""TEST-ONLY: drive the reference's UNMODIFIED ``reset()`` / ``step()`` and record what it returns.

The reference env (imported from its read-only mount through ``ref_harness``; ``mujoco`` = the closed-form stand-in of
``mujoco_standin.py``) is run on K independent trajectories of T env-steps each, with gymnasium's ``TimeLimit(50)``
emulated by the driver (__init__.py:25-38) and a ``reset()`` after every finished episode.  Everything the env returns —
start / goal positions sampled by its own ``_reset_callback`` (PCG64), observations, rewards, terminated, info, and the
MjData state after every call — is recorded into arrays shaped [K, T, ...].  The oracle (and, on the GPU box, the CUDA
library) then replays the SAME actions with the recorded starts / goals injected, and must reproduce every recorded value
bit for bit.

Sensor noise.  NumPy's PCG64 + ziggurat stream cannot be reproduced by a counter-based generator, but the
reference only touches its noise generator through ``self.rng_noise.normal(loc, scale, size)``.  ``OracleNoise`` is a
drop-in object for that attribute which serves, for every call site of the step path in the order the reference makes
them, the variates of the oracle's Philox streams (include/gpr_rng.h).  With it the reference's unmodified code computes
``qpos + scale * z`` on the very same ``z`` the oracle uses, so noisy trajectories (the reference's default
``std_noise=1e-5``) are comparable bit for bit as well — which pins the DRAW SITES and their order (6+7+7 per mover and
cycle, 7+7 at reset, 7+6 for the observation; basic_envs.py:801-855, 1797-1805, 1888-1901, planning:430, 554-555).
"""

from __future__ import annotations

import numpy as np

# stream ids of include/gpr_rng.h
BLOCK_VEL_WALL, BLOCK_MOVER, BLOCK_WALL_QUAT, BLOCK_MOVER_QUAT = 0, 1, 2, 3
RNG_OBS, RNG_RESET_CHECK, RNG_RESET_WQUAT, RNG_RESET_MQUAT, RNG_OBJECT = 0x40000000, 0x40000001, 0x40000002, 0x40000003, 0x40000004


class OracleNoise:
    """Stand-in for ``env.rng_noise``: ``normal()`` returns the oracle's variates for the call site it is called from."""

    def __init__(self, seed: int, env_global: int, num_movers: int, box: bool, pushing: bool = False):
        import gpr_oracle

        self._normals = gpr_oracle.normals
        self.seed, self.env_global, self.N, self.box, self.pushing = int(seed), int(env_global), int(num_movers), bool(box), pushing
        self.event = 0
        self.mode = 'idle'
        self.k = 0  # counted calls since begin_*
        self.in_obs = False
        self.obs_k = 0

    def _n4(self, stream: int, lane: int) -> np.ndarray:
        return self._normals(self.seed, self.env_global, self.event, stream, lane, 1).astype(np.float64)

    def begin_reset(self, event: int) -> None:
        self.mode, self.event, self.k, self.in_obs, self.obs_k = 'reset', int(event), 0, False, 0

    def begin_step(self, event: int) -> None:
        self.mode, self.event, self.k, self.in_obs, self.obs_k = 'step', int(event), 0, False, 0

    def _obs_call(self, size: int) -> np.ndarray:
        # planning:554-555 / pushing:551-553: qpos of every mover (7 each), then qvel of every mover (6 each)
        N, k = self.N, self.obs_k
        self.obs_k += 1
        z = np.zeros(size)
        if k < N:
            assert size == 7, 'observation: expected the qpos draws first'
            o = self._n4(RNG_OBS, k)
            z[0], z[1] = o[0], o[1]
        elif k < 2 * N:
            assert size == 6, 'observation: expected the qvel draws second'
            o = self._n4(RNG_OBS, k - N)
            z[0], z[1] = o[2], o[3]
        else:
            raise AssertionError('more noise draws than the step path has sites')
        return z

    def normal(self, loc=0.0, scale=1.0, size=None):
        size = int(size if size is not None else 1)
        if scale == 0.0:  # add_noise=False (qacc; planning:433, 558): NumPy returns exact zeros as well
            return np.zeros(size) + loc
        if self.pushing and size == 2:  # pushing:565 object-position noise (its own, always-on sigma)
            o = self._n4(RNG_OBJECT, 0)
            return loc + scale * np.array([o[0], o[1]])
        N = self.N
        z = np.zeros(size)
        if self.in_obs:
            z = self._obs_call(size)
        elif self.mode == 'reset':
            # basic_envs.py:1799-1805: wall check qpos (7 per mover), mover check qpos (7 per mover), then the observation
            k = self.k
            self.k += 1
            if k < 2 * N:
                assert size == 7
                m = k % N
                r = self._n4(RNG_RESET_CHECK, m)
                if k < N:
                    z[0], z[1] = r[0], r[1]
                    if self.box:
                        z[3:7] = self._n4(RNG_RESET_WQUAT, m)
                else:
                    z[0], z[1] = r[2], r[3]
                    if self.box:
                        z[3:7] = self._n4(RNG_RESET_MQUAT, m)
            else:
                self.in_obs = True
                z = self._obs_call(size)
        elif self.mode == 'step':
            # per cycle: N x qvel (6) in _mujoco_step_callback, N x qpos (7) wall check, N x qpos (7) mover check
            per = 3 * N
            cyc, j = divmod(self.k, per)
            if j == 0 and size == 7:  # a cycle would start with a velocity draw: this is the observation
                self.in_obs = True
                z = self._obs_call(size)
            else:
                self.k += 1
                base = cyc * 4
                if j < N:
                    assert size == 6, 'cycle: expected a qvel draw'
                    v = self._n4(base + BLOCK_VEL_WALL, j)
                    z[0], z[1] = v[0], v[1]
                elif j < 2 * N:
                    assert size == 7
                    m = j - N
                    v = self._n4(base + BLOCK_VEL_WALL, m)
                    z[0], z[1] = v[2], v[3]
                    if self.box:
                        z[3:7] = self._n4(base + BLOCK_WALL_QUAT, m)
                else:
                    assert size == 7
                    m = j - 2 * N
                    v = self._n4(base + BLOCK_MOVER, m)
                    z[0], z[1] = v[0], v[1]
                    if self.box:
                        z[3:7] = self._n4(base + BLOCK_MOVER_QUAT, m)
        else:
            raise AssertionError('noise drawn outside reset()/step()')
        return loc + scale * z


def _state_of(env, N):
    """MjData -> (pos, vel, acc) [N,2] float64 (x, y of every mover's free joint)."""
    pos, vel, acc = np.zeros((N, 2)), np.zeros((N, 2)), np.zeros((N, 2))
    for m in range(N):
        pos[m] = env.get_mover_qpos(env.mover_names[m], add_noise=False)[:2]
        vel[m] = env.get_mover_qvel(env.mover_names[m], add_noise=False)[:2]
        acc[m] = env.get_mover_qacc(env.mover_names[m], add_noise=False)[:2]
    return pos, vel, acc


def _seek_action(obs, N, J, lim, arng):
    """A goal-seeking PD law (so that episodes also END IN SUCCESS, which random actions never do) with a little dither."""
    vel = obs['observation'][:2 * N]
    want = 12.0 * (obs['desired_goal'] - obs['achieved_goal']) - 6.0 * vel
    if J:
        want = (want - obs['observation'][2 * N:]) / 0.04
    return np.clip(want + arng.normal(0.0, 0.02 * lim, 2 * N), -1.2 * lim, 1.2 * lim)


def record_planning(kwargs: dict, K: int, T: int, seed: int, noise_seed: int | None, max_episode_steps: int = 50,
                    action_scale: float = 1.2, policy: str = 'random') -> dict[str, np.ndarray]:
    """K trajectories x T env-steps of the reference's BenchmarkPlanningEnv.  noise_seed None -> std_noise must be 0.
    policy: 'random' (uniform actions, 20 % beyond the action box so the clip of basic_envs.py:1871-1873 is exercised) or
    'seek' (PD law towards the goals: successes, +50 rewards)."""
    import ref_harness

    N = int(kwargs['num_movers'])
    J = int(bool(kwargs.get('learn_jerk', False)))
    box = kwargs.get('collision_params', {}).get('shape', 'circle') == 'box'
    obs_dim = 2 * N * (1 + J)
    rec = {
        'action': np.zeros((K, T, 2 * N), np.float32),
        'reset_before': np.zeros((K, T), np.uint8),
        'start': np.zeros((K, T, N, 2)), 'goal': np.zeros((K, T, N, 2)),
        'reset_obs': np.zeros((K, T, obs_dim)), 'reset_ag': np.zeros((K, T, 2 * N)), 'reset_dg': np.zeros((K, T, 2 * N)),
        'reset_info': np.zeros((K, T, 3), np.uint8),
        'obs': np.zeros((K, T, obs_dim)), 'ag': np.zeros((K, T, 2 * N)), 'dg': np.zeros((K, T, 2 * N)),
        'reward': np.zeros((K, T)), 'terminated': np.zeros((K, T), np.uint8), 'truncated': np.zeros((K, T), np.uint8),
        'info': np.zeros((K, T, 3), np.uint8),
        'pos': np.zeros((K, T, N, 2)), 'vel': np.zeros((K, T, N, 2)), 'acc': np.zeros((K, T, N, 2)),
    }
    keys = ('is_success', 'mover_collision', 'wall_collision')
    for k in range(K):
        env = ref_harness.make_planning_env(**kwargs)
        lim = env.j_max if env.learn_jerk else env.a_max
        arng = np.random.default_rng(100003 * seed + k)
        env.np_random = np.random.default_rng(7919 * seed + k)  # what gymnasium's Env.reset(seed=...) creates
        noise = None
        if noise_seed is not None:
            noise = OracleNoise(noise_seed, k, N, box)
            env.rng_noise = noise
        else:
            assert not np.any(env.std_noise), 'without OracleNoise the trajectory must be noise-free'
        event, elapsed, need_reset = 0, 0, True
        for t in range(T):
            if need_reset:
                if noise is not None:
                    noise.begin_reset(event)
                obs, info = env.reset()  # seed=None: the generators installed above stay in place
                event += 1
                elapsed = 0
                rec['reset_before'][k, t] = 1
                p, _, _ = _state_of(env, N)
                rec['start'][k, t], rec['goal'][k, t] = p, env.goals
                rec['reset_obs'][k, t], rec['reset_ag'][k, t], rec['reset_dg'][k, t] = obs['observation'], obs['achieved_goal'], obs['desired_goal']
                rec['reset_info'][k, t] = [bool(info[q]) for q in keys]
            if policy == 'seek':
                a32 = _seek_action(obs, N, J, lim, arng).astype(np.float32)
            else:
                a32 = arng.uniform(-action_scale * lim, action_scale * lim, 2 * N).astype(np.float32)
            rec['action'][k, t] = a32
            if noise is not None:
                noise.begin_step(event)
            obs, reward, term, trunc, info = env.step(a32.astype(np.float64))
            event += 1
            elapsed += 1
            assert trunc is False or trunc == False  # noqa: E712  (the bare env never truncates, planning:481-500)
            tl = elapsed >= max_episode_steps  # gymnasium TimeLimit
            rec['obs'][k, t], rec['ag'][k, t], rec['dg'][k, t] = obs['observation'], obs['achieved_goal'], obs['desired_goal']
            rec['reward'][k, t], rec['terminated'][k, t], rec['truncated'][k, t] = reward, bool(term), tl
            rec['info'][k, t] = [bool(info[q]) for q in keys]
            rec['pos'][k, t], rec['vel'][k, t], rec['acc'][k, t] = _state_of(env, N)
            need_reset = bool(term) or tl
        env.close()
    return rec


def replay_planning_on_oracle(rec: dict[str, np.ndarray], cfg) -> list[str]:
    """Replay a recorded reference trajectory on the oracle (B = K envs, auto-reset off, recorded starts / goals injected).
    Returns the list of mismatches (empty = bit-identical)."""
    import gpr_oracle

    K, T = rec['action'].shape[:2]
    ora = gpr_oracle.OracleEnv(cfg)
    ora.reset(seed=int(cfg.seed), mask=np.zeros(K, np.uint8))  # (only installs the seed / zeroes the event counters)
    bad: list[str] = []

    def cmp(name, t, got, want, rows=None):
        got, want = np.asarray(got), np.asarray(want)
        if rows is not None:
            got, want = got[rows], want[rows]
        if not np.array_equal(got, want):
            bad.append(f'{name} @ step {t}: max |diff| {np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))):.3e}')

    for t in range(T):
        m = rec['reset_before'][:, t].astype(bool)
        if m.any():
            ora.reset(mask=m.astype(np.uint8), inject_start=rec['start'][:, t], inject_goal=rec['goal'][:, t])
            cmp('reset observation', t, ora.observation, rec['reset_obs'][:, t], m)
            cmp('reset achieved_goal', t, ora.achieved_goal, rec['reset_ag'][:, t], m)
            cmp('reset desired_goal', t, ora.desired_goal, rec['reset_dg'][:, t], m)
            info = np.stack([ora.is_success, ora.mover_collision, ora.wall_collision], axis=1)
            cmp('reset info', t, info, rec['reset_info'][:, t], m)
        ora.step(rec['action'][:, t])
        cmp('observation', t, ora.observation, rec['obs'][:, t])
        cmp('achieved_goal', t, ora.achieved_goal, rec['ag'][:, t])
        cmp('desired_goal', t, ora.desired_goal, rec['dg'][:, t])
        cmp('reward', t, ora.reward, rec['reward'][:, t])
        cmp('terminated', t, ora.terminated, rec['terminated'][:, t])
        cmp('truncated', t, ora.truncated, rec['truncated'][:, t])
        cmp('info', t, np.stack([ora.is_success, ora.mover_collision, ora.wall_collision], axis=1), rec['info'][:, t])
        cmp('pos', t, ora.pos, rec['pos'][:, t])
        cmp('vel', t, ora.vel, rec['vel'][:, t])
        cmp('acc', t, ora.acc, rec['acc'][:, t])
    return bad


# ---- the committed trajectory fixtures (tests/golden/reference_trajectories.npz) ---------------------------------------
def _L(shape, holes=()):
    a = np.ones(shape)
    for h in holes:
        a[h] = 0
    return a


_BOX88 = {'shape': 'box', 'size': np.array([0.08, 0.08])}
# name -> (reference kwargs, K trajectories, T steps, sensor noise on?, policy)
PLANNING_CASES = {
    'n4_circle_acc_clean': (dict(layout_tiles=_L((3, 3)), num_movers=4), 3, 60, False, 'random'),      # BASELINE configs[1]
    'n4_circle_acc_noise': (dict(layout_tiles=_L((3, 3)), num_movers=4), 3, 60, True, 'random'),
    'n2_circle_acc_limit': (dict(layout_tiles=_L((3, 3)), num_movers=2, v_max=0.3), 2, 120, True, 'random'),  # TimeLimit
    'n1_circle_acc_seek': (dict(layout_tiles=_L((3, 3)), num_movers=1), 2, 60, True, 'seek'),
    'n2_circle_jerk_seek': (dict(layout_tiles=_L((4, 4)), num_movers=2, learn_jerk=True), 2, 60, True, 'seek'),
    'n2_circle_hole_offsets': (dict(layout_tiles=_L((4, 3), [(3, 1)]), num_movers=2,
                                    collision_params={'shape': 'circle', 'size': 0.1, 'offset': 0.01, 'offset_wall': 0.005}),
                               3, 60, True, 'seek'),
    'n3_circle_ring_jerk': (dict(layout_tiles=_L((4, 4), [(1, 1), (2, 2)]), num_movers=3, learn_jerk=True, j_max=400.0), 2, 50, True, 'random'),
    'n8_circle_jerk_clean': (dict(layout_tiles=_L((5, 5)), num_movers=8, learn_jerk=True), 2, 50, False, 'random'),
    'n8_box_jerk_noise': (dict(layout_tiles=_L((5, 5)), num_movers=8, learn_jerk=True, collision_params=_BOX88), 2, 50, True, 'random'),  # configs[3]
    'n4_box_acc_clean': (dict(layout_tiles=_L((4, 4)), num_movers=4, collision_params={'shape': 'box', 'size': np.array([0.09, 0.07]), 'offset': 0.005}),
                         2, 50, False, 'random'),
    'n2_box_jerk_seek_limits': (dict(layout_tiles=_L((4, 4)), num_movers=2, learn_jerk=True, num_cycles=42, v_max=0.5, a_max=4.0, j_max=150.0,
                                     mover_params={'mass': 0.628}, collision_params={'shape': 'box', 'size': np.array([0.09, 0.07]), 'offset': 0.005}),
                                2, 60, False, 'seek'),
    'n2_box_hole_noise': (dict(layout_tiles=_L((3, 3), [(0, 2)]), num_movers=2, collision_params=_BOX88), 3, 50, True, 'random'),
}
NOISE_SEED = 11


def planning_case_config(name: str, **over):
    """The ``gpr_config`` the oracle / the CUDA library replays case `name` with (B = K envs, auto-reset off)."""
    import gymnasium_planar_robotics_b200 as gpr

    kw, K, T, noise, policy = PLANNING_CASES[name]
    kw = dict(kw)
    if not noise:
        kw['std_noise'] = 0.0
    args = dict(num_envs=K, autoreset_mode='off', max_episode_steps=50, seed=NOISE_SEED)
    args.update(over)
    return gpr.planning_config(**args, **kw)


def record_planning_case(name: str, seed: int) -> dict[str, np.ndarray]:
    kw, K, T, noise, policy = PLANNING_CASES[name]
    kw = dict(kw)
    if not noise:
        kw['std_noise'] = 0.0
    return record_planning(kw, K, T, seed, NOISE_SEED if noise else None, policy=policy)


# ---- BenchmarkPushingEnv, contact-free part ---------------------------------------------------------------------------
# The stand-in cannot model contact: a trajectory is recorded until the mover's geom first reaches the object's
# (``ContactError``); `valid[k, t]` marks the steps that were recorded.  What this pins against the reference's own code:
# push:419-455 (control limiting on the real qacc, actuators, the impedance controller's zero wrench at rest), the noise
# draw sites incl. the always-on object noise (push:565), observation layout, reward / terminated / info (push:457-608),
# reset sampling order and the reset-time checks — everything of the pushing step path EXCEPT the contact dynamics.
def record_pushing(kwargs: dict, K: int, T: int, seed: int, noise_seed: int, max_episode_steps: int = 50,
                   action_scale: float = 1.2) -> dict[str, np.ndarray]:
    import ref_harness

    ref_harness.install()
    import mujoco  # the stand-in

    J = int(bool(kwargs.get('learn_jerk', False)))
    box = kwargs.get('collision_params', {}).get('shape', 'circle') == 'box'
    obs_dim = 2 * (2 + J)
    rec = {
        'action': np.zeros((K, T, 2), np.float32), 'valid': np.zeros((K, T), np.uint8), 'reset_before': np.zeros((K, T), np.uint8),
        'start': np.zeros((K, T, 1, 2)), 'object_start': np.zeros((K, T, 2)), 'goal': np.zeros((K, T, 1, 2)),
        'reset_obs': np.zeros((K, T, obs_dim)), 'reset_ag': np.zeros((K, T, 2)), 'reset_dg': np.zeros((K, T, 2)),
        'reset_info': np.zeros((K, T, 3), np.uint8),
        'obs': np.zeros((K, T, obs_dim)), 'ag': np.zeros((K, T, 2)), 'dg': np.zeros((K, T, 2)),
        'reward': np.zeros((K, T)), 'terminated': np.zeros((K, T), np.uint8), 'truncated': np.zeros((K, T), np.uint8),
        'info': np.zeros((K, T, 3), np.uint8),
        'pos': np.zeros((K, T, 1, 2)), 'vel': np.zeros((K, T, 1, 2)), 'acc': np.zeros((K, T, 1, 2)), 'object_pos': np.zeros((K, T, 2)),
    }
    keys = ('is_success', 'mover_collision', 'wall_collision')
    for k in range(K):
        env = ref_harness.make_pushing_env(**kwargs)
        lim = env.j_max if env.learn_jerk else env.a_max
        arng = np.random.default_rng(100003 * seed + k)
        env.np_random = np.random.default_rng(7919 * seed + k)
        noise = OracleNoise(noise_seed, k, 1, box, pushing=True)
        env.rng_noise = noise
        event, elapsed, need_reset = 0, 0, True
        try:
            for t in range(T):
                if need_reset:
                    noise.begin_reset(event)
                    obs, info = env.reset()
                    event += 1
                    elapsed = 0
                    rec['reset_before'][k, t] = 1
                    p, _, _ = _state_of(env, 1)
                    rec['start'][k, t], rec['object_start'][k, t], rec['goal'][k, t, 0] = p, env.object_xy_start_pos, env.object_xy_goal_pos
                    rec['reset_obs'][k, t], rec['reset_ag'][k, t], rec['reset_dg'][k, t] = obs['observation'], obs['achieved_goal'], obs['desired_goal']
                    rec['reset_info'][k, t] = [bool(info[q]) for q in keys]
                a32 = arng.uniform(-action_scale * lim, action_scale * lim, 2).astype(np.float32)
                noise.begin_step(event)
                obs, reward, term, trunc, info = env.step(a32.astype(np.float64))  # raises ContactError on contact
                rec['action'][k, t] = a32
                rec['valid'][k, t] = 1
                event += 1
                elapsed += 1
                tl = elapsed >= max_episode_steps
                rec['obs'][k, t], rec['ag'][k, t], rec['dg'][k, t] = obs['observation'], obs['achieved_goal'], obs['desired_goal']
                rec['reward'][k, t], rec['terminated'][k, t], rec['truncated'][k, t] = reward, bool(term), tl
                rec['info'][k, t] = [bool(info[q]) for q in keys]
                rec['pos'][k, t], rec['vel'][k, t], rec['acc'][k, t] = _state_of(env, 1)
                rec['object_pos'][k, t] = env.data.qpos[env.model.joint('object_joint').qposadr[0]:][:2]
                need_reset = bool(term) or tl
        except mujoco.ContactError:
            # the reset that may have been recorded for step t stays in the file but is not replayed (valid = 0)
            rec['reset_before'][k, t] = 0
        env.close()
    return rec


def replay_pushing_on_oracle(rec: dict[str, np.ndarray], cfg) -> list[str]:
    import gpr_oracle

    K, T = rec['action'].shape[:2]
    ora = gpr_oracle.OracleEnv(cfg)
    ora.reset(seed=int(cfg.seed), mask=np.zeros(K, np.uint8))
    bad: list[str] = []

    def cmp(name, t, got, want, rows):
        got, want = np.asarray(got)[rows], np.asarray(want)[rows]
        if not np.array_equal(got, want):
            bad.append(f'{name} @ step {t}: max |diff| {np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))):.3e}')

    for t in range(T):
        v = rec['valid'][:, t].astype(bool)
        m = rec['reset_before'][:, t].astype(bool) & v
        if m.any():
            ora.reset(mask=m.astype(np.uint8), inject_start=rec['start'][:, t], inject_goal=rec['goal'][:, t], inject_object=rec['object_start'][:, t])
            cmp('reset observation', t, ora.observation, rec['reset_obs'][:, t], m)
            cmp('reset achieved_goal', t, ora.achieved_goal, rec['reset_ag'][:, t], m)
            cmp('reset desired_goal', t, ora.desired_goal, rec['reset_dg'][:, t], m)
            cmp('reset info', t, np.stack([ora.is_success, ora.mover_collision, ora.wall_collision], axis=1), rec['reset_info'][:, t], m)
        if not v.any():
            break
        ora.step(rec['action'][:, t])
        cmp('observation', t, ora.observation, rec['obs'][:, t], v)
        cmp('achieved_goal', t, ora.achieved_goal, rec['ag'][:, t], v)
        cmp('desired_goal', t, ora.desired_goal, rec['dg'][:, t], v)
        cmp('reward', t, ora.reward, rec['reward'][:, t], v)
        cmp('terminated', t, ora.terminated, rec['terminated'][:, t], v)
        cmp('truncated', t, ora.truncated, rec['truncated'][:, t], v)
        cmp('info', t, np.stack([ora.is_success, ora.mover_collision, ora.wall_collision], axis=1), rec['info'][:, t], v)
        cmp('pos', t, ora.pos, rec['pos'][:, t], v)
        cmp('vel', t, ora.vel, rec['vel'][:, t], v)
        cmp('acc', t, ora.acc, rec['acc'][:, t], v)
        cmp('object_pos', t, ora.object_pos[:, :2], rec['object_pos'][:, t], v)
    return bad


PUSHING_CASES = {
    'push_acc_noise': (dict(), 4, 60),                                              # BASELINE configs[2] kwargs
    'push_jerk_noise': (dict(learn_jerk=True), 4, 60),
    'push_acc_slow': (dict(v_max=0.2, a_max=2.0), 3, 120),                          # long episodes: TimeLimit
    'push_box_jerk': (dict(learn_jerk=True, collision_params={'shape': 'box', 'size': np.array([0.09, 0.09]), 'offset': 0.005, 'offset_wall': 0.002}), 3, 60),
}


def pushing_case_config(name: str, **over):
    import gymnasium_planar_robotics_b200 as gpr

    kw, K, T = PUSHING_CASES[name]
    args = dict(num_envs=K, autoreset_mode='off', max_episode_steps=50, seed=NOISE_SEED)
    args.update(over)
    return gpr.pushing_config(**args, **kw)


def record_pushing_case(name: str, seed: int) -> dict[str, np.ndarray]:
    kw, K, T = PUSHING_CASES[name]
    return record_pushing(dict(kw), K, T, seed, NOISE_SEED)
