"""Host-side checks of the pushing env in the oracle.

What the reference pins (tests/test_benchmark_pushing_env.py): with the object far away the mover follows the planning
closed form.  Everything involving contact lives inside MuJoCo in the reference and is not pinned by any reference test
(SURVEY.md §0.5): for that part these tests check the planar physics SPECIFICATION (include/gpr_push_physics.h) through
physical properties — rest stays rest, momentum balance, Coulomb friction, no deep penetration, no energy creation.
"""

import ctypes

import numpy as np
import pytest

import gpr_oracle as oracle
import gymnasium_planar_robotics_b200 as gpr

_D = ctypes.c_double


def _env(num_envs=1, **kw):
    cfg, d = gpr.pushing_config(num_envs=num_envs, **kw)
    return oracle.OracleEnv(cfg), cfg, d


_WARM = {}


def _substep(cfg, mover, obj, ux, uy):
    """one 1 ms substep on explicit bodies [x, y, cos, sin, vx, vy, w]; returns (contacts, mover qacc).  The warm-start state
    of the contact solve is carried from call to call per `obj` array, as the env carries it from substep to substep."""
    qacc = np.zeros(2)
    lib = oracle.lib()
    warm = _WARM.setdefault(id(obj), (obj, np.zeros(13, dtype=np.float32)))[1]
    nc = lib.gpro_push_substep(ctypes.byref(cfg), mover.ctypes.data_as(ctypes.POINTER(_D)), obj.ctypes.data_as(ctypes.POINTER(_D)),
                               _D(ux), _D(uy), qacc.ctypes.data_as(ctypes.POINTER(_D)), warm.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return nc, qacc


@pytest.mark.parametrize('mass,jerk,num_cycles,tx,ty', [(0.628, 100, 1, True, True), (1.237, -100, 42, True, True), (0.628, 100, 42, True, False),
                                                        (0.628, -100, 1, False, True)])
def test_jerk_closed_form_of_reference_tests(mass, jerk, num_cycles, tx, ty):
    """tests/test_benchmark_pushing_env.py:10-100 without MuJoCo: mover at (.48,.48), object far away, tiny limits."""
    env, cfg, _ = _env(1, mover_params={'mass': mass}, std_noise=0.0, num_cycles=num_cycles, v_max=0.01, a_max=0.2, j_max=150.0,
                       learn_jerk=True, autoreset_mode='off', max_episode_steps=0)
    start = np.array([[[0.48, 0.48]]])
    env.reset(seed=0, inject_start=start, inject_goal=np.array([[[0.3, 0.3]]]), inject_object=np.array([[0.12, 0.36]]))
    j = np.array([jerk / 2 if tx and ty else (jerk if tx else 0), jerk / 2 if tx and ty else (jerk if ty else 0)], dtype=np.float64)
    p, v, a = start[0, 0].copy(), np.zeros(2), np.zeros(2)
    for step in range(100):
        for _ in range(num_cycles):
            na, _ = oracle.ensure_max_dyn_val(a, 0.2, j, 0.001)
            v, a = oracle.ensure_max_dyn_val(v, 0.01, na, 0.001)
            p = 0.001 * v + p
        env.step(j[None].astype(np.float32))
        assert np.linalg.norm(env.vel[0, 0]) <= 0.01 * (1 + 1e-9) and np.linalg.norm(env.acc[0, 0]) <= 0.2 * (1 + 1e-9)
    assert np.allclose(env.pos[0, 0], p) and np.allclose(env.vel[0, 0], v) and np.allclose(env.acc[0, 0], a)
    # the object never moved, the mover never rotated
    assert np.array_equal(env.object_pos[0], [0.12, 0.36, 1.0, 0.0]) and np.array_equal(env.object_vel[0], [0, 0, 0])
    assert np.array_equal(env.mover_rot[0], [1.0, 0.0, 0.0])
    assert np.array_equal(env.observation[0], np.concatenate([env.pos[0, 0], env.vel[0, 0], env.acc[0, 0]]))  # push:529-560


@pytest.mark.parametrize('acc,num_cycles', [(10, 1), (-10, 42)])
def test_acc_closed_form_of_reference_tests(acc, num_cycles):
    """tests/test_benchmark_pushing_env.py:103-191 (acceleration actuator)."""
    env, cfg, _ = _env(1, std_noise=0.0, num_cycles=num_cycles, v_max=0.01, a_max=0.2, learn_jerk=False, autoreset_mode='off', max_episode_steps=0)
    env.reset(seed=0, inject_start=np.array([[[0.48, 0.48]]]), inject_goal=np.array([[[0.3, 0.3]]]), inject_object=np.array([[0.12, 0.36]]))
    u = np.clip(np.array([acc / 2, acc / 2]), -0.2, 0.2)
    p, v = np.array([0.48, 0.48]), np.zeros(2)
    for step in range(100):
        for _ in range(num_cycles):
            v, _ = oracle.ensure_max_dyn_val(v, 0.01, u, 0.001)
            p = 0.001 * v + p
        env.step(np.array([[acc / 2, acc / 2]], dtype=np.float32))
    assert np.allclose(env.pos[0, 0], p) and np.allclose(env.vel[0, 0], v)
    assert env.observation.shape == (1, 4)


def test_reset_sampling_follows_the_reference_ranges():
    """push:250-288, 386-413: mover ~U[.11,.55]^2, object and goal ~U[.22,.44]^2, object farther than min_mo_dist."""
    env, cfg, d = _env(4096, std_noise=0.0, seed=3)
    env.reset(seed=3)
    assert np.isclose(d['min_mo_dist'], np.linalg.norm([0.035 + 0.155 / 2] * 2))
    mp, op, g = env.pos[:, 0], env.object_pos[:, :2], env.goal[:, 0]
    assert (mp >= 0.11).all() and (mp <= 0.55).all() and (op >= 0.22).all() and (op <= 0.44).all() and (g >= 0.22).all() and (g <= 0.44).all()
    ok = ~env.reset_failed.astype(bool)
    assert (np.linalg.norm(op - mp, axis=1)[ok] > d['min_mo_dist']).all()
    # pushing:392-407 cannot succeed when the mover sits within a few mm of the layout centre (the whole object box is
    # then inside min_mo_dist; the reference would loop forever): only such envs may report a failed reset
    corners = np.array([[0.22, 0.22], [0.22, 0.44], [0.44, 0.22], [0.44, 0.44]])
    far = np.linalg.norm(mp[:, None, :] - corners[None], axis=2).max(axis=1)
    assert (far[~ok] < d['min_mo_dist'] + 4e-3).all() and (~ok).sum() <= 8  # (a sliver of < 4 mm: < 1e-4 per attempt)
    assert mp.std(axis=0).min() > 0.1 and op.std(axis=0).min() > 0.05  # actually spread over the boxes
    assert not env.wall_collision.any()
    # achieved goal = object position (+ N(0,1e-5), push:565), desired = goal
    assert np.abs(env.achieved_goal - op).max() < 1e-4 and np.abs(env.achieved_goal - op).max() > 0
    assert np.array_equal(env.desired_goal, g)


def test_reward_termination_and_wall_collision():
    """push:457-527, 578-608: -50 and terminated on a wall collision, 0 within the threshold (success does NOT terminate),
    -1 otherwise; break at the colliding cycle."""
    env, cfg, _ = _env(3, std_noise=0.0, autoreset_mode='off', max_episode_steps=0)
    start = np.array([[[0.3, 0.3]], [[0.3, 0.3]], [[0.1105, 0.3]]])  # env 2: 0.5 mm from the wall limit x - c > 0
    goal = np.array([[[0.4, 0.4]], [[0.25, 0.42]], [[0.4, 0.4]]])
    obj = np.array([[0.4 + 0.03, 0.4], [0.42, 0.25], [0.4, 0.4]])
    env.reset(seed=0, inject_start=start, inject_goal=goal, inject_object=obj)
    env.step(np.array([[0, 0], [0, 0], [-10, 0]], dtype=np.float32))
    assert list(env.reward) == [0.0, -1.0, -50.0]
    assert list(env.terminated) == [0, 0, 1] and list(env.is_success) == [1, 0, 0] and list(env.wall_collision) == [0, 0, 1]
    assert not env.mover_collision.any()
    # x(t) = x0 - 10 * dt^2 * k(k+1)/2 crosses 0.11 during cycle 10 (0.55 mm): the state is frozen there
    assert np.isclose(env.pos[2, 0, 0], 0.1105 - 10 * 1e-6 * 55) and np.isclose(env.vel[2, 0, 0], -0.1)
    r, t = oracle.compute_reward(cfg, env.achieved_goal, env.desired_goal, None, env.wall_collision)
    assert np.array_equal(r, env.reward.astype(np.float32)) and np.array_equal(t, env.terminated.astype(bool))


def test_object_at_rest_stays_exactly_at_rest():
    env, cfg, _ = _env(1, std_noise=0.0)
    m = np.array([0.2, 0.2, 1.0, 0.0, 0.0, 0.0, 0.0])
    o = np.array([0.4, 0.4, 1.0, 0.0, 0.0, 0.0, 0.0])
    for _ in range(100):
        nc, q = _substep(cfg, m, o, 1.0, -2.0)
        assert nc == 0 and np.array_equal(q, [1.0, -2.0])
    assert np.array_equal(o, [0.4, 0.4, 1.0, 0.0, 0.0, 0.0, 0.0])


def test_sliding_object_is_stopped_by_coulomb_friction():
    """A free-sliding object decelerates with about mu*g (soft friction: a little less at low speed) and comes to rest
    without reversing."""
    env, cfg, _ = _env(1, std_noise=0.0)
    m = np.array([0.15, 0.15, 1.0, 0.0, 0.0, 0.0, 0.0])
    o = np.array([0.4, 0.4, 1.0, 0.0, 0.3, 0.1, 0.0])
    v0 = np.hypot(0.3, 0.1)
    speeds = []
    for _ in range(200):
        _substep(cfg, m, o, 0.0, 0.0)
        speeds.append(np.hypot(o[4], o[5]))
        assert o[4] >= -1e-9 and o[5] >= -1e-9  # never reverses
    speeds = np.array(speeds)
    assert (np.diff(speeds) <= 1e-12).all()  # monotone
    dec = (v0 - speeds[9]) / 0.010
    assert 0.8 * 9.81 <= dec <= 1.05 * 9.81  # mu = 1: Coulomb deceleration mu*g (+ joint damping 0.01/0.01 = 1/s * v)
    assert speeds[-1] < 1e-3
    assert abs(o[6]) < 1e-3 and abs(o[3]) < 1e-4  # straight slide: (practically) no spin — Gauss-Seidel order leaves ~1e-6


def test_push_momentum_balance_and_penetration():
    """Mover pushes the object head-on: the pair's momentum changes only through the actuator and ground friction, the
    boxes never interpenetrate by more than the soft-contact depth (~ v * 0.02 s / e at a 0.25 m/s impact), the object ends up riding on the mover's face."""
    env, cfg, _ = _env(1, std_noise=0.0)
    mm, mo = cfg.mover_mass, cfg.object_mass
    m = np.array([0.2, 0.36, 1.0, 0.0, 0.25, 0.0, 0.0])
    o = np.array([0.2 + 0.0775 + 0.035 + 0.002, 0.36, 1.0, 0.0, 0.0, 0.0, 0.0])  # 2 mm gap
    touched = False
    for k in range(400):
        p_before = mm * m[4] + mo * o[4]
        nc, q = _substep(cfg, m, o, 0.0, 0.0)
        p_after = mm * m[4] + mo * o[4]
        gap = (o[0] - 0.035) - (m[0] + 0.0775)
        assert gap > -2.5e-3, f'penetration {-gap} m at substep {k}'  # soft contact, time constant 0.02 s: ~v*tc/e at impact
        # external impulse on the pair = ground friction (|F| <= mu m g) + object joint damping (0.01 * v)
        assert abs(p_after - p_before) <= (cfg.friction * mo * cfg.gravity + 0.01 * abs(o[4]) + 1e-9) * 1e-3 * 1.001
        if nc:
            touched = True
            assert q[0] <= 1e-9  # contact can only decelerate the pushing mover
    assert touched
    assert abs(o[4] - m[4]) < 5e-3 and o[4] > 0.2  # moving together
    assert abs(o[1] - 0.36) < 2e-3 and abs(m[1] - 0.36) < 1e-4  # head-on push stays (nearly) straight
    assert abs(m[3]) < 1e-3  # yaw impedance keeps the mover aligned
    ke = 0.5 * mm * (m[4] ** 2 + m[5] ** 2) + 0.5 * mo * (o[4] ** 2 + o[5] ** 2)
    assert ke <= 0.5 * mm * 0.25 ** 2 + 1e-12  # no energy created


def test_off_centre_push_rotates_the_object():
    env, cfg, _ = _env(1, std_noise=0.0)
    m = np.array([0.2, 0.36, 1.0, 0.0, 0.3, 0.0, 0.0])
    o = np.array([0.2 + 0.0775 + 0.035 + 0.001, 0.36 + 0.0775 + 0.02, 1.0, 0.0, 0.0, 0.0, 0.0])  # only 1.5 cm of face overlap
    for _ in range(300):
        _substep(cfg, m, o, 0.0, 0.0)
    assert o[6] != 0.0 or abs(o[3]) > 1e-3  # the object was turned
    assert abs(np.hypot(o[2], o[3]) - 1.0) < 1e-12 and abs(np.hypot(m[2], m[3]) - 1.0) < 1e-12  # orientations stay unit


@pytest.mark.parametrize('mode', ['same_step', 'next_step'])
def test_autoreset_modes_and_timelimit(mode):
    env, cfg, _ = _env(64, std_noise=1e-5, autoreset_mode=mode, max_episode_steps=7, seed=11)
    env.reset(seed=11)
    rng = np.random.default_rng(0)
    ends = 0
    for t in range(30):
        env.step(rng.uniform(-10, 10, (64, 2)).astype(np.float32))
        done = (env.terminated | env.truncated).astype(bool)
        ends += int(done.sum())
        assert (env.elapsed_steps <= 7).all()
        if mode == 'same_step':
            assert (env.elapsed_steps[done] == 0).all()
            assert (env.vel[done] == 0).all() and (env.object_vel[done] == 0).all()
    assert ends >= 64 * 3


def test_oracle_sharding_invariance():
    """results depend on the GLOBAL env index only."""
    kw = dict(std_noise=1e-5, seed=5)
    whole, _, _ = _env(32, **kw)
    halves = [_env(16, env_index_base=16 * r, **kw)[0] for r in range(2)]
    rng = np.random.default_rng(1)
    whole.reset(seed=5)
    for h in halves:
        h.reset(seed=5)
    for _ in range(12):
        a = rng.uniform(-10, 10, (32, 2)).astype(np.float32)
        whole.step(a)
        for r, h in enumerate(halves):
            h.step(a[16 * r:16 * (r + 1)])
        assert np.array_equal(whole.observation, np.concatenate([h.observation for h in halves]))
        assert np.array_equal(whole.object_pos, np.concatenate([h.object_pos for h in halves]))
        assert np.array_equal(whole.reward, np.concatenate([h.reward for h in halves]))
