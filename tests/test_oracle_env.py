"""Host-side checks of the oracle's batched env (and through it of the env semantics the CUDA path must reproduce):
closed-form integrator from the reference's own tests, reward/termination bookkeeping, auto-reset modes, sharding."""

import numpy as np
import pytest

import gpr_oracle as oracle
import gymnasium_planar_robotics_b200 as gpr


def _env(num_envs=4, **kw):
    kw.setdefault('layout_tiles', np.ones((9, 9)))
    kw.setdefault('num_movers', 2)
    cfg, d = gpr.planning_config(num_envs=num_envs, **kw)
    return oracle.OracleEnv(cfg), cfg


@pytest.mark.parametrize('jerk,num_cycles,tx,ty', [(100, 1, True, True), (100, 1, True, False), (-100, 42, True, True), (100, 42, False, True)])
def test_jerk_closed_form_of_reference_tests(jerk, num_cycles, tx, ty):
    """tests/test_benchmark_planning_env.py:10-121 re-expressed without MuJoCo: the env's control+integration equals the
    closed form  a<-clip(a+dt*j); (v,a)<-ensure(v,v_max,a); p<-p+dt*v  to np.allclose."""
    env, cfg = _env(1, std_noise=0.0, num_cycles=num_cycles, v_max=0.01, a_max=0.2, j_max=150.0, learn_jerk=True, autoreset_mode='off', max_episode_steps=0)
    start = np.array([[[0.96, 0.96], [1.2, 1.2]]])
    env.reset(seed=0, inject_start=start, inject_goal=start)
    j = np.array([jerk / 2 if tx and ty else (jerk if tx else 0), jerk / 2 if tx and ty else (jerk if ty else 0)])
    p, v, a = start[0].copy(), np.zeros((2, 2)), np.zeros((2, 2))
    dt = 0.001
    for step in range(100):
        for m in range(2):
            for _ in range(num_cycles):
                na, _ = oracle.ensure_max_dyn_val(a[m], 0.2, j, dt)
                v[m], a[m] = oracle.ensure_max_dyn_val(v[m], 0.01, na, dt)
                p[m] = dt * v[m] + p[m]
        env.step(np.tile(j, 2)[None].astype(np.float32))
        assert np.linalg.norm(env.vel[0], axis=1).max() <= 0.01 * (1 + 1e-9)
        assert np.linalg.norm(env.acc[0], axis=1).max() <= 0.2 * (1 + 1e-9)
    assert np.allclose(env.pos[0], p) and np.allclose(env.vel[0], v) and np.allclose(env.acc[0], a)


@pytest.mark.parametrize('acc,num_cycles', [(10, 1), (-10, 42)])
def test_acc_closed_form_of_reference_tests(acc, num_cycles):
    """tests/test_benchmark_planning_env.py:124-231 (acceleration actuator)."""
    env, cfg = _env(1, std_noise=0.0, num_cycles=num_cycles, v_max=0.01, a_max=0.2, learn_jerk=False, autoreset_mode='off', max_episode_steps=0)
    start = np.array([[[0.96, 0.96], [1.2, 1.2]]])
    env.reset(seed=0, inject_start=start, inject_goal=start)
    u = np.clip(np.array([acc / 2, acc / 2]), -0.2, 0.2)  # the env clips the action to the Box first (basic:1871-1873)
    p, v = start[0].copy(), np.zeros((2, 2))
    for step in range(100):
        for m in range(2):
            for _ in range(num_cycles):
                v[m], _ = oracle.ensure_max_dyn_val(v[m], 0.01, u, 0.001)
                p[m] = 0.001 * v[m] + p[m]
        env.step(np.tile([acc / 2, acc / 2], 2)[None].astype(np.float32))
    assert np.allclose(env.pos[0], p) and np.allclose(env.vel[0], v)


def test_break_on_collision_freezes_state_at_the_colliding_cycle():
    """basic:1904: the loop stops at the first colliding cycle; flags, reward -50, terminated."""
    env, cfg = _env(1, layout_tiles=np.ones((3, 3)), num_movers=2, std_noise=0.0, autoreset_mode='off')
    start = np.array([[[0.3, 0.3], [0.3 + 0.2205, 0.3]]])  # 0.5 mm outside the 0.22 collision distance
    env.reset(seed=0, inject_start=start, inject_goal=start)
    assert not env.mover_collision[0]
    env.step(np.array([[10.0, 0, -10.0, 0]], dtype=np.float32))  # approach at +-10 m/s^2
    assert env.mover_collision[0] and not env.wall_collision[0] and env.terminated[0] and env.reward[0] == -50
    gap = env.pos[0, 1, 0] - env.pos[0, 0, 0]
    assert gap <= 0.22 and gap > 0.22 - 2e-4  # stopped right at the threshold (7 cycles), not 40 cycles later
    k = round(env.vel[0, 0, 0] / 0.01)  # cycles executed: v = k*dt*a
    assert 1 <= k < 40


def test_rewards_and_success():
    env, cfg = _env(3, layout_tiles=np.ones((3, 3)), num_movers=2, std_noise=0.0, autoreset_mode='off')
    st = np.array([[[0.2, 0.2], [0.5, 0.5]]] * 3)
    gl = st.copy()
    gl[1, 1] = [0.2, 0.5]     # second mover far from its goal
    gl[2] = [[0.5, 0.2], [0.2, 0.5]]
    env.reset(seed=0, inject_start=st, inject_goal=gl)
    assert env.is_success.tolist() == [1, 0, 0]
    env.step(np.zeros((3, 4), dtype=np.float32))
    assert env.reward.tolist() == [50.0, -1.0, -2.0] and env.terminated.tolist() == [1, 0, 0]  # plan:526-528


@pytest.mark.parametrize('mode', ['same_step', 'next_step'])
def test_autoreset_modes_and_timelimit(mode):
    env, cfg = _env(64, layout_tiles=np.ones((3, 3)), num_movers=2, std_noise=1e-5, autoreset_mode=mode, max_episode_steps=5, seed=3)
    env.reset(seed=3)
    assert oracle.lib().gpro_assert_trips(1) == 0
    rng = np.random.default_rng(0)
    truncs = 0
    for t in range(30):
        before = env.elapsed_steps.copy()
        pending = env.needs_reset.copy()
        env.step(rng.uniform(-3, 3, (64, 4)).astype(np.float32))
        done = (env.terminated | env.truncated).astype(bool)
        truncs += int(env.truncated.sum())
        if mode == 'same_step':
            assert (env.elapsed_steps[done] == 0).all() and (env.elapsed_steps[~done] == before[~done] + 1).all()
            assert (np.abs(env.vel[done]) == 0).all()
        else:
            assert (env.elapsed_steps[pending == 1] == 0).all() and (env.reward[pending == 1] == 0).all()
            assert not done[pending == 1].any() and (env.needs_reset == done).all()
    assert truncs > 0 and env.reset_failed.sum() == 0 and oracle.lib().gpro_assert_trips(1) == 0


def test_oracle_sharding_invariance():
    """RNG keyed by the global env index: two shards == one whole (what the multi-GPU path relies on)."""
    kw = dict(layout_tiles=np.ones((3, 3)), num_movers=3, std_noise=1e-5, seed=11, autoreset_mode='same_step', max_episode_steps=6)
    whole = oracle.OracleEnv(gpr.planning_config(num_envs=100, **kw)[0])
    parts = [oracle.OracleEnv(gpr.planning_config(num_envs=50, env_index_base=50 * r, **kw)[0]) for r in range(2)]
    whole.reset(seed=11)
    for p in parts:
        p.reset(seed=11)
    rng = np.random.default_rng(1)
    for t in range(20):
        a = rng.uniform(-10, 10, (100, 6)).astype(np.float32)
        whole.step(a)
        for r, p in enumerate(parts):
            p.step(a[50 * r:50 * r + 50])
        assert np.array_equal(whole.observation, np.concatenate([p.observation for p in parts]))
        assert np.array_equal(whole.desired_goal, np.concatenate([p.desired_goal for p in parts]))
        assert np.array_equal(whole.reward, np.concatenate([p.reward for p in parts]))


def test_baseline_config0_shape_single_env_1000_steps():
    """BASELINE.json configs[0]: BenchmarkPlanningEnv-v0, 2 movers, ONE env, random actions, 1000 steps with a reset whenever
    an episode ends (the reference's own CPU-runnable case) — run on the oracle with reference-default kwargs.
    Invariants of the reference's step contract: rewards in {-50, +50, -2, -1}, terminated <=> |reward| == 50, TimeLimit(50),
    a collision flag with every -50, state inside the layout while no wall collision is reported."""
    cfg, _ = gpr.planning_config(num_envs=1, layout_tiles=np.ones((3, 3)), num_movers=2, autoreset_mode='same_step', seed=0)
    env = oracle.OracleEnv(cfg)
    env.reset(seed=0)
    rng = np.random.default_rng(0)
    episodes, lengths, length = 0, [], 0
    for _ in range(1000):
        act = rng.uniform(-10.0, 10.0, (1, 4)).astype(np.float32)
        env.step(act)
        r, term, trunc = float(env.reward[0]), bool(env.terminated[0]), bool(env.truncated[0])
        length += 1
        assert r in (-50.0, 50.0, -2.0, -1.0)
        assert term == (abs(r) == 50.0)
        assert (r == -50.0) == bool(env.mover_collision[0] or env.wall_collision[0])
        if term or trunc:
            assert length <= 50 and (trunc == (length == 50) or term)
            episodes += 1
            lengths.append(length)
            length = 0
            assert env.elapsed_steps[0] == 0 and not np.any(env.vel[0])  # SAME_STEP: the new episode has already begun
            assert np.all(env.pos[0] >= 0.11) and np.all(env.pos[0] <= 0.55)  # spawn box incl. the plan:264-267 quirk
        else:
            assert np.all(env.pos[0] > 0.0) and np.all(env.pos[0] < 0.72)
    assert episodes > 50 and max(lengths) <= 50
