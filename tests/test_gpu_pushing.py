"""GPU parity tests of BenchmarkPushingEnv-v0 (``-m gpu``): the CUDA path through the C ABI against the CPU oracle.

Everything around mj_step (control limiting, wall check, observation, reward, termination, reset sampling, auto-reset)
restates the reference and must agree exactly.  The contact substep is this project's planar specification
(include/gpr_push_physics.h; MuJoCo parity unpinned): the kernels are built with -fmad=false and the
oracle with -ffp-contract=off, so even the contact trajectories must be bit-identical between CPU and GPU.
"""

import numpy as np
import pytest
import torch

import gpr_oracle as oracle
import gymnasium_planar_robotics_b200 as gpr

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'
STATE_KEYS = ('pos', 'vel', 'acc', 'goal', 'act', 'mover_rot', 'object_pos', 'object_vel', 'contact_warm')


def make_pair(num_envs, **kw):
    env = gpr.BenchmarkPushingVecEnv(num_envs, device=DEV, **kw)
    cfg, _ = gpr.pushing_config(num_envs=num_envs, **kw)
    return env, oracle.OracleEnv(cfg, nthreads=oracle.max_threads())


def assert_outputs_equal(env, ora, what):
    b = env.core.buf
    torch.cuda.synchronize()
    for k in ('terminated', 'truncated', 'is_success', 'mover_collision', 'wall_collision'):
        got, ref = b[k].cpu().numpy(), getattr(ora, k)
        assert np.array_equal(got, ref), f'{what}: {k} differs in {np.count_nonzero(got != ref)} envs'
    assert np.array_equal(b['reward'].cpu().numpy(), ora.reward.astype(np.float32)), f'{what}: reward'
    for k in ('observation', 'achieved_goal', 'desired_goal'):
        got, ref = b[k].cpu().numpy(), getattr(ora, k).astype(np.float32)
        assert np.array_equal(got, ref), f'{what}: {k} max|d|={np.abs(got - ref).max()}'


def assert_state_equal(env, ora, what):
    st = env.get_state()
    torch.cuda.synchronize()
    for k in STATE_KEYS:
        got, ref = st[k].cpu().numpy(), getattr(ora, k)
        assert np.array_equal(got, ref), f'{what}: state {k} differs in {np.count_nonzero(got != ref)} values, max|d|={np.abs(got - ref).max()}'
    assert np.array_equal(st['elapsed_steps'].cpu().numpy(), ora.elapsed_steps), f'{what}: elapsed'
    assert np.array_equal(st['rng_counter'].cpu().numpy().view(np.uint32), ora.rng_counter), f'{what}: rng counter'


def contact_starts(rng, B):
    """mover next to the object so that random pushes produce contact within a few env-steps"""
    obj = rng.uniform(0.25, 0.41, (B, 2))
    ang = rng.uniform(0, 2 * np.pi, B)
    dist = rng.uniform(0.115, 0.16, B)
    mover = obj + dist[:, None] * np.stack([np.cos(ang), np.sin(ang)], axis=1)
    goal = rng.uniform(0.22, 0.44, (B, 1, 2))
    return mover.reshape(B, 1, 2), obj, goal


@pytest.mark.parametrize('learn_jerk', [False, True])
def test_pushing_contact_trajectories_bit_identical(learn_jerk):
    """BASELINE configs[2] shape (1 mover + box object), no noise, starts injected next to the object, actions that push
    towards it: contacts occur in most envs and every state variable must match the oracle bit for bit."""
    B = 2048
    rng = np.random.default_rng(31 + learn_jerk)
    env, ora = make_pair(B, std_noise=0.0, learn_jerk=learn_jerk, autoreset_mode='off', max_episode_steps=0)
    mover, obj, goal = contact_starts(rng, B)
    env.reset(seed=1, options={'mover_start_xy_pos': mover, 'object_start_xy_pos': obj, 'object_goal_xy_pos': goal})
    ora.reset(seed=1, inject_start=mover, inject_goal=goal, inject_object=obj)
    assert_state_equal(env, ora, 'reset')
    lim = 100.0 if learn_jerk else 10.0
    for t in range(25):
        towards = (ora.object_pos[:, :2] - ora.pos[:, 0])
        towards /= np.linalg.norm(towards, axis=1, keepdims=True)
        a = (lim * (0.7 * towards + rng.uniform(-0.6, 0.6, (B, 2)))).astype(np.float32)
        env.step(torch.as_tensor(a, device=DEV))
        ora.step(a)
        assert_outputs_equal(env, ora, f'step {t}')
    assert_state_equal(env, ora, 'final')
    moved = np.linalg.norm(ora.object_pos[:, :2] - obj, axis=1) > 1e-3
    assert moved.mean() > 0.5, 'the pushes were supposed to move most objects'
    assert (ora.object_pos[:, 3] != 0).mean() > 0.3  # and to rotate many of them
    env.close()


@pytest.mark.parametrize('shape', ['circle', 'box'])
def test_pushing_with_sensor_noise_and_box_shape(shape):
    """Reference-default noise (1e-5) and a large one; box collision shape uses the mover's yaw (rotated by contact)."""
    for sigma in (1e-5, np.array([1e-3, 2e-3, 0.0])):
        B = 1024
        rng = np.random.default_rng(41)
        cp = {'shape': 'circle', 'size': 0.11} if shape == 'circle' else {'shape': 'box', 'size': np.array([0.08, 0.08]), 'offset_wall': 0.002}
        env, ora = make_pair(B, std_noise=sigma, learn_jerk=True, collision_params=cp, autoreset_mode='off', seed=8)
        mover, obj, goal = contact_starts(rng, B)
        env.reset(seed=8, options={'mover_start_xy_pos': mover, 'object_start_xy_pos': obj, 'object_goal_xy_pos': goal})
        ora.reset(seed=8, inject_start=mover, inject_goal=goal, inject_object=obj)
        assert_outputs_equal_reset(env, ora)
        for t in range(15):
            a = rng.uniform(-130, 130, (B, 2)).astype(np.float32)
            env.step(torch.as_tensor(a, device=DEV))
            ora.step(a)
            assert_outputs_equal(env, ora, f'{shape} step {t}')
        assert_state_equal(env, ora, shape)
        assert ora.wall_collision.any()
        env.close()


def assert_outputs_equal_reset(env, ora):
    torch.cuda.synchronize()
    b = env.core.buf
    for k in ('observation', 'achieved_goal', 'desired_goal'):
        assert np.array_equal(b[k].cpu().numpy(), getattr(ora, k).astype(np.float32)), f'reset {k}'
    for k in ('is_success', 'mover_collision', 'wall_collision'):
        assert np.array_equal(b[k].cpu().numpy(), getattr(ora, k)), f'reset {k}'


@pytest.mark.parametrize('mode', ['same_step', 'next_step'])
def test_pushing_autoreset_sampled_on_device(mode):
    """Full episodes: on-device sampling of mover / object / goal (pushing:373-417), TimeLimit, both auto-reset modes."""
    B = 2048
    rng = np.random.default_rng(51)
    env, ora = make_pair(B, std_noise=1e-5, autoreset_mode=mode, seed=77, max_episode_steps=9)
    env.reset(seed=77)
    ora.reset(seed=77)
    assert_outputs_equal_reset(env, ora)
    assert_state_equal(env, ora, 'reset')
    ends = 0
    for t in range(40):
        a = rng.uniform(-10, 10, (B, 2)).astype(np.float32)
        env.step(torch.as_tensor(a, device=DEV))
        ora.step(a)
        assert_outputs_equal(env, ora, f'{mode} step {t}')
        done = (ora.terminated | ora.truncated).astype(bool)
        ends += int(done.sum())
        if mode == 'same_step' and done.any():
            for k in ('observation', 'achieved_goal', 'desired_goal'):
                got = env.core.buf['final_' + k].cpu().numpy()[done]
                assert np.array_equal(got, getattr(ora, 'final_' + k).astype(np.float32)[done]), f'final_{k}'
    assert ends > B * 3
    assert_state_equal(env, ora, mode)
    env.close()


def test_pushing_reset_failure_is_reported_not_hung():
    """The reference's object-placement loop (pushing:392-407) never ends for a mover at the layout centre; here the loop
    is capped, the failure counted, and CPU and GPU agree on the kept draw."""
    B = 64
    env, ora = make_pair(B, std_noise=0.0, autoreset_mode='off', max_reset_attempts=300, seed=2)
    mover = np.tile(np.array([[[0.33, 0.33]]]), (B, 1, 1))
    mover[B // 2:] = [[0.2, 0.2]]
    env.reset(seed=2, options={'mover_start_xy_pos': mover})
    ora.reset(seed=2, inject_start=mover)
    assert_state_equal(env, ora, 'centre')
    assert env.core.reset_failures() == int(ora.reset_failed.sum()) == B // 2
    env.close()


def test_pushing_host_api_and_single_env_form():
    # host buffers (rows written over PCIe by BOTH kernels of the population split: the free kernel writes filler rows for the
    # envs it parks, the contact kernel the real ones afterwards) against the device-pointer call, every output, long enough
    # for a fifth of the envs to be in the contact regime and for episodes to end (TimeLimit 50, walls, successes)
    rng = np.random.default_rng(61)
    B = 4099
    e1 = gpr.BenchmarkPushingVecEnv(B, device=DEV, seed=4)
    e2 = gpr.BenchmarkPushingVecEnv(B, device=DEV, seed=4)
    e1.reset(seed=4)
    e2.reset(seed=4)
    finished = moving = 0
    for t in range(60):
        a = rng.uniform(-10, 10, (B, 2)).astype(np.float32)
        o1, r1, t1, tr1, i1 = e1.step(torch.as_tensor(a, device=DEV))
        o2, r2, t2, tr2, i2 = e2.step_host(a)
        torch.cuda.synchronize()
        for k in ('observation', 'achieved_goal', 'desired_goal'):
            assert np.array_equal(o1[k].cpu().numpy(), o2[k]), (t, k)
        assert np.array_equal(r1.cpu().numpy(), r2) and np.array_equal(t1.cpu().numpy(), t2) and np.array_equal(tr1.cpu().numpy(), tr2), t
        for k in ('is_success', 'mover_collision', 'wall_collision'):
            assert np.array_equal(i1[k].cpu().numpy(), i2[k]), (t, k)
        d = t2 | tr2
        fo = e2.final_obs_dense(i2)
        for k in ('observation', 'achieved_goal', 'desired_goal'):
            assert np.array_equal(i1['final_obs'][k].cpu().numpy()[d], fo[k][d]), (t, 'final ' + k)
        finished += int(d.sum())
        moving += int((e1.state_dict()['object_vel'] != 0).any(dim=1).sum()) if t == 40 else 0
    assert finished > B and moving > B // 10, (finished, moving)
    e1.close()
    e2.close()
    env = gpr.BenchmarkPushingEnv(render_mode=None, std_noise=0.0, learn_jerk=True)
    obs, info = env.reset(seed=0, options={'mover_start_xy_pos': np.array([[0.48, 0.48]]), 'object_start_xy_pos': np.array([0.25, 0.3])})
    assert obs['observation'].shape == (6,) and obs['achieved_goal'].shape == (2,) and obs['observation'].dtype == np.float64
    o, r, term, trunc, info = env.step(np.array([50.0, 0.0]))
    # one env-step of 40 cycles from rest with jerk 50: a = 40*dt*j = 2, v = dt^2*j*820, (push:529-560 layout: pos, vel, acc)
    assert np.allclose(o['observation'][4:], [2.0, 0.0]) and np.allclose(o['observation'][2:4], [50 * 1e-6 * 820, 0.0])
    assert r in (-1.0, 0.0) and term is False and set(info) == {'is_success', 'mover_collision', 'wall_collision'}
    env.close()


def test_pushing_full_size_properties():
    """BASELINE configs[2] at full size (65,536 envs): size-independent properties over 60 steps with auto-reset."""
    B = 65536
    env = gpr.BenchmarkPushingVecEnv(B, device=DEV, seed=99)
    env2 = gpr.BenchmarkPushingVecEnv(B, device=DEV, seed=99)
    g = torch.Generator(device=DEV).manual_seed(0)
    env.reset(seed=99)
    env2.reset(seed=99)
    finished = 0
    for t in range(60):
        a = (torch.rand((B, 2), device=DEV, generator=g) * 2 - 1) * 10
        obs, r, term, trunc, info = env.step(a)
        o2, r2, term2, _, _ = env2.step(a)
        assert torch.equal(r, r2) and torch.equal(term, term2) and torch.equal(obs['achieved_goal'], o2['achieved_goal'])  # deterministic
        assert torch.equal(term, info['wall_collision']) and not info['mover_collision'].any()  # push:475, 592
        assert torch.equal(r == -50, info['wall_collision']) and torch.equal((r == 0), info['is_success'])
        assert ((r == -50) | (r == -1) | (r == 0)).all()
        st = env.get_state()
        assert (st['vel'].norm(dim=-1) <= 2.0 + 1e-3).all()
        assert ((st['mover_rot'][:, 0] ** 2 + st['mover_rot'][:, 1] ** 2 - 1).abs() < 1e-12).all()
        assert ((st['object_pos'][:, 2] ** 2 + st['object_pos'][:, 3] ** 2 - 1).abs() < 1e-12).all()
        assert torch.isfinite(st['object_pos']).all() and torch.isfinite(st['object_vel']).all()
        assert (st['object_vel'][:, :2].norm(dim=-1) < 4.0).all()  # nothing is ever launched faster than ~2 v_max
        done = term | trunc
        finished += int(done.sum())
        if done.any():
            assert (st['vel'][done] == 0).all() and (st['object_vel'][done] == 0).all() and (st['elapsed_steps'][done] == 0).all()
            op = st['object_pos'][done][:, :2]
            assert (op >= 0.22).all() and (op <= 0.44).all()
    stats = env.episode_stats(reset=True)
    assert stats['episodes'] == finished and finished > B // 4
    # only the centre-of-layout case, where pushing:392-407 has no (or a vanishing) feasible region: p < 1e-3 per reset
    assert env.core.reset_failures() <= 2e-3 * (finished + B)
    env.close()
    env2.close()
