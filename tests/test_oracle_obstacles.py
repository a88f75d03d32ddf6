"""Static obstacles: the typed form of the reference's extension point
``_check_for_other_collisions_callback`` (basic_envs.py:1976-1986, called at :1807 and :1903).  The reference ships no
obstacle implementation (the hook returns False in both benchmark envs), so the rules are this project's and are fixed
here on the CPU oracle: mover-mover rules between movers and fixed shapes (include/gpr.h, gpr_config.num_obstacles).
The GPU parity tests then demand bit-identical results from the CUDA path."""

import numpy as np
import pytest

import gpr_oracle as oracle
import gymnasium_planar_robotics_b200 as gpr


def qpos_of(xy):
    xy = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
    q = np.zeros((xy.shape[0], 7))
    q[:, :2] = xy
    q[:, 3] = 1.0
    return q


def test_circle_rule_is_inclusive_distance_against_radius_sum():
    obst = [[0.5, 0.5, 0.0625], [0.8, 0.2, 0.02]]
    cfg, _ = gpr.planning_config(num_envs=1, layout_tiles=np.ones((4, 4)), num_movers=2, obstacles=obst)
    assert cfg.num_obstacles == 2
    r = 0.125  # (binary fractions: 0.6875 - 0.5 == 0.125 + 0.0625 exactly)
    assert oracle.check_obstacle_collision(cfg, qpos_of([[0.6875, 0.5], [0.2, 0.8]]), r)                     # distance == radius sum: inclusive
    assert not oracle.check_obstacle_collision(cfg, qpos_of([[np.nextafter(0.6875, 1), 0.5], [0.2, 0.8]]), r)
    assert oracle.check_obstacle_collision(cfg, qpos_of([[0.2, 0.2], [0.8 - 0.09, 0.2 + 0.09]]), r)          # second mover, second obstacle
    rng = np.random.default_rng(0)
    for _ in range(2000):
        xy = rng.uniform(0.1, 0.9, (2, 2))
        want = any(np.sqrt((xy[m, 0] - o[0]) ** 2 + (xy[m, 1] - o[1]) ** 2) <= r + o[2] for m in range(2) for o in obst)
        assert oracle.check_obstacle_collision(cfg, qpos_of(xy), r) == want


def test_box_rule_is_edge_intersection_or_centre_inside():
    box = {'shape': 'box', 'size': np.array([0.08, 0.06])}
    cfg, _ = gpr.planning_config(num_envs=1, layout_tiles=np.ones((4, 4)), num_movers=1, collision_params=box, obstacles=[[0.5, 0.5, 0.2, 0.1]])
    s = np.array([0.08, 0.06])
    assert oracle.check_obstacle_collision(cfg, qpos_of([[0.5 + 0.2 + 0.08, 0.5]]), s)                      # edges touch (1e-7 tolerance, geom:37-45)
    assert not oracle.check_obstacle_collision(cfg, qpos_of([[0.5 + 0.2 + 0.0801, 0.5]]), s)
    assert oracle.check_obstacle_collision(cfg, qpos_of([[0.5, 0.5]]), s)                                    # fully inside: no edges cross
    assert oracle.check_obstacle_collision(cfg, qpos_of([[0.45, 0.52]]), s)
    assert not oracle.check_obstacle_collision(cfg, qpos_of([[0.5, 0.5 + 0.1 + 0.0601]]), s)
    q = qpos_of([[0.5 + 0.2 + 0.085, 0.5]])
    assert not oracle.check_obstacle_collision(cfg, q, s)
    q[0, 3:] = [np.cos(0.35), 0, 0, np.sin(0.35)]                                                            # yaw 0.7 rad swings a corner into it
    assert oracle.check_obstacle_collision(cfg, q, s)


def test_obstacle_kwarg_validation():
    kw = dict(num_envs=1, layout_tiles=np.ones((3, 3)), num_movers=1)
    with pytest.raises(ValueError):
        gpr.planning_config(obstacles=[[0.3, 0.3, 0.05, 0.05]], **kw)                                        # circle shape wants (K, 3)
    with pytest.raises(ValueError):
        gpr.planning_config(obstacles=[[0.3, 0.3, 0.0]], **kw)
    with pytest.raises(ValueError):
        gpr.planning_config(obstacles=np.tile([[0.3, 0.3, 0.01]], (9, 1)), **kw)
    with pytest.raises(TypeError):
        gpr.pushing_config(num_envs=1, obstacles=[[0.3, 0.3, 0.05]])
    cfg, d = gpr.planning_config(**kw)
    assert cfg.num_obstacles == 0 and d['obstacles'].shape[0] == 0


@pytest.mark.parametrize('shape', ['circle', 'box'])
def test_driving_into_an_obstacle_ends_the_step_like_any_collision(shape):
    """basic:1903-1904: the hook's verdict breaks the cycle loop; here it also counts as a collision in reward and
    termination and is reported as other_collision (mover / wall flags stay clear)."""
    cp = {'shape': 'circle', 'size': 0.11} if shape == 'circle' else {'shape': 'box', 'size': np.array([0.1, 0.1])}
    obst = [[1.2, 0.6, 0.1]] if shape == 'circle' else [[1.2, 0.6, 0.1, 0.3]]
    cfg, _ = gpr.planning_config(num_envs=2, layout_tiles=np.ones((8, 5)), num_movers=1, collision_params=cp, obstacles=obst,
                                 std_noise=0.0, autoreset_mode='off', max_episode_steps=0)
    env = oracle.OracleEnv(cfg)
    start = np.array([[[0.6, 0.6]], [[0.6, 0.9 if shape == 'circle' else 1.05]]])  # env 1 passes beside the obstacle
    env.reset(seed=0, inject_start=start, inject_goal=np.full((2, 1, 2), [1.7, 1.0]))
    assert not env.other_collision.any()
    act = np.array([[10.0, 0.0], [10.0, 0.0]], dtype=np.float32)
    hit_step = None
    for t in range(12):
        env.step(act)
        if env.other_collision[0]:
            hit_step = t
            break
        assert env.reward[0] == -1.0 and not env.terminated[0]
    assert hit_step is not None
    gap = 1.2 - 0.1 - (0.11 if shape == 'circle' else 0.1)                       # centre x at first contact
    assert env.pos[0, 0, 0] >= gap and env.pos[0, 0, 0] < gap + 2 * 0.002 * 1.01  # stopped within a cycle of contact
    assert env.reward[0] == -50.0 and env.terminated[0] and not env.is_success[0]
    assert not env.mover_collision[0] and not env.wall_collision[0]
    assert not env.other_collision[1] and env.reward[1] == -1.0 and not env.terminated[1]
    assert env.pos[1, 0, 0] > env.pos[0, 0, 0]                                   # env 1 kept integrating all 40 cycles


def test_sampled_starts_and_goals_clear_the_obstacles_by_the_safety_offset():
    obst = np.array([[0.36, 0.36, 0.08], [0.2, 0.5, 0.03]])
    cp = {'shape': 'circle', 'size': 0.08, 'offset': 0.01}
    cfg, d = gpr.planning_config(num_envs=3000, layout_tiles=np.ones((3, 3)), num_movers=2, collision_params=cp, obstacles=obst, seed=4)
    env = oracle.OracleEnv(cfg, nthreads=oracle.max_threads())
    env.reset(seed=4)
    assert not env.reset_failed.any() and not env.other_collision.any()
    for arr in (env.pos, env.goal):
        for o in obst:
            dist = np.hypot(arr[..., 0] - o[0], arr[..., 1] - o[1])
            assert dist.min() > 0.08 + 0.01 + o[2]
    # and the same seed without obstacles samples different (closer) positions: the rejection really was at work
    cfg0, _ = gpr.planning_config(num_envs=3000, layout_tiles=np.ones((3, 3)), num_movers=2, collision_params=cp, seed=4)
    env0 = oracle.OracleEnv(cfg0, nthreads=oracle.max_threads())
    env0.reset(seed=4)
    assert np.hypot(env0.pos[..., 0] - 0.36, env0.pos[..., 1] - 0.36).min() < 0.17


def test_no_obstacles_is_the_reference_path():
    """num_obstacles = 0 changes nothing: same trajectories as a config that never heard of obstacles."""
    kw = dict(num_envs=64, layout_tiles=np.ones((3, 3)), num_movers=3, seed=9)
    a = oracle.OracleEnv(gpr.planning_config(**kw)[0])
    b = oracle.OracleEnv(gpr.planning_config(obstacles=np.zeros((0, 3)), **kw)[0])
    a.reset(seed=9)
    b.reset(seed=9)
    rng = np.random.default_rng(1)
    for _ in range(10):
        act = rng.uniform(-10, 10, (64, 6)).astype(np.float32)
        a.step(act)
        b.step(act)
        assert np.array_equal(a.pos, b.pos) and np.array_equal(a.reward, b.reward) and not b.other_collision.any()


# ---- typed extra bodies with a prescribed velocity (gpr_config.obstacle_vel) --------------------------
def test_extra_body_kwarg_is_the_obstacle_list_with_velocities():
    cfg, d = gpr.planning_config(num_envs=1, layout_tiles=np.ones((4, 4)), num_movers=1, obstacles=[[0.5, 0.5, 0.05]],
                                 extra_bodies=[{'shape': 'circle', 'pos': (0.2, 0.7), 'size': 0.03, 'vel': (0.25, -0.125)}, {'pos': (0.8, 0.8), 'size': 0.02}])
    assert cfg.num_obstacles == 3 and d['obstacles'].shape == (3, 5)
    assert [cfg.obstacle_vel[0][0], cfg.obstacle_vel[0][1]] == [0.0, 0.0] and [cfg.obstacle_vel[1][0], cfg.obstacle_vel[1][1]] == [0.25, -0.125]
    assert cfg.obstacle_xy[2][0] == 0.8 and cfg.obstacle_size[2][0] == 0.02
    with pytest.raises(NotImplementedError):
        gpr.planning_config(num_envs=1, layout_tiles=np.ones((4, 4)), num_movers=1, extra_bodies=[{'shape': 'box', 'pos': (0.2, 0.7), 'size': (0.1, 0.1)}])
    with pytest.raises(ValueError):
        gpr.planning_config(num_envs=1, layout_tiles=np.ones((4, 4)), num_movers=1, obstacles=np.zeros((9, 3)) + 0.1)


@pytest.mark.parametrize('shape', ['circle', 'box'])
def test_moving_body_reaches_a_resting_mover_at_the_predicted_cycle(shape):
    """A body moving at constant velocity towards a mover at rest: the env must flag `other_collision` in the control cycle
    in which  p0 + v * ((s * num_cycles + c + 1) * dt)  first touches the mover's collision shape, and freeze there."""
    if shape == 'circle':
        cp, body, reach = {'shape': 'circle', 'size': 0.08}, {'pos': (0.2, 0.48), 'size': 0.04, 'vel': (0.5, 0.0)}, 0.08 + 0.04
    else:
        cp, reach = {'shape': 'box', 'size': np.array([0.07, 0.05])}, 0.07 + 0.03
        body = {'shape': 'box', 'pos': (0.2, 0.48), 'size': (0.03, 0.2), 'vel': (0.5, 0.0)}
    cfg, _ = gpr.planning_config(num_envs=1, layout_tiles=np.ones((4, 4)), num_movers=1, std_noise=0.0, collision_params=cp,
                                 extra_bodies=[body], autoreset_mode='off', max_episode_steps=0)
    env = oracle.OracleEnv(cfg)
    env.reset(seed=0, inject_start=np.array([[[0.48, 0.48]]]), inject_goal=np.array([[[0.7, 0.7]]]))
    assert not env.other_collision[0]
    # gap at t = 0: 0.28 - reach; the body closes 0.5 mm per cycle.  The box test has the reference's 1e-7 edge tolerance.
    gap = 0.48 - 0.2 - reach
    hit_cycle = next(n for n in range(1, 4000) if 0.2 + 0.5 * (n * 0.001) + reach >= 0.48 - (1e-7 if shape == 'box' else 0.0))
    assert abs(hit_cycle - gap / 0.0005) <= 1
    steps = 0
    while not env.other_collision[0]:
        env.step(np.zeros((1, 2), np.float32))
        steps += 1
        assert steps <= 20
    assert steps == (hit_cycle + 39) // 40  # the env-step that contains that cycle
    assert env.terminated[0] and env.reward[0] == -50.0 and not env.mover_collision[0] and not env.wall_collision[0]
    assert np.array_equal(env.pos[0, 0], [0.48, 0.48])  # the mover never moved


def test_custom_env_without_xml_conveyor_example():
    """docs/make_own_env.rst builds custom envs by adding bodies to the MuJoCo XML and checking them in
    `_check_for_other_collisions_callback`.  The same env, typed: a fixed robot base and two pallets on a conveyor crossing the
    tiles; movers must reach their goals without touching either.  (CUDA == oracle for this env: tests/test_gpu_parity.py)"""
    kw = custom_conveyor_env_kwargs()
    cfg, _ = gpr.planning_config(num_envs=64, seed=3, **kw)
    env = oracle.OracleEnv(cfg)
    env.reset(seed=3)
    rng = np.random.default_rng(0)
    hits = 0
    for _ in range(40):
        env.step(rng.uniform(-6, 6, (64, 4)).astype(np.float32))
        hits += int(env.other_collision.sum())
        assert ((env.reward == -50) == (env.other_collision | env.mover_collision | env.wall_collision).astype(bool)).all()
    assert hits > 0


def custom_conveyor_env_kwargs():
    return dict(layout_tiles=np.ones((4, 4)), num_movers=2, collision_params={'shape': 'circle', 'size': 0.09, 'offset': 0.005},
                extra_bodies=[{'shape': 'circle', 'pos': (0.48, 0.48), 'size': 0.07},                        # robot base, fixed
                              {'shape': 'circle', 'pos': (0.10, 0.80), 'size': 0.04, 'vel': (0.30, 0.0)},    # pallets on a conveyor
                              {'shape': 'circle', 'pos': (0.86, 0.16), 'size': 0.04, 'vel': (-0.30, 0.0)}])
