"""Generate the committed golden fixtures by running the UNMODIFIED reference (read-only mount /root/reference).

Run once in the build container (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (all small, committed):
    reference_test_vectors.json : the reference's own test vectors — 100 `qpos_is_valid` cases
                                  (tests/test_basic_env.py:10-1633), 17+17 segment / rectangle pairs
                                  (tests/test_geometry_2D_utils.py:10-164) — with the reference's expected results AND the
                                  results the reference code returned when executed here (they agree).
    reference_random_vectors.npz: seeded random inputs with the outputs of the reference functions
                                  `ensure_max_dyn_val`, `qpos_is_valid`, `check_mover_collision`,
                                  `check_rectangles_intersect`, `get_2D_rect_vertices`, `compute_reward`,
                                  `compute_terminated` (planning and pushing).
"""

from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_harness  # noqa: E402

ref_harness.install()
from gymnasium_planar_robotics.envs.basic_envs import BasicPlanarRoboticsEnv  # noqa: E402
from gymnasium_planar_robotics.utils import geometry_2D_utils as geom  # noqa: E402


def _load_test_module(name):
    path = os.path.join(ref_harness.REFERENCE_ROOT, 'tests', name)
    spec = importlib.util.spec_from_file_location(name[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _params(fn):
    return fn.pytestmark[0].args[1]


def reference_test_vectors():
    out = {'wall': [], 'segments': [], 'rectangles': []}
    tb = _load_test_module('test_basic_env.py')
    for layout, qpos, cparams, safety, expected, c_size in _params(tb.test_mover_position_is_valid_check):
        env = BasicPlanarRoboticsEnv(layout_tiles=layout, num_movers=1, collision_params=cparams)
        cs = c_size if c_size is not None else env.c_size
        got = env.qpos_is_valid(qpos, c_size=cs, add_safety_offset=safety)
        assert (got == expected).all()
        # basic_envs.py:487, evaluated by the reference's own expression
        total = cs + env.c_size_offset_wall + int(safety) * env.c_size_offset
        total_arr = env.get_c_size_arr(c_size=total, num_reps=qpos.shape[0])
        out['wall'].append(
            {
                'layout': np.asarray(layout).tolist(),
                'qpos': np.asarray(qpos, dtype=np.float64).tolist(),
                'shape': env.c_shape,
                'csize_total': np.asarray(total_arr, dtype=np.float64).tolist(),
                'expected': np.asarray(expected).astype(int).tolist(),
                'reference_output': np.asarray(got).astype(int).tolist(),
            }
        )
    tg = _load_test_module('test_geometry_2D_utils.py')
    for p1, p2, q1, q2, expected in _params(tg.test_line_segments_intersect_check):
        got = geom.check_line_segments_intersect(p1=p1, p2=p2, q1=q1, q2=q2)
        assert (got == expected).all()
        out['segments'].append(
            {
                'p1': np.asarray(p1, float).tolist(),
                'p2': np.asarray(p2, float).tolist(),
                'q1': np.asarray(q1, float).tolist(),
                'q2': np.asarray(q2, float).tolist(),
                'expected': np.asarray(expected).astype(int).tolist(),
            }
        )
    for r1, r2, s1, s2, expected in _params(tg.test_rectangles_intersect_check):
        got = geom.check_rectangles_intersect(qpos_r1=r1, qpos_r2=r2, size_r1=s1, size_r2=s2)
        assert (got == expected).all()
        out['rectangles'].append(
            {
                'qpos_r1': np.asarray(r1, float).tolist(),
                'qpos_r2': np.asarray(r2, float).tolist(),
                'size_r1': np.asarray(s1, float).tolist(),
                'size_r2': np.asarray(s2, float).tolist(),
                'expected': np.asarray(expected).astype(int).tolist(),
            }
        )
    return out


LAYOUTS = {
    'full3x3': np.ones((3, 3)),
    'full5x5': np.ones((5, 5)),
    'L': np.array([[1, 1], [1, 0]]),
    'hole3x3': np.array([[1, 1, 1], [1, 0, 1], [1, 1, 1]]),
    'ragged': np.array([[1, 1, 0, 1], [1, 1, 1, 1], [0, 1, 1, 0], [1, 1, 1, 1], [1, 0, 1, 1]]),
    'strip': np.ones((1, 4)),
}


def _yaw_quat(yaw):
    q = np.zeros((yaw.shape[0], 4))
    q[:, 0] = np.cos(yaw / 2)
    q[:, 3] = np.sin(yaw / 2)
    return q


def random_qpos(rng, env, n, yaw=True, snap=True):
    """Positions over the whole grid (plus a small margin kept inside, the reference asserts outside), a share of them
    snapped onto tile borders / thresholds to exercise the inclusive/strict comparisons."""
    W = env.num_tiles_x * 2 * env.tile_size[0]
    H = env.num_tiles_y * 2 * env.tile_size[1]
    q = np.zeros((n, 7))
    q[:, 0] = rng.uniform(0, W, n)
    q[:, 1] = rng.uniform(0, H, n)
    if snap:
        k = n // 4
        gx = np.round(q[:k, 0] / 0.01) * 0.01
        gy = np.round(q[:k, 1] / 0.01) * 0.01
        q[:k, 0] = np.clip(gx, 0, W)
        q[:k, 1] = np.clip(gy, 0, H)
    q[:, 2] = 0.003
    if yaw:
        ang = rng.uniform(-np.pi, np.pi, n)
        ang[: n // 5] = rng.choice([0, np.pi / 2, np.pi / 4, -np.pi / 2, np.pi], n // 5)
        q[:, 3:] = _yaw_quat(ang)
    else:
        q[:, 3] = 1
    return q


def reference_random_vectors(seed=20240607):
    rng = np.random.default_rng(seed)
    out = {}
    penv = ref_harness.make_planning_env(layout_tiles=np.ones((3, 3)), num_movers=2)

    # --- ensure_max_dyn_val (planning:610-645)
    n = 4000
    cur = rng.normal(0, 1.5, (n, 2))
    der = rng.uniform(-150, 150, (n, 2))
    maxv = rng.choice([2.0, 10.0, 0.01, 0.2], n)
    cur[: n // 8] *= 1e-3
    nv = np.zeros((n, 2))
    nd = np.zeros((n, 2))
    for i in range(n):
        a, b = penv.ensure_max_dyn_val(cur[i], float(maxv[i]), der[i])
        nv[i], nd[i] = a[0], b[0]
    out.update(emdv_cur=cur, emdv_der=der, emdv_max=maxv, emdv_next=nv, emdv_next_der=nd)

    # --- qpos_is_valid (basic_envs.py:459-788)
    wall_meta = []
    for lname, layout in LAYOUTS.items():
        for shape, size, off_w, off in [
            ('circle', 0.11, 0.0, 0.0),
            ('circle', 0.09, 0.001, 0.005),
            ('box', np.array([0.08, 0.08]), 0.0, 0.0),
            ('box', np.array([0.1, 0.06]), 0.002, 0.01),
        ]:
            env = BasicPlanarRoboticsEnv(
                layout_tiles=layout, num_movers=1, collision_params={'shape': shape, 'size': size, 'offset': off, 'offset_wall': off_w}
            )
            for safety in (False, True):
                nq = 300
                qpos = random_qpos(rng, env, nq, yaw=shape == 'box')
                got = env.qpos_is_valid(qpos, c_size=env.c_size, add_safety_offset=safety)
                total = env.c_size + env.c_size_offset_wall + int(safety) * env.c_size_offset
                tot = env.get_c_size_arr(c_size=total, num_reps=nq)
                key = f'wall_{len(wall_meta)}'
                out[key + '_qpos'] = qpos
                out[key + '_csize'] = np.asarray(tot, dtype=np.float64)
                out[key + '_valid'] = np.asarray(got, dtype=np.int32)
                wall_meta.append({'key': key, 'layout': lname, 'shape': shape})
    out['wall_meta'] = np.array(json.dumps(wall_meta))

    # --- check_mover_collision (basic_envs.py:355-424)
    mov_meta = []
    for shape, size, off in [('circle', 0.11, 0.0), ('circle', 0.1, 0.01), ('box', np.array([0.08, 0.08]), 0.0), ('box', np.array([0.1, 0.06]), 0.005)]:
        for N in (2, 3, 4, 8):
            env = BasicPlanarRoboticsEnv(layout_tiles=np.ones((5, 5)), num_movers=N, collision_params={'shape': shape, 'size': size, 'offset': off})
            ncase = 250
            qs = np.zeros((ncase, N, 7))
            res = np.zeros((ncase, 2), dtype=np.int32)
            for c in range(ncase):
                q = np.zeros((N, 7))
                centre = rng.uniform(0.3, 0.9, 2)
                spread = rng.choice([0.15, 0.25, 0.5])
                q[:, :2] = centre + rng.uniform(-spread, spread, (N, 2))
                if c % 7 == 0:  # exact-threshold pairs
                    d = 2 * (size if shape == 'circle' else size[0])
                    q[1, :2] = q[0, :2] + np.array([d, 0.0])
                if shape == 'box':
                    ang = rng.uniform(-np.pi, np.pi, N)
                    if c % 3 == 0:
                        ang[:] = 0.0
                    q[:, 3:] = _yaw_quat(ang)
                else:
                    q[:, 3] = 1
                qs[c] = q
                for s, safety in enumerate((False, True)):
                    res[c, s] = int(env.check_mover_collision(mover_names=[], c_size=env.c_size, add_safety_offset=safety, mover_qpos=q))
            key = f'mov_{len(mov_meta)}'
            out[key + '_qpos'] = qs
            out[key + '_res'] = res
            mov_meta.append({'key': key, 'shape': shape, 'size': np.asarray(size).tolist(), 'offset': off, 'N': N})
    out['mov_meta'] = np.array(json.dumps(mov_meta))

    # --- rectangles / vertices (geometry_2D_utils.py:72-138)
    n = 3000
    q1 = np.zeros((n, 7))
    q2 = np.zeros((n, 7))
    q1[:, :2] = rng.uniform(0, 0.5, (n, 2))
    q2[:, :2] = q1[:, :2] + rng.uniform(-0.3, 0.3, (n, 2))
    a1 = rng.uniform(-np.pi, np.pi, n)
    a2 = rng.uniform(-np.pi, np.pi, n)
    a1[: n // 4] = 0
    a2[: n // 4] = rng.choice([0, np.pi / 2], n // 4)
    q1[:, 3:] = _yaw_quat(a1)
    q2[:, 3:] = _yaw_quat(a2)
    # un-normalised and slightly tilted quaternions too (the reference normalises in float32 and projects)
    q1[n // 2 :, 3:] *= rng.uniform(0.5, 2.0, (n - n // 2, 1))
    q2[3 * n // 4 :, 4:6] += rng.normal(0, 1e-3, (n - 3 * n // 4, 2))
    s1 = rng.uniform(0.03, 0.12, (n, 2))
    s2 = rng.uniform(0.03, 0.12, (n, 2))
    s1[: n // 4] = 0.08
    s2[: n // 4] = 0.08
    k = n // 8  # touching axis-aligned boxes
    q2[:k, 0] = q1[:k, 0] + 0.16
    q2[:k, 1] = q1[:k, 1] + rng.uniform(-0.2, 0.2, k)
    out['rect_q1'], out['rect_q2'], out['rect_s1'], out['rect_s2'] = q1, q2, s1, s2
    out['rect_res'] = np.asarray(geom.check_rectangles_intersect(qpos_r1=q1, qpos_r2=q2, size_r1=s1, size_r2=s2), dtype=np.int32)
    out['rect_v1'] = geom.get_2D_rect_vertices(qpos=q1, size=s1)

    # --- rewards (planning:459-534, pushing:457-527)
    from gymnasium_planar_robotics.envs.manipulation.benchmark_pushing_env import BenchmarkPushingEnv

    for N in (1, 2, 4):
        env = ref_harness.make_planning_env(layout_tiles=np.ones((3, 3)), num_movers=N)
        b = 400
        dg = rng.uniform(0.11, 0.55, (b, 2 * N))
        ag = dg + rng.normal(0, 0.08, (b, 2 * N))
        ag[: b // 8] = dg[: b // 8]
        ag[b // 8 : b // 4, 0] = dg[b // 8 : b // 4, 0] + 0.1  # exactly on the threshold in x
        ag[b // 8 : b // 4, 1] = dg[b // 8 : b // 4, 1]
        mc = rng.random(b) < 0.15
        wc = rng.random(b) < 0.15
        info = np.array([{'mover_collision': bool(mc[i]), 'wall_collision': bool(wc[i])} for i in range(b)])
        out[f'rew_plan{N}_ag'], out[f'rew_plan{N}_dg'] = ag, dg
        out[f'rew_plan{N}_mc'], out[f'rew_plan{N}_wc'] = mc, wc
        out[f'rew_plan{N}_reward'] = env.compute_reward(ag, dg, info)
        out[f'rew_plan{N}_term'] = env.compute_terminated(ag, dg, info)
    try:
        penv2 = BenchmarkPushingEnv(render_mode=None)
    except Exception:  # the pushing ctor touches mocked MuJoCo objects; build without running __init__
        penv2 = BenchmarkPushingEnv.__new__(BenchmarkPushingEnv)
        penv2.num_movers = 1
        penv2.threshold_pos = 0.05
        penv2.reward_wall_collision = -50
    b = 400
    dg = rng.uniform(0.22, 0.44, (b, 2))
    ag = dg + rng.normal(0, 0.04, (b, 2))
    ag[: b // 8] = dg[: b // 8]
    wc = rng.random(b) < 0.2
    info = np.array([{'mover_collision': False, 'wall_collision': bool(wc[i])} for i in range(b)])
    out['rew_push_ag'], out['rew_push_dg'], out['rew_push_wc'] = ag, dg, wc
    out['rew_push_reward'] = penv2.compute_reward(ag, dg, info)
    out['rew_push_term'] = penv2.compute_terminated(ag, dg, info)
    return out


OBSTACLES_BOX = np.array([[0.5, 0.5, 0.2, 0.1], [0.9, 0.3, 0.05, 0.05]])
OBSTACLES_CIRCLE = np.array([[0.5, 0.5, 0.07], [0.3, 0.9, 0.02]])


def reference_obstacle_vectors(seed=20241018, n=3000):
    """Verdicts of the REFERENCE's own functions for a mover against fixed shapes — what the static-obstacle rules of
    include/gpr.h (gpr_config.num_obstacles) are built from: box  ``geom.check_rectangles_intersect`` (geom:107-138) or the
    mover's centre inside the obstacle; circle  ``check_mover_collision`` (basic:355-424) on the pair (mover, obstacle)."""
    ref_harness.install()
    geom = ref_harness.geometry()
    rng = np.random.default_rng(seed)
    out = {'obst_box': OBSTACLES_BOX, 'obst_circle': OBSTACLES_CIRCLE}
    q = np.zeros((n, 7))
    q[:, :2] = rng.uniform(0.1, 1.1, (n, 2))
    yaw = rng.uniform(-np.pi, np.pi, n)
    q[:, 3], q[:, 6] = np.cos(yaw / 2), np.sin(yaw / 2)
    s = rng.uniform(0.03, 0.1, (n, 2))
    want = np.zeros(n, dtype=bool)
    for o in OBSTACLES_BOX:
        qo = np.tile(np.array([[o[0], o[1], 0.0, 1.0, 0.0, 0.0, 0.0]]), (n, 1))
        want |= geom.check_rectangles_intersect(q, qo, s, np.tile(o[None, 2:], (n, 1)))
        want |= (np.abs(q[:, 0] - o[0]) <= o[2]) & (np.abs(q[:, 1] - o[1]) <= o[3])
    out['box_qpos'], out['box_size'], out['box_hit'] = q, s, want
    env = ref_harness.make_planning_env(layout_tiles=np.ones((5, 5)), num_movers=2, collision_params={'shape': 'circle', 'size': 0.1})
    xy = rng.uniform(0.2, 1.0, (n, 2))
    r = rng.uniform(0.03, 0.12, n)
    hit = np.zeros(n, dtype=bool)
    for i in range(n):
        for o in OBSTACLES_CIRCLE:
            pair = np.zeros((2, 7))
            pair[:, 3] = 1.0
            pair[0, :2], pair[1, :2] = xy[i], o[:2]
            hit[i] |= bool(env.check_mover_collision(mover_names=['m0', 'm1'], c_size=np.array([r[i], o[2]]), add_safety_offset=False,
                                                     mover_qpos=pair))
    out['circle_xy'], out['circle_r'], out['circle_hit'] = xy, r, hit
    return out


if __name__ == '__main__':
    ov = reference_obstacle_vectors()
    np.savez_compressed(os.path.join(HERE, 'reference_obstacle_vectors.npz'), **ov)
    print('obstacle vectors: box hits', int(ov['box_hit'].sum()), 'circle hits', int(ov['circle_hit'].sum()))
    tv = reference_test_vectors()
    with open(os.path.join(HERE, 'reference_test_vectors.json'), 'w') as f:
        json.dump(tv, f)
    print('wall cases', len(tv['wall']), 'segment sets', len(tv['segments']), 'rectangle sets', len(tv['rectangles']))
    rv = reference_random_vectors()
    np.savez_compressed(os.path.join(HERE, 'reference_random_vectors.npz'), **rv)
    print('random vectors', len(rv), 'arrays')
