"""Pin the CPU oracle (oracle/gpr_oracle.c) against the committed golden vectors.

The vectors were produced by executing the UNMODIFIED reference functions (tests/golden/make_golden.py): the reference's
own 100 wall-check cases and 34 geometry cases, plus seeded random inputs.  Every comparison is exact: flags bit-for-bit,
float64 outputs bit-for-bit (the oracle evaluates the reference's expressions in the reference's order).
"""

import json
import os

import numpy as np
import pytest

import gpr_oracle as oracle
import gymnasium_planar_robotics_b200 as gpr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

LAYOUTS = {
    'full3x3': np.ones((3, 3)),
    'full5x5': np.ones((5, 5)),
    'L': np.array([[1, 1], [1, 0]]),
    'hole3x3': np.array([[1, 1, 1], [1, 0, 1], [1, 1, 1]]),
    'ragged': np.array([[1, 1, 0, 1], [1, 1, 1, 1], [0, 1, 1, 0], [1, 1, 1, 1], [1, 0, 1, 1]]),
    'strip': np.ones((1, 4)),
}


@pytest.fixture(scope='module')
def tv():
    with open(os.path.join(GOLDEN, 'reference_test_vectors.json')) as f:
        return json.load(f)


@pytest.fixture(scope='module')
def rv():
    return np.load(os.path.join(GOLDEN, 'reference_random_vectors.npz'))


def _cfg(layout, shape, num_movers=1, **kw):
    cp = {'shape': shape}
    if shape == 'box':
        cp['size'] = np.array([0.08, 0.08])
    cfg, _ = gpr.planning_config(num_envs=1, layout_tiles=np.asarray(layout), num_movers=num_movers, collision_params=cp, std_noise=0.0, **kw)
    return cfg


def test_reference_wall_cases(tv):
    """tests/test_basic_env.py:10-1633 — all 100 cases, against the reference's expected vectors."""
    assert len(tv['wall']) == 100
    oracle.lib().gpro_assert_trips(1)
    for i, case in enumerate(tv['wall']):
        cfg = _cfg(case['layout'], case['shape'])
        got = oracle.qpos_is_valid(cfg, np.array(case['qpos']), np.array(case['csize_total']))
        assert got.tolist() == case['expected'], f'wall case {i}'
        assert case['reference_output'] == case['expected']
    assert oracle.lib().gpro_assert_trips(1) == 0


def test_reference_segment_cases(tv):
    """tests/test_geometry_2D_utils.py:10-100"""
    n = 0
    for case in tv['segments']:
        for p1, p2, q1, q2, e in zip(case['p1'], case['p2'], case['q1'], case['q2'], case['expected']):
            assert oracle.segments_intersect(p1, p2, q1, q2) == bool(e)
            n += 1
    assert n >= 17


def test_reference_rectangle_cases(tv):
    """tests/test_geometry_2D_utils.py:103-164"""
    n = 0
    for case in tv['rectangles']:
        for r1, r2, s1, s2, e in zip(case['qpos_r1'], case['qpos_r2'], case['size_r1'], case['size_r2'], case['expected']):
            assert oracle.rectangles_intersect(r1, r2, s1, s2) == bool(e)
            n += 1
    assert n >= 17


def test_ensure_max_dyn_val_bit_exact(rv):
    """planning:610-645 on 4000 random inputs: float64 outputs identical to the last bit."""
    cur, der, mx = rv['emdv_cur'], rv['emdv_der'], rv['emdv_max']
    for i in range(cur.shape[0]):
        nv, nd = oracle.ensure_max_dyn_val(cur[i], float(mx[i]), der[i], 0.001)
        assert np.array_equal(nv, rv['emdv_next'][i]) and np.array_equal(nd, rv['emdv_next_der'][i]), i


def test_wall_random(rv):
    meta = json.loads(str(rv['wall_meta']))
    total = 0
    for m in meta:
        cfg = _cfg(LAYOUTS[m['layout']], m['shape'])
        got = oracle.qpos_is_valid(cfg, rv[m['key'] + '_qpos'], rv[m['key'] + '_csize'])
        ref = rv[m['key'] + '_valid']
        assert np.array_equal(got, ref), (m, np.nonzero(got != ref)[0][:10])
        total += ref.size
    assert total >= 10000


def test_mover_collision_random(rv):
    meta = json.loads(str(rv['mov_meta']))
    for m in meta:
        N = m['N']
        cfg = _cfg(np.ones((5, 5)), m['shape'], num_movers=N)
        qs, res = rv[m['key'] + '_qpos'], rv[m['key'] + '_res']
        size = np.asarray(m['size'], dtype=np.float64)
        for c in range(qs.shape[0]):
            for s in (0, 1):
                cs = np.tile(np.atleast_1d(size + m['offset'] * s), (N, 1))
                assert oracle.check_mover_collision(cfg, qs[c], cs) == bool(res[c, s]), (m, c, s)


def test_rectangles_random(rv):
    q1, q2, s1, s2 = rv['rect_q1'], rv['rect_q2'], rv['rect_s1'], rv['rect_s2']
    got = np.array([oracle.rectangles_intersect(q1[i], q2[i], s1[i], s2[i]) for i in range(q1.shape[0])])
    assert np.array_equal(got, rv['rect_res'].astype(bool))
    assert 0.2 < got.mean() < 0.9


def test_rect_vertices(rv):
    """geom:72-104 incl. the float32 quaternion normalisation (rot:447). The reference's 3x3 @ 3x4 product goes through
    BLAS (FMA), so the last bit may differ; 4 ulp of the coordinate magnitude is the bound checked."""
    q1, s1, v1 = rv['rect_q1'], rv['rect_s1'], rv['rect_v1']
    worst = 0.0
    for i in range(q1.shape[0]):
        v = oracle.rect_vertices(q1[i], s1[i])
        worst = max(worst, np.abs(v - v1[i]).max())
    assert worst <= 4 * np.finfo(np.float64).eps, worst


@pytest.mark.parametrize('N', [1, 2, 4])
def test_planning_reward(rv, N):
    cfg = _cfg(np.ones((3, 3)), 'circle', num_movers=N)
    ag, dg = rv[f'rew_plan{N}_ag'], rv[f'rew_plan{N}_dg']
    # the oracle entry point takes float32 goals (the product's I/O dtype): evaluate the reference on the same values
    r, t = oracle.compute_reward(cfg, ag, dg, rv[f'rew_plan{N}_mc'], rv[f'rew_plan{N}_wc'])
    same = (ag.astype(np.float32) == ag).all(axis=1) & (dg.astype(np.float32) == dg).all(axis=1)
    ref_r, ref_t = rv[f'rew_plan{N}_reward'], rv[f'rew_plan{N}_term']
    # float32 rounding of the goals can only move a distance across the threshold when it is within ~1e-8 of it
    d = np.linalg.norm((ag - dg).reshape(-1, N, 2), axis=2)
    safe = (np.abs(d - 0.1) > 1e-6).all(axis=1) | same
    assert np.array_equal(r[safe], ref_r[safe].astype(np.float32))
    assert np.array_equal(t[safe], ref_t[safe])
    assert safe.mean() > 0.8


def test_pushing_reward(rv):
    cfg, _ = gpr.pushing_config(num_envs=1, std_noise=0.0)
    ag, dg = rv['rew_push_ag'], rv['rew_push_dg']
    r, t = oracle.compute_reward(cfg, ag, dg, None, rv['rew_push_wc'])
    d = np.linalg.norm(ag - dg, axis=1)
    safe = np.abs(d - 0.05) > 1e-6
    assert np.array_equal(r[safe], rv['rew_push_reward'][safe].astype(np.float32))
    assert np.array_equal(t[safe], rv['rew_push_term'][safe])


def check_obstacle_vectors(ov):
    """The oracle's static-obstacle test against verdicts of the reference's own geometry / mover-check functions."""
    cfg_b, _ = gpr.planning_config(num_envs=1, layout_tiles=np.ones((5, 5)), num_movers=1, obstacles=ov['obst_box'], std_noise=0.0,
                                   collision_params={'shape': 'box', 'size': np.array([0.08, 0.06])})
    got = np.array([oracle.check_obstacle_collision(cfg_b, ov['box_qpos'][i:i + 1], ov['box_size'][i:i + 1]) for i in range(len(ov['box_hit']))])
    assert np.array_equal(got, ov['box_hit'])
    cfg_c, _ = gpr.planning_config(num_envs=1, layout_tiles=np.ones((5, 5)), num_movers=1, obstacles=ov['obst_circle'], std_noise=0.0)
    q = np.zeros((1, 7))
    q[0, 3] = 1.0
    got = []
    for xy, r in zip(ov['circle_xy'], ov['circle_r']):
        q[0, :2] = xy
        got.append(oracle.check_obstacle_collision(cfg_c, q, r))
    assert np.array_equal(np.array(got), ov['circle_hit'])
    assert 100 < ov['box_hit'].sum() < len(ov['box_hit']) - 100 and 100 < ov['circle_hit'].sum() < len(ov['circle_hit']) - 100


def test_obstacle_rules_against_reference_vectors():
    check_obstacle_vectors(np.load(os.path.join(GOLDEN, 'reference_obstacle_vectors.npz')))


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10 and philox4x32-7."""
    assert oracle.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2).tolist() == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]).tolist() == [
        0xD16CFE09,
        0x94FDCCEB,
        0x5001E420,
        0x24126EA1,
    ]
    # philox4x32-7 (the rejection-sampling streams, gpr_rng_block_sampling): Random123's kat_vectors for 7 rounds
    assert oracle.philox([0, 0, 0, 0], [0, 0], 7).tolist() == [0x5F6FB709, 0x0D893F64, 0x4F121F81, 0x4F730A48]
    assert oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, 7).tolist() == [0x5207DDC2, 0x45165E59, 0x4D8EE751, 0x8C52F662]
    assert oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], 7).tolist() == [
        0x4DFCCABA, 0x190A87F0, 0xC47362BA, 0xB6B5242A]


def test_portable_normals_are_standard_normal():
    from scipy import stats

    n = oracle.normals(20240607, 3, 5, 0, 0, 100000).astype(np.float64)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1.0) < 0.01
    assert stats.kstest(n, 'norm').pvalue > 1e-3
    assert abs(stats.skew(n)) < 0.02 and abs(stats.kurtosis(n)) < 0.05
    assert abs(np.corrcoef(n[0::2], n[1::2])[0, 1]) < 0.01
