"""The planar pushing substep against a second, independently written implementation.

``include/gpr_push_physics.h`` is compiled into BOTH the CUDA kernels and the C oracle, so their bit-exact agreement cannot
show that the header implements the model it documents.  ``tests/push_model_numpy.py`` restates that model from MuJoCo's
documentation in generalised coordinates with explicit Jacobians and shares no code with the header.  Stated tolerances:

    * one substep from identical state, any contact configuration: the same number of contact points, and the velocity
      change of each body within float32-solver precision of the float64 model — median relative error < 1e-6, 99th
      percentile < 1e-4, maximum < 5e-3 (measured over 3000 random states: 4e-8 / 3e-6 / 9e-4; with the float64 solve the
      header used until round 2 the same comparison gave 1.2e-11, i.e. the two derivations are the same model; what remains
      is the rounding of the float32 constraint solve)
    * 600-substep push trajectory, both sides warm-started: poses within 1e-5 m / rad, velocities within 1e-4
    * solver truncation + warm start: the shipped setting — 3 sweeps started from the previous substep's forces — against
      the converged solution (300 cold sweeps per substep) over 400-substep pushes that move the object ~17 cm: final
      object position within 3e-4 m (measured mean 2.4e-5, max 1.2e-4), yaw within 3e-4 rad; 8 COLD sweeps (the setting
      until round 2) give 2.2e-5 / 1.5e-4, i.e. the warm start buys the same accuracy with 2.7x fewer sweeps.

Parity with MuJoCo itself stays UNPINNED: ``tests/test_mujoco_parity.py`` runs the moment ``import mujoco`` works.
"""

import ctypes

import numpy as np
import pytest

import gpr_oracle as oracle
import gymnasium_planar_robotics_b200 as gpr
import push_model_numpy as pm

_D = ctypes.c_double


def _c_substep(cfg, M, O, u, warm=None):
    """gpro_push_substep of the oracle library (= include/gpr_push_physics.h); warm: float32[13] in/out or None (cold)."""
    M, O, q = M.copy(), O.copy(), np.zeros(2)
    P = ctypes.POINTER(_D)
    w = None if warm is None else warm.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    nc = oracle.lib().gpro_push_substep(ctypes.byref(cfg), M.ctypes.data_as(P), O.ctypes.data_as(P), _D(u[0]), _D(u[1]), q.ctypes.data_as(P), w)
    return M, O, q, nc


def _random_state(rng, lo=0.09, hi=0.17, rest_share=0.2):
    ang, yaw_m, yaw_o = rng.uniform(-np.pi, np.pi), rng.normal(0, 0.01), rng.uniform(-np.pi, np.pi)
    dist = rng.uniform(lo, hi)
    M = np.array([0.3, 0.3, np.cos(yaw_m), np.sin(yaw_m), *rng.normal(0, 0.5, 2), rng.normal(0, 0.2)])
    O = np.array([0.3 + dist * np.cos(ang), 0.3 + dist * np.sin(ang), np.cos(yaw_o), np.sin(yaw_o), *rng.normal(0, 0.5, 2), rng.normal(0, 3)])
    if rng.random() < rest_share:
        O[4:] = 0.0
    return M, O, rng.uniform(-10, 10, 2)


@pytest.mark.parametrize('kw', [dict(), dict(mover_params={'mass': 0.628, 'size': np.array([0.06, 0.09, 0.006])}, contact_iterations=5)])
def test_single_substep_matches_the_independent_model(kw):
    cfg, _ = gpr.pushing_config(num_envs=1, **kw)
    P = pm.Params(iterations=int(cfg.contact_iterations), m_M=float(cfg.mover_mass), half_M=(cfg.mover_half[0], cfg.mover_half[1]))
    rng = np.random.default_rng(0)
    hist = [0, 0, 0]
    rel = []
    for _ in range(600):
        M, O, u = _random_state(rng)
        a, b = _c_substep(cfg, M, O, u), pm.substep(P, M, O, u)
        assert a[3] == b[3]  # same number of contact points
        hist[a[3]] += 1
        for got, want, before in ((a[0], b[0], M), (a[1], b[1], O)):
            dv_got, dv_want = (got - before)[4:7], (want - before)[4:7]
            rel.append(np.abs(dv_got - dv_want).max() / (np.abs(dv_want).max() + 1e-9))
            assert np.allclose(got[:4], want[:4], rtol=0, atol=1e-8)  # poses after one substep
        assert np.allclose(a[2], b[2], rtol=1e-4, atol=1e-6)  # the mover's qacc
    rel = np.array(rel)
    assert min(hist) > 60  # free, one-point and two-point manifolds all exercised
    assert np.median(rel) < 1e-6 and np.percentile(rel, 99) < 1e-4 and rel.max() < 5e-3, (np.median(rel), np.percentile(rel, 99), rel.max())


def test_push_trajectory_matches_the_independent_model():
    cfg, _ = gpr.pushing_config(num_envs=1)
    P = pm.Params()
    M = np.array([0.3, 0.3, 1, 0, 0, 0, 0.0])
    O = np.array([0.45, 0.32, np.cos(0.3), np.sin(0.3), 0, 0, 0.0])
    M2, O2 = M.copy(), O.copy()
    touched = 0
    wn, wc = {}, np.zeros(13, dtype=np.float32)  # warm-start state of either side
    for k in range(600):
        u = np.array([4.0, 0.5]) if k < 300 else np.array([-4.0, 0.0])
        M, O, _, nc = pm.substep(P, M, O, u, warm=wn)
        M2, O2, _, nc2 = _c_substep(cfg, M2, O2, u, wc)
        assert nc == nc2 or abs(k - 300) < 3
        touched += nc > 0
    assert touched > 50 and O[0] > 0.6  # the object really was pushed
    assert np.abs(M - M2)[:4].max() < 1e-5 and np.abs(O - O2)[:4].max() < 1e-5, (np.abs(M - M2), np.abs(O - O2))
    assert np.abs(M - M2)[4:].max() < 1e-4 and np.abs(O - O2)[4:].max() < 1e-4


def test_three_warm_started_sweeps_track_the_converged_solution():
    P = pm.Params()
    rng = np.random.default_rng(0)
    err_w3, err_c8 = [], []
    for _ in range(5):
        ang, yaw_o = rng.uniform(0, 2 * np.pi), rng.uniform(-1, 1)
        O0 = np.array([0.35, 0.35, np.cos(yaw_o), np.sin(yaw_o), 0, 0, 0.0])
        M0 = np.array([0.35 + 0.16 * np.cos(ang), 0.35 + 0.16 * np.sin(ang), 1, 0, 0, 0, 0.0])
        d = -np.array([np.cos(ang), np.sin(ang)]) + rng.normal(0, 0.3, 2)

        def run(iterations, warm):
            M, O = M0.copy(), O0.copy()
            for k in range(400):
                M, O, _, _ = pm.substep(P, M, O, 8 * d if k < 150 else -8 * d, iterations, warm)
            return O

        ref = run(300, None)
        assert np.linalg.norm(ref[:2] - O0[:2]) > 0.05  # the object really was pushed
        for errs, O in ((err_w3, run(3, {})), (err_c8, run(8, None))):
            errs.append((np.linalg.norm(O[:2] - ref[:2]), abs(np.arctan2(O[3], O[2]) - np.arctan2(ref[3], ref[2]))))
    err_w3, err_c8 = np.array(err_w3), np.array(err_c8)
    assert err_w3[:, 0].max() < 3e-4 and err_w3[:, 1].max() < 3e-4, err_w3
    assert err_w3[:, 0].mean() < 2.0 * err_c8[:, 0].mean() + 1e-5  # as good as 8 cold sweeps
