"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA step path, called through the C ABI, against the CPU
oracle on the same seeded inputs.

Bar (BASELINE.json north_star): collision / termination / done flags bit-exact; positions, velocities, observations and
rewards within a float32 tolerance.  The path computes in float64 in the reference's operation order, so the tests demand
MORE than the bar: the float64 state is compared for exact equality, and the float32 outputs must equal the oracle's
float64 values rounded once.
"""

import os

import numpy as np
import pytest
import torch

import gpr_oracle as oracle
import gymnasium_planar_robotics_b200 as gpr

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'
ROOT = __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))

L_RAGGED = np.array([[1, 1, 0, 1], [1, 1, 1, 1], [0, 1, 1, 0], [1, 1, 1, 1], [1, 0, 1, 1]])
L_HOLE = np.array([[1, 1, 1, 1], [1, 0, 1, 1], [1, 1, 1, 1], [1, 1, 1, 1]])


def make_pair(num_envs, **kw):
    env = gpr.BenchmarkPlanningVecEnv(num_envs, device=DEV, **kw)
    cfg, _ = gpr.planning_config(num_envs=num_envs, **kw)
    return env, oracle.OracleEnv(cfg, nthreads=oracle.max_threads())


def sample_starts(rng, ora, spread=1.0):
    """Collision-free-ish injected starts/goals inside the spawn box (validity is not required: flags must still agree)."""
    c = ora.cfg
    lo = np.array([c.min_xy_pos[0], c.min_xy_pos[1]])
    hi = np.array([c.max_xy_pos[0], c.max_xy_pos[1]])
    st = rng.uniform(lo, lo + (hi - lo) * spread, (ora.B, ora.N, 2))
    gl = rng.uniform(lo, hi, (ora.B, ora.N, 2))
    return st, gl


def assert_outputs_equal(env, ora, what):
    b = env.core.buf
    torch.cuda.synchronize()
    for k in ('terminated', 'truncated', 'is_success', 'mover_collision', 'wall_collision', 'other_collision'):
        if k not in b:
            continue  # (other_collision exists only with static obstacles)
        got, ref = b[k].cpu().numpy(), getattr(ora, k)
        assert np.array_equal(got, ref), f'{what}: {k} differs in {np.count_nonzero(got != ref)} envs'
    assert np.array_equal(b['reward'].cpu().numpy(), ora.reward.astype(np.float32)), f'{what}: reward'
    for k in ('observation', 'achieved_goal', 'desired_goal'):
        got, ref = b[k].cpu().numpy(), getattr(ora, k).astype(np.float32)
        assert np.array_equal(got, ref), f'{what}: {k} max|d|={np.abs(got - ref).max()}'


def assert_state_equal(env, ora, what):
    st = env.get_state()
    torch.cuda.synchronize()
    for k in ('pos', 'vel', 'acc', 'goal'):
        got, ref = st[k].cpu().numpy(), getattr(ora, k)
        assert np.array_equal(got, ref), f'{what}: state {k} max|d|={np.abs(got - ref).max()}'
    assert np.array_equal(st['elapsed_steps'].cpu().numpy(), ora.elapsed_steps), f'{what}: elapsed'
    assert np.array_equal(st['rng_counter'].cpu().numpy().view(np.uint32), ora.rng_counter), f'{what}: rng counter'


def run_lockstep(env, ora, steps, rng, scale, inject=True, seed=3):
    if inject:
        st, gl = sample_starts(rng, ora)
        env.reset(seed=seed, options={'mover_start_xy_pos': st, 'mover_goal_xy_pos': gl})
        ora.reset(seed=seed, inject_start=st, inject_goal=gl)
    else:
        env.reset(seed=seed)
        ora.reset(seed=seed)
    torch.cuda.synchronize()
    for k in ('observation', 'achieved_goal', 'desired_goal'):
        assert np.array_equal(env.core.buf[k].cpu().numpy(), getattr(ora, k).astype(np.float32)), f'reset {k}'
    for k in ('is_success', 'mover_collision', 'wall_collision', 'other_collision'):
        if k in env.core.buf:
            assert np.array_equal(env.core.buf[k].cpu().numpy(), getattr(ora, k)), f'reset {k}'
    assert_state_equal(env, ora, 'reset')
    events = 0
    for t in range(steps):
        a = rng.uniform(-scale, scale, (ora.B, ora.action_dim)).astype(np.float32)
        env.step(torch.as_tensor(a, device=DEV))
        ora.step(a)
        assert_outputs_equal(env, ora, f'step {t}')
        events += int(ora.terminated.sum())
    assert_state_equal(env, ora, 'final')
    return events


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('learn_jerk', [False, True])
@pytest.mark.parametrize('num_movers', [1, 2, 3, 4])
def test_planning_circle_injected(num_movers, learn_jerk):
    """BASELINE configs[0]/[1] shape: circle r=0.11 on 3x3, acc and jerk mode, no noise, injected starts and goals."""
    rng = np.random.default_rng(100 + num_movers + 10 * learn_jerk)
    env, ora = make_pair(1024, layout_tiles=np.ones((3, 3)), num_movers=num_movers, std_noise=0.0, learn_jerk=learn_jerk,
                         autoreset_mode='off', max_episode_steps=50)
    lim = 100.0 if learn_jerk else 10.0
    ev = run_lockstep(env, ora, 30, rng, 1.3 * lim)
    assert ev > 0  # collisions did occur and were reproduced
    env.close()


def test_planning_velocity_and_acceleration_clipping():
    """The reference's own closed-form test setup (tests/test_benchmark_planning_env.py:31-45): tiny v_max / a_max on a
    9x9 layout so both ensure_max_dyn_val clips are active nearly every cycle."""
    rng = np.random.default_rng(5)
    for jerk in (False, True):
        env, ora = make_pair(512, layout_tiles=np.ones((9, 9)), num_movers=2, std_noise=0.0, learn_jerk=jerk, v_max=0.01,
                             a_max=0.2, j_max=150.0, num_cycles=42, autoreset_mode='off', max_episode_steps=0)
        st = np.tile(np.array([[0.96, 0.96], [1.2, 1.2]]), (512, 1, 1))
        gl = np.tile(np.array([[0.5, 0.5], [1.5, 1.5]]), (512, 1, 1))
        env.reset(seed=1, options={'mover_start_xy_pos': st, 'mover_goal_xy_pos': gl})
        ora.reset(seed=1, inject_start=st, inject_goal=gl)
        for t in range(40):
            a = rng.uniform(-150, 150, (512, 4)).astype(np.float32)
            env.step(torch.as_tensor(a, device=DEV))
            ora.step(a)
        assert_state_equal(env, ora, f'clip jerk={jerk}')
        s = env.get_state()
        vn = s['vel'].norm(dim=-1).max().item()
        an = s['acc'].norm(dim=-1).max().item()
        assert vn <= 0.01 * (1 + 1e-12)  # tests/test_benchmark_planning_env.py:116-117
        if jerk:  # acc mode: the action Box clips per component, only jerk mode norm-clips the acceleration
            assert an <= 0.2 * (1 + 1e-9)
        env.close()


@pytest.mark.parametrize('layout', [np.ones((5, 5)), L_RAGGED, L_HOLE], ids=['5x5', 'ragged', 'hole'])
@pytest.mark.parametrize('shape', ['circle', 'box'])
def test_planning_layouts_and_shapes(layout, shape):
    """Arbitrary tile layouts (inner corners, holes, missing tiles) with circle and box collision shapes."""
    rng = np.random.default_rng(11)
    cp = {'shape': shape, 'size': 0.09 if shape == 'circle' else np.array([0.08, 0.06]), 'offset': 0.004, 'offset_wall': 0.002}
    env, ora = make_pair(1024, layout_tiles=layout, num_movers=4, std_noise=0.0, collision_params=cp, autoreset_mode='off')
    # starts anywhere over the grid, including over missing tiles and borders
    W, H = layout.shape[0] * 0.24, layout.shape[1] * 0.24
    st = rng.uniform([0.01, 0.01], [W - 0.01, H - 0.01], (1024, 4, 2))
    st[:64] = np.round(st[:64] / 0.12) * 0.12  # exactly on tile borders / centres
    st = np.clip(st, 0.0, [W, H])
    gl = rng.uniform([0.1, 0.1], [W - 0.1, H - 0.1], (1024, 4, 2))
    env.reset(seed=2, options={'mover_start_xy_pos': st, 'mover_goal_xy_pos': gl})
    ora.reset(seed=2, inject_start=st, inject_goal=gl)
    torch.cuda.synchronize()
    assert np.array_equal(env.core.buf['wall_collision'].cpu().numpy(), ora.wall_collision)
    assert np.array_equal(env.core.buf['mover_collision'].cpu().numpy(), ora.mover_collision)
    assert 0.05 < ora.wall_collision.mean() < 0.999
    for t in range(10):
        a = rng.uniform(-10, 10, (1024, 8)).astype(np.float32)
        env.step(torch.as_tensor(a, device=DEV))
        ora.step(a)
        assert_outputs_equal(env, ora, f'{shape} step {t}')
    assert_state_equal(env, ora, shape)
    env.close()


def test_planning_box_eight_movers_jerk():
    """BASELINE configs[3] shape: 8 movers, box 0.08 x 0.08, learn_jerk=True, 5x5 tiles."""
    rng = np.random.default_rng(12)
    env, ora = make_pair(1024, layout_tiles=np.ones((5, 5)), num_movers=8, std_noise=0.0, learn_jerk=True,
                         collision_params={'shape': 'box', 'size': np.array([0.08, 0.08])}, autoreset_mode='off')
    ev = run_lockstep(env, ora, 25, rng, 120.0)
    assert ev > 0
    env.close()


@pytest.mark.parametrize('shape', ['circle', 'box'])
def test_planning_with_sensor_noise(shape):
    """std_noise = reference default 1e-5 (and a large 1e-3): the portable RNG makes the noise bit-identical on CPU and
    GPU, so even noisy runs must agree exactly."""
    for sigma in (1e-5, np.array([1e-3, 2e-3, 0.0])):
        rng = np.random.default_rng(13)
        cp = {'shape': shape} if shape == 'circle' else {'shape': 'box', 'size': np.array([0.08, 0.08])}
        env, ora = make_pair(768, layout_tiles=np.ones((4, 4)), num_movers=3, std_noise=sigma, learn_jerk=True,
                             collision_params=cp, autoreset_mode='off')
        run_lockstep(env, ora, 12, rng, 100.0)
        env.close()


@pytest.mark.parametrize('mode', ['same_step', 'next_step'])
@pytest.mark.parametrize('num_movers,layout', [(4, np.ones((3, 3))), (2, L_RAGGED), (1, np.ones((3, 3)))], ids=['4m3x3', '2mragged', '1m'])
def test_planning_autoreset_sampled_on_device(mode, num_movers, layout):
    """Full episodes with on-device rejection sampling of starts and goals (planning:355-418) and TimeLimit(50)."""
    rng = np.random.default_rng(14)
    env, ora = make_pair(512, layout_tiles=layout, num_movers=num_movers, std_noise=1e-5, autoreset_mode=mode, seed=99,
                         max_episode_steps=12)
    env.reset(seed=99)
    ora.reset(seed=99)
    done_total = 0
    for t in range(40):
        a = rng.uniform(-10, 10, (512, 2 * num_movers)).astype(np.float32)
        env.step(torch.as_tensor(a, device=DEV))
        ora.step(a)
        assert_outputs_equal(env, ora, f'{mode} step {t}')
        done = (ora.terminated | ora.truncated).astype(bool)
        done_total += int(done.sum())
        if mode == 'same_step' and done.any():
            torch.cuda.synchronize()
            for k in ('observation', 'achieved_goal', 'desired_goal'):
                got = env.core.buf['final_' + k].cpu().numpy()[done]
                assert np.array_equal(got, getattr(ora, 'final_' + k).astype(np.float32)[done]), f'final_{k}'
    assert done_total > 100
    assert_state_equal(env, ora, mode)
    assert env.core.reset_failures() == int(ora.reset_failed.sum()) == 0
    # sampled starts respect the reference's spawn box and separation
    st = env.get_state()
    assert (st['goal'] >= 0.11 - 1e-12).all() and (st['goal'] <= layout.shape[0] * 0.24).all()
    env.close()


@pytest.mark.parametrize('num_movers,shape,jerk', [(12, 'circle', False), (20, 'box', True), (32, 'circle', True), (7, 'box', False)],
                         ids=['12circle', '20box-jerk', '32circle-jerk', '7box'])
def test_planning_wide_lane_groups(num_movers, shape, jerk):
    """Lane groups of 8, 16 and 32 (one warp per env at the maximum of 32 movers): full episodes with on-device
    rejection sampling (the one-group-per-attempt sampler), sensor noise, auto-reset, on a 10x10 (16x16) layout with a hole."""
    rng = np.random.default_rng(70 + num_movers)
    n = 16 if num_movers > 24 else 10  # 32 movers need room: acceptance of all-at-once sampling is ~1e-2 on 16x16
    layout = np.ones((n, n))
    layout[4, 5] = 0
    cp = {'shape': 'circle', 'size': 0.1, 'offset': 0.002} if shape == 'circle' else {'shape': 'box', 'size': np.array([0.08, 0.07]), 'offset_wall': 0.001}
    env, ora = make_pair(256, layout_tiles=layout, num_movers=num_movers, std_noise=1e-5, learn_jerk=jerk, collision_params=cp,
                         autoreset_mode='same_step', max_episode_steps=8, seed=3)
    env.reset(seed=3)
    ora.reset(seed=3)
    assert_state_equal(env, ora, 'reset')
    lim = 100.0 if jerk else 10.0
    ends = 0
    for t in range(24):
        a = rng.uniform(-lim, lim, (256, 2 * num_movers)).astype(np.float32)
        env.step(torch.as_tensor(a, device=DEV))
        ora.step(a)
        assert_outputs_equal(env, ora, f'N={num_movers} step {t}')
        ends += int((ora.terminated | ora.truncated).sum())
    assert ends > 256
    assert_state_equal(env, ora, f'N={num_movers}')
    assert env.core.reset_failures() == int(ora.reset_failed.sum())
    env.close()


def test_reference_wall_vectors_directly_on_the_gpu():
    """The reference's OWN wall-check test vectors (tests/test_basic_env.py, committed as tests/golden/
    reference_test_vectors.json together with the outputs of the unmodified reference function) against the CUDA path
    directly, without the oracle in between: every test position becomes the injected start of a one-mover env on the
    test's layout and collision size, and reset()'s wall flag (basic_envs.py:1799-1801) must equal 1 - expected.
    Cases with rotated movers (yaw 45 / 90 deg) are not reachable in the planning env (movers never rotate) and are
    covered through the oracle (tests/test_oracle_golden.py)."""
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_test_vectors.json')) as f:
        cases = json.load(f)['wall']
    used = points = 0
    for c in cases:
        q = np.asarray(c['qpos'], dtype=np.float64)
        cs = np.asarray(c['csize_total'], dtype=np.float64)
        if not np.allclose(q[:, 3:], [1, 0, 0, 0]):
            continue
        assert c['expected'] == c['reference_output']  # the reference passes its own test
        layout = np.asarray(c['layout'])
        for size in np.unique(cs, axis=0):
            sel = np.all(cs == size, axis=1)
            cp = {'shape': c['shape'], 'size': float(size[0]) if c['shape'] == 'circle' else size.copy()}
            try:
                env = gpr.BenchmarkPlanningVecEnv(int(sel.sum()), layout_tiles=layout, num_movers=1, device=DEV, std_noise=0.0,
                                                  collision_params=cp, autoreset_mode='off')
            except gpr.GprError:
                continue  # shapes as large as a tile trip an assert in the reference (basic_envs.py:650): refused up front
            st = q[sel][:, None, :2]
            env.reset(seed=0, options={'mover_start_xy_pos': st, 'mover_goal_xy_pos': st})
            torch.cuda.synchronize()
            got = env.core.buf['wall_collision'].cpu().numpy().astype(int)
            exp = 1 - np.asarray(c['expected'])[sel]
            assert np.array_equal(got, exp), (c['shape'], size, layout.tolist(), q[sel][got != exp])
            points += int(sel.sum())
            env.close()
        used += 1
    assert used >= 40 and points >= 400


def test_sampled_starts_follow_the_reference_distribution():
    """Reset sampling cannot be compared draw for draw (NumPy's PCG64 vs Philox), so it is compared in distribution
    against 4,096 resets drawn by the reference's OWN ``_reset_callback`` loop (planning:355-418, run unmodified through
    the harness; committed as tests/golden/reference_reset_samples.npz by tests/golden/make_reset_samples.py):
    two-sample Kolmogorov-Smirnov on coordinates and pair distances of starts and goals."""
    from scipy import stats

    B, N = 16384, 4
    ref = np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_reset_samples.npz'))
    env = gpr.BenchmarkPlanningVecEnv(B, layout_tiles=np.ones((3, 3)), num_movers=N, device=DEV, std_noise=0.0, seed=11)
    env.reset(seed=11)
    p = env.get_state()['pos'].cpu().numpy()
    g = env.get_state()['goal'].cpu().numpy()
    env.close()

    def feats(x):
        d = np.linalg.norm(x[:, :, None] - x[:, None], axis=-1)[:, np.triu_indices(N, 1)[0], np.triu_indices(N, 1)[1]]
        return {'x': x[..., 0].ravel(), 'y': x[..., 1].ravel(), 'min_pair': d.min(axis=1), 'max_pair': d.max(axis=1), 'x0': x[:, 0, 0]}

    for name, arr in (('start', p), ('goal', g)):
        fa, fb = feats(arr), feats(ref[name])
        assert (fb['min_pair'] >= 0.22).all() and ref[name].min() >= 0.11 and ref[name].max() <= 0.55  # the reference's own sample
        for k in fa:
            pv = stats.ks_2samp(fa[k], fb[k]).pvalue
            assert pv > 1e-4, f'{name} {k}: KS p-value {pv}'
        assert (fa['min_pair'] >= 0.22).all() and arr.min() >= 0.11 and arr.max() <= 0.55


def test_planning_per_mover_radii_and_quirk():
    """Per-mover circle radii: pairwise semantics by default, the reference's broadcast quirk (basic_envs.py:409) on demand."""
    for quirk in (False, True):
        rng = np.random.default_rng(15)
        cp = {'shape': 'circle', 'size': np.array([0.06, 0.11, 0.08])}
        env, ora = make_pair(1024, layout_tiles=np.ones((4, 4)), num_movers=3, std_noise=0.0, collision_params=cp,
                             autoreset_mode='off', reference_quirks=quirk)
        st = rng.uniform(0.12, 0.84, (1024, 3, 2))
        st[:, 1] = st[:, 0] + rng.uniform(-0.2, 0.2, (1024, 2))
        gl = rng.uniform(0.12, 0.84, (1024, 3, 2))
        env.reset(seed=4, options={'mover_start_xy_pos': st, 'mover_goal_xy_pos': gl})
        ora.reset(seed=4, inject_start=st, inject_goal=gl)
        torch.cuda.synchronize()
        assert np.array_equal(env.core.buf['mover_collision'].cpu().numpy(), ora.mover_collision)
        assert 0.05 < ora.mover_collision.mean() < 0.95
        env.close()


def test_threshold_edges_are_bit_exact():
    """Movers placed exactly on the decision thresholds: centre distance == r_i + r_j (collision, '<='), one ulp beyond
    (no collision); wall distance == c (collision, strict '<') and one ulp inside."""
    B = 8
    env, ora = make_pair(B, layout_tiles=np.ones((3, 3)), num_movers=2, std_noise=0.0, autoreset_mode='off')
    st = np.zeros((B, 2, 2))
    st[:, 0] = [0.3, 0.3]
    st[:, 1] = [0.3 + 0.22, 0.3]
    st[1, 1, 0] = np.nextafter(0.3 + 0.22, 1.0)
    st[2, 0] = [0.11, 0.3]                       # x - c == 0 -> not strictly inside
    st[2, 1] = [0.5, 0.5]
    st[3, 0] = [np.nextafter(0.11, 1.0), 0.3]
    st[3, 1] = [0.5, 0.5]
    st[4, 0] = [0.61, 0.3]                       # x + c == 0.72 (computed as centre + half)
    st[4, 1] = [0.3, 0.6]
    st[5, 0] = [np.nextafter(0.61, 0.0), 0.3]
    st[5, 1] = [0.3, 0.6]
    st[6, 0] = [0.24, 0.24]                      # on the corner shared by four tiles
    st[6, 1] = [0.48, 0.48]
    st[7, 0] = [0.36, 0.36]
    st[7, 1] = [0.36, 0.36 + 0.22]
    gl = np.tile(np.array([[0.2, 0.2], [0.5, 0.5]]), (B, 1, 1))
    env.reset(seed=0, options={'mover_start_xy_pos': st, 'mover_goal_xy_pos': gl})
    ora.reset(seed=0, inject_start=st, inject_goal=gl)
    torch.cuda.synchronize()
    mc, wc = env.core.buf['mover_collision'].cpu().numpy(), env.core.buf['wall_collision'].cpu().numpy()
    assert np.array_equal(mc, ora.mover_collision) and np.array_equal(wc, ora.wall_collision)
    # independent NumPy evaluation of the reference's expressions (basic_envs.py:409 and, for a full rectangular
    # layout, SURVEY.md spec 5: 0 < x-c and x+c < W strictly, W = last centre + half size)
    exp_mc = np.linalg.norm(st[:, 0] - st[:, 1], axis=1) <= (0.11 + 0.11)
    hi = np.linspace(0.12, 0.6, 3)[-1] + 0.12
    ok = (0.0 < st - 0.11) & (st + 0.11 < hi)
    exp_wc = ~ok.all(axis=(1, 2))
    assert np.array_equal(mc.astype(bool), exp_mc) and np.array_equal(wc.astype(bool), exp_wc)
    assert exp_mc[7] and not exp_mc[1] and exp_wc[2] and not exp_wc[3] and exp_wc[4] and not exp_wc[5] and not exp_wc[6]
    env.close()


def test_compute_reward_kernel_matches_oracle():
    """HER relabelling entry point (planning:459-534): batched rewards / terminated on device."""
    rng = np.random.default_rng(16)
    env, ora = make_pair(4, layout_tiles=np.ones((3, 3)), num_movers=4, std_noise=0.0)
    b = 20000
    dg = rng.uniform(0.11, 0.55, (b, 8)).astype(np.float32)
    ag = (dg + rng.normal(0, 0.07, (b, 8))).astype(np.float32)
    ag[:500] = dg[:500]
    mc, wc = rng.random(b) < 0.1, rng.random(b) < 0.1
    r, t = env.core.compute_reward(ag, dg, mc, wc)
    r2, t2 = oracle.compute_reward(ora.cfg, ag, dg, mc, wc)
    assert np.array_equal(r.cpu().numpy(), r2) and np.array_equal(t.cpu().numpy(), t2)
    assert len(np.unique(r2)) >= 5
    env.close()


def test_float64_outputs_make_compute_reward_consistent_with_the_step():
    """the step decides goal-reached / reward / terminated on float64 values; float32 outputs are those values
    rounded once, so re-deriving the reward from float32 goals can disagree within ~3e-8 m of threshold_pos.  With
    GPR_OUT_FLOAT64 (what the single-env drop-in classes use) the outputs are the reference's dtype, equal the oracle's
    float64 values exactly, and compute_reward(achieved_goal, desired_goal, info) == the step's reward for EVERY env —
    here 200,000 movers parked within +-2e-7 m of the threshold."""
    B = 100000
    kw = dict(layout_tiles=np.ones((4, 4)), num_movers=2, std_noise=0.0, autoreset_mode='off', max_episode_steps=0, seed=2)
    env = gpr.BenchmarkPlanningVecEnv(B, device=DEV, float64_outputs=True, **kw)
    e32 = gpr.BenchmarkPlanningVecEnv(B, device=DEV, **kw)
    cfg, _ = gpr.planning_config(num_envs=B, **kw)
    ora = oracle.OracleEnv(cfg, nthreads=oracle.max_threads())
    rng = np.random.default_rng(4)
    goal = np.tile(np.array([[0.25, 0.25], [0.7, 0.7]]), (B, 1, 1))
    ang = rng.uniform(0, 2 * np.pi, (B, 2))
    dist = 0.1 + rng.uniform(-2e-7, 2e-7, (B, 2))
    start = goal + dist[..., None] * np.stack([np.cos(ang), np.sin(ang)], axis=-1)
    opts = {'mover_start_xy_pos': start, 'mover_goal_xy_pos': goal}
    obs, info = env.reset(seed=2, options=opts)
    e32.reset(seed=2, options=opts)
    ora.reset(seed=2, inject_start=start, inject_goal=goal)
    assert obs['achieved_goal'].dtype == torch.float64
    a = torch.zeros((B, 4), device=DEV)
    obs, r, term, trunc, info = env.step(a)
    o32, r32, *_, i32 = e32.step(a)
    ora.step(np.zeros((B, 4), np.float32))
    torch.cuda.synchronize()
    for k in ('observation', 'achieved_goal', 'desired_goal'):
        assert np.array_equal(obs[k].cpu().numpy(), getattr(ora, k)), k  # float64, unrounded
    assert torch.equal(r, r32) and np.array_equal(r.cpu().numpy(), ora.reward.astype(np.float32))
    assert 0.2 < float((r == 50).float().mean()) < 0.3 and not bool((r == -50).any())  # both movers inside in about a quarter of the envs
    rr = env.compute_reward(obs['achieved_goal'], obs['desired_goal'], info)
    tt = env.compute_terminated(obs['achieved_goal'], obs['desired_goal'], info)
    assert torch.equal(rr, r) and torch.equal(tt, term)  # exact for every env
    rr32 = e32.compute_reward(o32['achieved_goal'], o32['desired_goal'], i32)
    assert int((rr32 != r32).sum()) > 0  # the float32 outputs cannot guarantee it this close to the threshold
    env.close()
    e32.close()
    # the single-env drop-in (NumPy float64 API) uses the float64 outputs
    one = gpr.BenchmarkPlanningEnv(layout_tiles=np.ones((4, 4)), num_movers=2, std_noise=0.0)
    o, i = one.reset(seed=1, options={'mover_start_xy_pos': start[0], 'mover_goal_xy_pos': goal[0]})
    o, r1, t1, _, i = one.step(np.zeros(4))
    assert o['achieved_goal'].dtype == np.float64 and np.array_equal(o['achieved_goal'], ora.achieved_goal[0])
    assert one.compute_reward(o['achieved_goal'], o['desired_goal'], i) == r1 == ora.reward[0]
    one.close()


def test_step_host_matches_device_step():
    """gpr_step_host (NumPy in / NumPy out) is the same computation as gpr_step on device tensors."""
    rng = np.random.default_rng(17)
    kw = dict(layout_tiles=np.ones((3, 3)), num_movers=4, std_noise=1e-5, seed=5)
    e1 = gpr.BenchmarkPlanningVecEnv(2048, device=DEV, **kw)
    e2 = gpr.BenchmarkPlanningVecEnv(2048, device=DEV, **kw)
    e1.reset(seed=5)
    e2.reset(seed=5)
    for _ in range(8):
        a = rng.uniform(-10, 10, (2048, 8)).astype(np.float32)
        o1, r1, t1, tr1, i1 = e1.step(torch.as_tensor(a, device=DEV))
        o2, r2, t2, tr2, i2 = e2.step_host(a)
        torch.cuda.synchronize()
        assert np.array_equal(o1['observation'].cpu().numpy(), o2['observation'])
        assert np.array_equal(o1['achieved_goal'].cpu().numpy(), o2['achieved_goal'])
        assert np.array_equal(r1.cpu().numpy(), r2) and np.array_equal(t1.cpu().numpy(), t2)
        assert np.array_equal(i1['wall_collision'].cpu().numpy(), i2['wall_collision'])
    e1.close()
    e2.close()


def test_host_calls_are_ordered_after_caller_stream_work():
    """gpr_step_host runs on the handle's private non-blocking stream.  A large gpr_reset / gpr_set_state on the
    caller's stream immediately followed by step_host — no synchronize in between — must still see the reset state."""
    B = 262144
    kw = dict(layout_tiles=np.ones((3, 3)), num_movers=4, std_noise=1e-5, seed=5)
    env, ora = make_pair(B, **kw)
    rng = np.random.default_rng(23)
    a0 = rng.uniform(-10, 10, (B, 8)).astype(np.float32)
    a1 = rng.uniform(-10, 10, (B, 8)).astype(np.float32)
    env.reset(seed=1)
    env.step_host(a0)  # (creates the private stream)
    side = torch.cuda.Stream(device=DEV)
    with torch.cuda.stream(side):
        for _ in range(4):  # keep the caller's stream busy: the reset below queues behind these
            torch.empty(1 << 28, dtype=torch.uint8, device=DEV).zero_()
        env.reset(seed=5)  # a full re-sample of 262,144 envs on a caller stream ...
    out = env.step_host(a1)  # ... immediately followed by a host-buffer step
    ora.reset(seed=5)
    ora.step(a1)
    assert np.array_equal(out[0]['achieved_goal'], ora.achieved_goal.astype(np.float32))
    assert np.array_equal(out[1], ora.reward.astype(np.float32)) and np.array_equal(out[2], ora.terminated.astype(bool))
    # same for gpr_set_state
    st = {k: v.clone() for k, v in env.get_state().items()}
    torch.cuda.synchronize()
    env.step_host(a0)
    with torch.cuda.stream(side):
        torch.empty(1 << 28, dtype=torch.uint8, device=DEV).zero_()
        env.set_state(st)
    o2 = env.step_host(a0)
    r_first = o2[1].copy()
    env.set_state(st)
    torch.cuda.synchronize()
    o3 = env.step_host(a0)
    assert np.array_equal(r_first, o3[1])
    env.close()


def test_step_host_copy_engine_route_matches(tmp_path):
    """GPR_HOST_IO=dma (results through device staging + copy engine instead of zero-copy stores; read once per process,
    hence the subprocess) gives the same results as the default route — with the COMPACT transport of the sparse rows
    (final_*, new desired_goal rows as lists scattered by the library; the default of the copy-engine route), with the dense
    transport (GPR_HOST_COMPACT=0), with float64 outputs, 1-8 movers, and across switches between device-pointer steps,
    host steps and resets (the library must notice when a host buffer no longer holds every env's goal)."""
    import subprocess
    import sys

    code = r"""
import sys, numpy as np, torch
sys.path[:0] = [%r, %r]
import os, gymnasium_planar_robotics_b200 as gpr
MODE = (os.environ['GPR_HOST_IO'], os.environ['GPR_HOST_COMPACT'])
for N, layout, f64 in ((4, (3, 3), False), (2, (3, 3), True), (8, (5, 5), False), (3, (4, 4), False)):
    kw = dict(layout_tiles=np.ones(layout), num_movers=N, std_noise=1e-5, seed=5, float64_outputs=f64, learn_jerk=(N == 8))
    B = 3001
    e1 = gpr.BenchmarkPlanningVecEnv(B, device='cuda:0', **kw)
    e2 = gpr.BenchmarkPlanningVecEnv(B, device='cuda:0', **kw)
    e1.reset(seed=5); e2.reset(seed=5)
    rng = np.random.default_rng(17)
    lim = 100.0 if N == 8 else 10.0
    finished = lists = 0
    for t in range(14):
        if t == 7:           # a masked reset through the device API: new goals the host buffer has not seen
            m = torch.as_tensor(rng.random(B) < 0.3, device='cuda:0')
            e1.reset(options={'mask': m}); e2.reset(options={'mask': m})
        a = rng.uniform(-lim, lim, (B, 2 * N)).astype(np.float32)
        o1, r1, t1, tr1, i1 = e1.step(torch.as_tensor(a, device='cuda:0'))
        if t in (5, 9):      # a device-pointer step in between: the host buffer misses the goals of the envs it resets
            o2d, r2d, *_ = e2.step(torch.as_tensor(a, device='cuda:0'))
            assert torch.equal(o2d['desired_goal'], o1['desired_goal']) and torch.equal(r2d, r1)
            continue
        o2, r2, t2, tr2, i2 = e2.step_host(a)
        torch.cuda.synchronize()
        for k in ('observation', 'achieved_goal', 'desired_goal'):
            assert np.array_equal(o1[k].cpu().numpy(), o2[k]), (N, t, k)
        assert np.array_equal(r1.cpu().numpy(), r2) and np.array_equal(t1.cpu().numpy(), t2) and np.array_equal(tr1.cpu().numpy(), tr2)
        for k in ('is_success', 'mover_collision', 'wall_collision'):
            assert np.array_equal(i1[k].cpu().numpy(), i2[k]), k
        d = t2 | tr2
        finished += int(d.sum())
        lists += 'index' in i2['final_obs']
        if 'index' in i2['final_obs']:
            assert sorted(i2['final_obs']['index'].tolist()) == np.flatnonzero(d).tolist()  # exactly the finished envs, once each
        fo = e2.final_obs_dense(i2)
        for k in ('observation', 'achieved_goal', 'desired_goal'):
            assert np.array_equal(i1['final_obs'][k].cpu().numpy()[d], fo[k][d]), (N, t, 'final ' + k)
    assert finished > B // 2
    assert (lists > 0) == (MODE == ('dma', '1')), (lists, MODE)  # the list form belongs to the compact transport only
    e1.close(); e2.close()
print('ok')
""" % (ROOT, ROOT + '/oracle')
    import os

    for mode, compact in (('dma', '1'), ('dma', '0'), ('zerocopy', '1')):
        env = dict(os.environ, GPR_HOST_IO=mode, GPR_HOST_COMPACT=compact)
        out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=600)
        assert out.returncode == 0 and 'ok' in out.stdout, (mode, compact, out.stdout[-500:], out.stderr[-1500:])


def test_desired_goal_on_change_survives_route_switches():
    """GPR_OUT_GOAL_ON_CHANGE (default in the env classes): steps rewrite desired_goal rows only for envs that were reset.
    Switching between the device-tensor route and the host-array route, masked resets and set_state must never leave a
    stale row in the buffer that is returned."""
    rng = np.random.default_rng(23)
    B = 2048
    env = gpr.BenchmarkPlanningVecEnv(B, device=DEV, layout_tiles=np.ones((3, 3)), num_movers=4, std_noise=0.0, seed=9, max_episode_steps=5)
    env.reset(seed=9)

    def goal_now():
        return env.get_state()['goal'].cpu().numpy().reshape(B, 8).astype(np.float32)

    for t in range(14):
        a = rng.uniform(-10, 10, (B, 8)).astype(np.float32)
        if t % 3 == 2:
            obs = env.step_host(a)[0]
            got = obs['desired_goal']
        else:
            obs = env.step(torch.as_tensor(a, device=DEV))[0]
            torch.cuda.synchronize()
            got = obs['desired_goal'].cpu().numpy()
        assert np.array_equal(got, goal_now()), f'step {t}'
        if t == 5:  # masked reset through the device route, then a host-route step
            mask = rng.random(B) < 0.3
            env.reset(options={'mask': mask})
        if t == 9:  # goals replaced behind the env's back
            st = env.get_state()
            env.set_state({'goal': st['goal'].flip(0)})
    env.close()


@pytest.mark.parametrize('mode', ['same_step', 'next_step'])
def test_checkpoint_resume_is_bit_exact(mode):
    """state_dict() / load_state_dict(): a second env built with the same kwargs continues an interrupted run with
    identical results (state, RNG event counters, RNG key after reset(seed), TimeLimit counters, NEXT_STEP bookkeeping)."""
    rng = np.random.default_rng(29)
    kw = dict(layout_tiles=np.ones((3, 3)), num_movers=4, std_noise=1e-5, seed=1, autoreset_mode=mode, max_episode_steps=6)
    a_env = gpr.BenchmarkPlanningVecEnv(3000, device=DEV, **kw)
    a_env.reset(seed=77)  # (not the constructor's seed: the checkpoint must carry the key in use)
    acts = [torch.as_tensor(rng.uniform(-10, 10, (3000, 8)).astype(np.float32), device=DEV) for _ in range(16)]
    for t in range(8):
        a_env.step(acts[t])
    ckpt = a_env.state_dict()
    b_env = gpr.BenchmarkPlanningVecEnv(3000, device=DEV, **kw)
    b_env.load_state_dict(ckpt)
    for t in range(8, 16):
        oa, ra, ta, tra, ia = a_env.step(acts[t])
        ob, rb, tb, trb, ib = b_env.step(acts[t])
        assert torch.equal(oa['observation'], ob['observation']) and torch.equal(oa['achieved_goal'], ob['achieved_goal'])
        assert torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(tra, trb) and torch.equal(ia['is_success'], ib['is_success'])
    sa, sb = a_env.get_state(), b_env.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert a_env.episode_stats()['episodes'] > 0
    a_env.close()
    b_env.close()


def test_sharding_does_not_change_results():
    """Multi-GPU rule: the RNG is keyed by the GLOBAL env index, so one handle with 4096 envs and two
    handles with 2048 each (env_index_base 0 / 2048) produce identical trajectories."""
    rng = np.random.default_rng(18)
    kw = dict(layout_tiles=np.ones((3, 3)), num_movers=4, std_noise=1e-5, seed=21)
    whole = gpr.BenchmarkPlanningVecEnv(4096, device=DEV, **kw)
    parts = [gpr.BenchmarkPlanningVecEnv(0, device=DEV, total_envs=4096, rank=r, world_size=2, **kw) for r in range(2)]
    whole.reset(seed=21)
    for p in parts:
        p.reset(seed=21)
    for _ in range(15):
        a = torch.as_tensor(rng.uniform(-10, 10, (4096, 8)).astype(np.float32), device=DEV)
        o, r, t, tr, _ = whole.step(a)
        outs = [p.step(a[i * 2048:(i + 1) * 2048].contiguous()) for i, p in enumerate(parts)]
        assert torch.equal(o['observation'], torch.cat([x[0]['observation'] for x in outs]))
        assert torch.equal(o['desired_goal'], torch.cat([x[0]['desired_goal'] for x in outs]))
        assert torch.equal(r, torch.cat([x[1] for x in outs])) and torch.equal(t, torch.cat([x[2] for x in outs]))
    whole.close()
    for p in parts:
        p.close()


def test_full_size_properties():
    """BASELINE configs[1] at full size (65,536 envs, 4 movers, 3x3): size-independent properties over 60 steps with
    auto-reset — kinematic limits, flag/reward consistency, spawn validity, determinism, episode accounting."""
    B = 65536
    kw = dict(layout_tiles=np.ones((3, 3)), num_movers=4, std_noise=1e-5, seed=1234)
    env = gpr.BenchmarkPlanningVecEnv(B, device=DEV, **kw)
    env2 = gpr.BenchmarkPlanningVecEnv(B, device=DEV, **kw)
    g = torch.Generator(device=DEV).manual_seed(0)
    env.reset(seed=1234)
    env2.reset(seed=1234)
    finished = 0
    for t in range(60):
        a = (torch.rand((B, 8), device=DEV, generator=g) * 2 - 1) * 10
        obs, r, term, trunc, info = env.step(a)
        o2, r2, term2, _, _ = env2.step(a)
        assert torch.equal(r, r2) and torch.equal(term, term2) and torch.equal(obs['achieved_goal'], o2['achieved_goal'])
        coll = info['mover_collision'] | info['wall_collision']
        assert torch.equal(term, coll | info['is_success'])                     # planning:477-478
        assert torch.equal(r == -50, coll) and torch.equal(r == 50, info['is_success'])
        assert ((r[~term] <= -1) & (r[~term] >= -4)).all()                     # -(N - reached), at least one missing
        st = env.get_state()
        assert (st["vel"].norm(dim=-1) <= 2.0 + 1e-4).all()  # plan:437 ||v_measured|| <= v_max; the true v may exceed it by the sensor noise (<= 5.77 sigma)
        assert (st['acc'].abs() <= 10.0).all()                                  # clipped action
        # envs that were just re-sampled: starts inside the spawn box, pairwise separation > 2r, zero velocity
        done = term | trunc
        finished += int(done.sum())
        if done.any():
            p = st['pos'][done]
            assert (p >= 0.11).all() and (p <= 0.55).all()                      # planning:262-267 quirk: max is 0.55
            d = torch.cdist(p, p) + torch.eye(4, device=DEV, dtype=torch.float64) * 10
            assert (d > 0.22).all()
            assert (st['vel'][done] == 0).all() and (st['elapsed_steps'][done] == 0).all()
            gl = st['goal'][done]
            dg = torch.cdist(gl, gl) + torch.eye(4, device=DEV, dtype=torch.float64) * 10
            assert (dg >= 0.22).all()                                           # planning:410
    stats = env.episode_stats(reset=True)
    assert stats['episodes'] == finished and finished > B // 2
    assert 1.0 <= stats['mean_length'] <= 50.0
    assert env.core.reset_failures() == 0
    env.close()
    env2.close()


@pytest.mark.parametrize('config', ['configs1_planning4', 'configs3_planning8box', 'configs2_pushing'])
def test_oracle_lockstep_at_baseline_sizes(config):
    """the oracle in lock-step with CUDA at BASELINE's FULL sizes — configs[1] 65,536 envs x 50 steps,
    configs[3] 262,144 x 10, configs[2] 65,536 x 20 — reference-default kwargs (std_noise = 1e-5), SAME_STEP auto-reset with
    on-device sampling, uniform random actions: flags, rewards and the float64 state must be bit-identical every step."""
    if config == 'configs1_planning4':
        B, steps, kw, cls, cfn = 65536, 50, dict(layout_tiles=np.ones((3, 3)), num_movers=4), gpr.BenchmarkPlanningVecEnv, gpr.planning_config
    elif config == 'configs3_planning8box':
        B, steps, cls, cfn = 262144, 10, gpr.BenchmarkPlanningVecEnv, gpr.planning_config
        kw = dict(layout_tiles=np.ones((5, 5)), num_movers=8, learn_jerk=True, collision_params={'shape': 'box', 'size': np.array([0.08, 0.08])})
    else:
        B, steps, kw, cls, cfn = 65536, 20, dict(), gpr.BenchmarkPushingVecEnv, gpr.pushing_config
    env = cls(B, device=DEV, seed=77, **kw)
    cfg, _ = cfn(num_envs=B, seed=77, **kw)
    ora = oracle.OracleEnv(cfg, nthreads=oracle.max_threads())
    env.reset(seed=77)
    ora.reset(seed=77)
    lim = cfg.j_max if cfg.learn_jerk else cfg.a_max
    rng = np.random.default_rng(5)
    keys = ('pos', 'vel', 'acc', 'goal') + (('act', 'mover_rot', 'object_pos', 'object_vel') if config == 'configs2_pushing' else ())
    resets = 0
    for t in range(steps):
        a = rng.uniform(-lim, lim, (B, ora.action_dim)).astype(np.float32)
        obs, r, term, trunc, info = env.step(torch.as_tensor(a, device=DEV))
        ora.step(a)
        torch.cuda.synchronize()
        for k in ('terminated', 'truncated', 'is_success', 'mover_collision', 'wall_collision'):
            assert np.array_equal(env.core.buf[k].cpu().numpy(), getattr(ora, k)), f'step {t}: {k}'
        assert np.array_equal(r.cpu().numpy(), ora.reward.astype(np.float32)), f'step {t}: reward'
        assert np.array_equal(obs['achieved_goal'].cpu().numpy(), ora.achieved_goal.astype(np.float32)), f'step {t}: achieved_goal'
        resets += int(ora.terminated.sum() + ora.truncated.sum())
        if t % 5 == 4 or t == steps - 1:
            st = env.get_state()
            torch.cuda.synchronize()
            for k in keys:
                assert np.array_equal(st[k].cpu().numpy(), getattr(ora, k)), f'step {t}: state {k}'
            assert np.array_equal(st['rng_counter'].cpu().numpy().view(np.uint32), ora.rng_counter)
    assert resets > B // 4  # the on-device rejection sampling really ran at this size
    assert config == 'configs2_pushing' or env.core.reset_failures() == 0
    env.close()


def test_full_size_pettingzoo_eight_movers_box_jerk():
    """BASELINE configs[3] at full size: PettingZoo-parallel form, 8 movers, box collision shape, learn_jerk=True,
    262,144 envs (5x5 tiles, see bench.py).  Size-independent properties over 12 steps with auto-reset, and the agent
    views must be exactly the columns of the vector env's tensors."""
    B, N = 262144, 8
    kw = dict(learn_jerk=True, collision_params={'shape': 'box', 'size': np.array([0.08, 0.08])}, seed=4321)
    pz = gpr.BenchmarkPlanningParallelEnv(B, np.ones((5, 5)), N, device=DEV, **kw)
    twin = gpr.BenchmarkPlanningVecEnv(B, np.ones((5, 5)), N, device=DEV, **kw)
    obs, infos = pz.reset(seed=4321)
    twin.reset(seed=4321)
    assert pz.agents == [f'mover_{i}' for i in range(N)] and obs['mover_7']['observation'].shape == (B, 4)
    g = torch.Generator(device=DEV).manual_seed(1)
    finished = 0
    for t in range(12):
        a = (torch.rand((B, N, 2), device=DEV, generator=g) * 2 - 1) * 100
        obs, rew, term, trunc, infos = pz.step({f'mover_{i}': a[:, i] for i in range(N)})
        o2, r2, t2, tr2, i2 = twin.step(a.reshape(B, 2 * N))
        assert torch.equal(rew['mover_0'], r2) and torch.equal(term['mover_3'], t2)          # same env, re-keyed
        full = o2['observation']                                                                # [vel(8x2), acc(8x2)]
        for i in (0, 5, 7):
            ob = obs[f'mover_{i}']
            assert torch.equal(ob['observation'][:, :2], full[:, 2 * i:2 * i + 2])
            assert torch.equal(ob['observation'][:, 2:], full[:, 2 * N + 2 * i:2 * N + 2 * i + 2])
            assert torch.equal(ob['achieved_goal'], o2['achieved_goal'][:, 2 * i:2 * i + 2])
        info = infos['mover_0']
        coll = info['mover_collision'] | info['wall_collision']
        assert torch.equal(t2, coll | info['is_success']) and torch.equal(r2 == -50, coll)
        st = twin.get_state()
        assert (st['vel'].norm(dim=-1) <= 2.0 + 1e-4).all() and (st['acc'].norm(dim=-1) <= 10.0 * (1 + 1e-9)).all()
        done = t2 | tr2
        finished += int(done.sum())
        if done.any():
            p = st['pos'][done]
            lo, hi = twin.cfg.min_xy_pos[0], twin.cfg.max_xy_pos[0]
            assert abs(lo - 0.08) < 1e-12 and abs(hi - (1.08 + 0.06 - 0.08)) < 1e-12  # planning:262-267 with the box margin
            assert (p >= lo).all() and (p <= hi).all()
            d = (p[:, :, None, :] - p[:, None, :, :]).abs()
            apart = (d[..., 0] > 0.16) | (d[..., 1] > 0.16) | torch.eye(N, device=DEV, dtype=torch.bool)
            assert apart.all()                                      # fresh starts: no two 0.16 m boxes overlap
    assert finished > B // 2 and twin.core.reset_failures() == 0
    pz.close()
    twin.close()


def test_single_env_and_pettingzoo_forms():
    """The reference's single-env call signatures (NumPy float64) and the PettingZoo-parallel re-keying."""
    env = gpr.BenchmarkPlanningEnv(layout_tiles=np.ones((3, 3)), num_movers=2, show_2D_plot=False, render_mode=None, std_noise=0.0)
    obs, info = env.reset(seed=0)
    assert set(obs) == {'observation', 'achieved_goal', 'desired_goal'} and obs['observation'].shape == (4,)
    assert obs['achieved_goal'].dtype == np.float64 and set(info) == {'is_success', 'mover_collision', 'wall_collision'}
    o, r, term, trunc, info = env.step(np.array([1.0, 0.0, 0.0, -1.0]))
    assert isinstance(r, float) and isinstance(term, bool) and trunc is False
    # closed form after one env-step of 40 cycles from rest: v = 40*dt*a, p = p0 + dt^2*a*(1+...+40)
    assert np.allclose(o['observation'], [0.04, 0, 0, -0.04], atol=1e-9)
    assert np.allclose(o['achieved_goal'][0] - obs['achieved_goal'][0], 1e-6 * 820, atol=1e-7)  # (float32 outputs)
    assert env.compute_reward(o['achieved_goal'], o['desired_goal'], info) == r
    env.close()

    pz = gpr.BenchmarkPlanningParallelEnv(64, np.ones((3, 3)), 3, device=DEV, std_noise=0.0, learn_jerk=True)
    obs, infos = pz.reset(seed=1)
    assert pz.agents == ['mover_0', 'mover_1', 'mover_2'] and obs['mover_1']['observation'].shape == (64, 4)
    acts = {a: torch.full((64, 2), 50.0 * (i + 1), device=DEV) for i, a in enumerate(pz.agents)}
    obs, rew, term, trunc, infos = pz.step(acts)
    full = pz._vec.core.buf['observation']
    assert torch.equal(obs['mover_2']['observation'][:, :2], full[:, 4:6]) and torch.equal(obs['mover_2']['observation'][:, 2:], full[:, 10:12])
    alive = ~term['mover_0']
    assert torch.allclose(obs['mover_1']['observation'][alive, 2:], torch.full((int(alive.sum()), 2), 4.0, device=DEV), atol=1e-5)
    pz.close()


@pytest.mark.parametrize('mode', ['same_step', 'next_step'])
@pytest.mark.parametrize('movers,num_envs', [(4, 65536), (8, 20000), (2, 262144 + 7)])
def test_streamed_autoreset_equals_serial_launches(mode, movers, num_envs):
    """gpr_step launches the auto-reset kernel as the step kernel's programmatic dependent: it consumes the work list while
    the step kernel's last wave is still publishing into it.  With per-kernel timing switched on the two kernels run one
    after the other (an event sits between them).  Both orders must give identical results, step after step, at batch
    sizes of several waves (random actions: a third of the envs is reset in every step)."""
    kw = dict(layout_tiles=np.ones((5, 5)) if movers == 8 else np.ones((3, 3)), num_movers=movers, std_noise=1e-5, seed=5,
              autoreset_mode=mode, max_episode_steps=7)
    a_env = gpr.BenchmarkPlanningVecEnv(num_envs, device=DEV, **kw)
    b_env = gpr.BenchmarkPlanningVecEnv(num_envs, device=DEV, **kw)
    b_env.core.kernel_times(True)  # serial launches
    a_env.reset(seed=5)
    b_env.reset(seed=5)
    gen = torch.Generator(device=DEV).manual_seed(3)
    for t in range(25):
        act = (torch.rand((num_envs, 2 * movers), device=DEV, generator=gen) * 2 - 1) * 10
        oa, ra, ta, tra, ia = a_env.step(act)
        ob, rb, tb, trb, ib = b_env.step(act)
        for k in ('observation', 'achieved_goal', 'desired_goal'):
            assert torch.equal(oa[k], ob[k]), (t, k)
        assert torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(tra, trb), t
        for k in ('is_success', 'mover_collision', 'wall_collision'):
            assert torch.equal(ia[k], ib[k]), (t, k)
    sa, sb = a_env.get_state(), b_env.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert a_env.core.reset_failures() == 0 and b_env.core.reset_failures() == 0
    assert a_env.episode_stats()['episodes'] > num_envs  # the work list really was busy
    b_env.core.kernel_times(False)
    a_env.close()
    b_env.close()


def test_debug_view_draws_the_chosen_env():
    """env.debug_view(i): the picture is made from the state of env i alone (movers where get_state says they are)."""
    from gymnasium_planar_robotics_b200 import debug_view as dv

    env = gpr.BenchmarkPlanningVecEnv(300, layout_tiles=np.ones((3, 3)), num_movers=3, device=DEV, seed=2)
    env.reset(seed=2)
    st = env.get_state()
    ppm = 400.0
    for i in (0, 123, 299):
        img = env.debug_view(i, ppm=ppm)
        assert img.shape == (288, 288, 3) and img.dtype == np.uint8
        pos = st['pos'][i].cpu().numpy()
        for m in range(3):
            # a point inside mover m's body, off the velocity arrow and the outlines
            x, y = pos[m, 0] + 0.04, pos[m, 1] - 0.04
            assert tuple(int(c) for c in img[img.shape[0] - 1 - int(y * ppm), int(x * ppm)]) == dv.PALETTE[m]
    push = gpr.BenchmarkPushingVecEnv(50, device=DEV, seed=3)
    push.reset(seed=3)
    img = push.debug_view(7, ppm=ppm)
    ob = push.get_state()['object_pos'][7].cpu().numpy()
    assert tuple(int(c) for c in img[img.shape[0] - 1 - int(ob[1] * ppm), int(ob[0] * ppm)]) == dv.OBJECT_COLOUR
    single = gpr.BenchmarkPlanningEnv(layout_tiles=np.ones((3, 3)), num_movers=2)
    single.reset(seed=1)
    assert single.debug_view().shape == (288, 288, 3) and single.render() is None
    for e in (env, push, single):
        e.close()


# ---------------------------------------------------------------------------------------------------------------------
# static obstacles (rules fixed on the oracle in tests/test_oracle_obstacles.py)
OBST_CIRCLE = np.array([[0.36, 0.36, 0.05], [0.62, 0.2, 0.03], [0.15, 0.8, 0.04]])
OBST_BOX = np.array([[0.36, 0.36, 0.05, 0.08], [0.62, 0.2, 0.03, 0.03], [0.15, 0.8, 0.06, 0.02]])


@pytest.mark.parametrize('noise', [0.0, 1e-5], ids=['exact', 'noisy'])
@pytest.mark.parametrize('shape,movers,jerk', [('circle', 1, False), ('circle', 3, False), ('circle', 4, True), ('box', 2, False), ('box', 8, True)])
def test_obstacles_injected_starts(shape, movers, jerk, noise):
    """Movers started anywhere (also on top of obstacles): other_collision, the loop break, reward and state must equal
    the oracle's bit for bit, with and without sensor noise, for lane groups of 1 to 8."""
    rng = np.random.default_rng(300 + movers + (7 if noise else 0))
    cp = {'shape': 'circle', 'size': 0.08} if shape == 'circle' else {'shape': 'box', 'size': np.array([0.07, 0.05])}
    env, ora = make_pair(1500, layout_tiles=np.ones((4, 4)), num_movers=movers, std_noise=noise, learn_jerk=jerk, collision_params=cp,
                         obstacles=OBST_CIRCLE if shape == 'circle' else OBST_BOX, autoreset_mode='off', max_episode_steps=50)
    ev = run_lockstep(env, ora, 25, rng, 130.0 if jerk else 13.0)
    assert ev > 0
    assert ora.other_collision.any()
    env.close()


@pytest.mark.parametrize('mode', ['same_step', 'next_step'])
@pytest.mark.parametrize('shape,movers', [('circle', 2), ('circle', 4), ('box', 3), ('box', 8)])
def test_obstacles_with_autoreset_sampling(shape, movers, mode):
    """Sampled starts / goals avoid the obstacles (rejection loops of both samplers), auto-reset in both modes."""
    rng = np.random.default_rng(400 + movers)
    cp = {'shape': 'circle', 'size': 0.08, 'offset': 0.005} if shape == 'circle' else {'shape': 'box', 'size': np.array([0.07, 0.05]), 'offset': 0.005}
    layout = np.ones((5, 5)) if movers == 8 else np.ones((4, 4))
    env, ora = make_pair(1200, layout_tiles=layout, num_movers=movers, std_noise=1e-5, collision_params=cp,
                         obstacles=OBST_CIRCLE if shape == 'circle' else OBST_BOX, autoreset_mode=mode, max_episode_steps=6, seed=11)
    run_lockstep(env, ora, 30, rng, 12.0, inject=False, seed=11)
    assert env.core.reset_failures() == 0 and not ora.reset_failed.any()
    st = env.get_state()
    for arr in (st['pos'].cpu().numpy(), st['goal'].cpu().numpy()):
        assert np.isfinite(arr).all()
    info = env._info(False)
    assert 'other_collision' in info and info['other_collision'].dtype == torch.bool
    env.close()


@pytest.mark.parametrize('shape,movers,mode', [('circle', 2, 'same_step'), ('circle', 4, 'next_step'), ('box', 3, 'same_step'), ('box', 8, 'off')])
def test_extra_bodies_with_prescribed_velocity(shape, movers, mode):
    """typed extra bodies — kinematic circles / boxes with a constant velocity (gpr_config.obstacle_vel).
    CUDA == oracle bit for bit, including the travel-budget bookkeeping of the step kernel (a moving body uses up the
    certified clearance on its own account)."""
    rng = np.random.default_rng(500 + movers)
    if shape == 'circle':
        cp = {'shape': 'circle', 'size': 0.08}
        bodies = [{'pos': (0.48, 0.48), 'size': 0.05}, {'pos': (0.1, 0.8), 'size': 0.04, 'vel': (0.4, -0.05)}, {'pos': (0.9, 0.2), 'size': 0.03, 'vel': (-0.5, 0.3)}]
    else:
        cp = {'shape': 'box', 'size': np.array([0.07, 0.05])}
        bodies = [{'shape': 'box', 'pos': (0.5, 0.5), 'size': (0.06, 0.03)}, {'shape': 'box', 'pos': (0.1, 0.75), 'size': (0.03, 0.08), 'vel': (0.45, 0.0)},
                  {'shape': 'box', 'pos': (0.95, 0.3), 'size': (0.05, 0.02), 'vel': (-0.3, 0.2)}]
    layout = np.ones((5, 5)) if movers == 8 else np.ones((4, 4))
    env, ora = make_pair(1500, layout_tiles=layout, num_movers=movers, std_noise=1e-5, collision_params=cp, extra_bodies=bodies,
                         autoreset_mode=mode, max_episode_steps=12, seed=9)
    run_lockstep(env, ora, 30, rng, 4.0, inject=(mode == 'off'), seed=9)
    assert ora.other_collision.any() or mode != 'off'
    env.close()


def test_custom_env_without_xml_matches_the_oracle():
    """The docs/make_own_env.rst-style custom env of tests/test_oracle_obstacles.py (robot base + conveyor pallets, no XML)."""
    from test_oracle_obstacles import custom_conveyor_env_kwargs

    env, ora = make_pair(4096, seed=3, **custom_conveyor_env_kwargs())
    run_lockstep(env, ora, 40, np.random.default_rng(1), 6.0, inject=False, seed=3)
    assert env.episode_stats()['episodes'] > 1000
    env.close()


def test_obstacles_through_the_host_path_and_her():
    """step_host delivers other_collision like the other flags; compute_reward treats it as a collision."""
    cp = {'shape': 'circle', 'size': 0.08}
    kw = dict(layout_tiles=np.ones((4, 4)), num_movers=2, collision_params=cp, obstacles=OBST_CIRCLE, seed=5, std_noise=1e-5)
    a_env = gpr.BenchmarkPlanningVecEnv(3000, device=DEV, **kw)
    b_env = gpr.BenchmarkPlanningVecEnv(3000, device=DEV, **kw)
    a_env.reset(seed=5)
    b_env.reset(seed=5)
    rng = np.random.default_rng(6)
    seen = 0
    for _ in range(20):
        act = rng.uniform(-10, 10, (3000, 4)).astype(np.float32)
        obs, r, term, trunc, info = a_env.step(torch.as_tensor(act, device=DEV))
        hobs, hr, hterm, htrunc, hinfo = b_env.step_host(act)
        assert np.array_equal(info['other_collision'].cpu().numpy(), hinfo['other_collision'])
        assert np.array_equal(r.cpu().numpy(), hr) and np.array_equal(term.cpu().numpy(), hterm)
        oc = info['other_collision']
        seen += int(oc.sum())
        assert bool((r[oc] == -50.0).all()) and bool(term[oc].all())
        fin = info['final_obs']
        rr = a_env.compute_reward(fin['achieved_goal'][oc], fin['desired_goal'][oc], {k: info[k][oc] for k in ('mover_collision', 'wall_collision', 'other_collision')})
        assert bool((rr == -50.0).all())
    assert seen > 0
    a_env.close()
    b_env.close()
