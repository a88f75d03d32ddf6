"""Host-side packing of the reference's constructor kwargs into the POD ``gpr_config`` (include/gpr.h).

Everything here is the part of ``BasicPlanarRoboticsEnv.__init__`` (basic_envs.py:162-289),
``BenchmarkPlanningEnv.__init__`` (planning/benchmark_planning_env.py:165-291) and ``BenchmarkPushingEnv.__init__``
(manipulation/benchmark_pushing_env.py:154-288) that produces NUMBERS the step path uses.  Derived quantities are
evaluated in NumPy float64 with the reference's own expressions (same operands, same order) so that every threshold the
kernels compare against is bit-identical to the reference's.  MuJoCo XML generation, viewers and meshes are out of scope
(SURVEY.md §2 rows 12-17); asking for them raises.
"""

from __future__ import annotations

import ctypes
import warnings
from typing import Any

import numpy as np

GPR_ABI_VERSION = 3
GPR_MAX_MOVERS = 32
GPR_MAX_TILES_1D = 32
GPR_MAX_OBSTACLES = 8

ENV_PLANNING, ENV_PUSHING = 0, 1
SHAPE_CIRCLE, SHAPE_BOX = 0, 1
AUTORESET_OFF, AUTORESET_SAME_STEP, AUTORESET_NEXT_STEP = 0, 1, 2
_AUTORESET = {
    None: AUTORESET_OFF,
    'off': AUTORESET_OFF,
    'none': AUTORESET_OFF,
    'same_step': AUTORESET_SAME_STEP,
    'next_step': AUTORESET_NEXT_STEP,
}

_D2M = (ctypes.c_double * 2) * GPR_MAX_MOVERS


class GprConfig(ctypes.Structure):
    """ctypes mirror of ``struct gpr_config``; field order and types must match include/gpr.h exactly
    (checked at load time against ``gpr_config_bytes()``)."""

    _fields_ = [
        ('struct_bytes', ctypes.c_uint32),
        ('abi_version', ctypes.c_uint32),
        ('env_kind', ctypes.c_int32),
        ('num_envs', ctypes.c_int32),
        ('env_index_base', ctypes.c_int64),
        ('seed', ctypes.c_uint64),
        ('num_tiles_x', ctypes.c_int32),
        ('num_tiles_y', ctypes.c_int32),
        ('layout', ctypes.c_uint8 * (GPR_MAX_TILES_1D * GPR_MAX_TILES_1D)),
        ('tile_half', ctypes.c_double * 2),
        ('tile_cx', ctypes.c_double * GPR_MAX_TILES_1D),
        ('tile_cy', ctypes.c_double * GPR_MAX_TILES_1D),
        ('c_shape', ctypes.c_int32),
        ('reference_quirks', ctypes.c_int32),
        ('c_wall', _D2M * 2),
        ('c_mover', _D2M * 2),
        ('num_movers', ctypes.c_int32),
        ('learn_jerk', ctypes.c_int32),
        ('num_cycles', ctypes.c_int32),
        ('max_episode_steps', ctypes.c_int32),
        ('cycle_time', ctypes.c_double),
        ('v_max', ctypes.c_double),
        ('a_max', ctypes.c_double),
        ('j_max', ctypes.c_double),
        ('threshold_pos', ctypes.c_double),
        ('min_xy_pos', ctypes.c_double * 2),
        ('max_xy_pos', ctypes.c_double * 2),
        ('min_goal_dist', ctypes.c_double),
        ('std_noise', ctypes.c_double * 3),
        ('autoreset_mode', ctypes.c_int32),
        ('max_reset_attempts', ctypes.c_int32),
        ('object_min_xy_pos', ctypes.c_double * 2),
        ('object_max_xy_pos', ctypes.c_double * 2),
        ('min_mo_dist', ctypes.c_double),
        ('object_noise_xy', ctypes.c_double),
        ('object_half_xy', ctypes.c_double),
        ('object_mass', ctypes.c_double),
        ('object_damping', ctypes.c_double),
        ('mover_half', ctypes.c_double * 2),
        ('mover_mass', ctypes.c_double),
        ('imp_k_rot', ctypes.c_double),
        ('gravity', ctypes.c_double),
        ('friction', ctypes.c_double),
        ('solref', ctypes.c_double * 2),
        ('solimp', ctypes.c_double * 5),
        ('contact_iterations', ctypes.c_int32),
        ('output_flags', ctypes.c_int32),
        ('num_obstacles', ctypes.c_int32),
        ('contact_warm_start', ctypes.c_int32),
        ('obstacle_xy', (ctypes.c_double * 2) * GPR_MAX_OBSTACLES),
        ('obstacle_size', (ctypes.c_double * 2) * GPR_MAX_OBSTACLES),
        ('obstacle_vel', (ctypes.c_double * 2) * GPR_MAX_OBSTACLES),
    ]


class GprOutputs(ctypes.Structure):
    _fields_ = [
        ('observation', ctypes.c_void_p),
        ('achieved_goal', ctypes.c_void_p),
        ('desired_goal', ctypes.c_void_p),
        ('reward', ctypes.c_void_p),
        ('terminated', ctypes.c_void_p),
        ('truncated', ctypes.c_void_p),
        ('is_success', ctypes.c_void_p),
        ('mover_collision', ctypes.c_void_p),
        ('wall_collision', ctypes.c_void_p),
        ('final_observation', ctypes.c_void_p),
        ('final_achieved_goal', ctypes.c_void_p),
        ('final_desired_goal', ctypes.c_void_p),
        ('other_collision', ctypes.c_void_p),
        ('final_index', ctypes.c_void_p),
        ('final_count', ctypes.c_void_p),
    ]


class GprState(ctypes.Structure):
    _fields_ = [
        ('pos', ctypes.c_void_p),
        ('vel', ctypes.c_void_p),
        ('acc', ctypes.c_void_p),
        ('goal', ctypes.c_void_p),
        ('elapsed_steps', ctypes.c_void_p),
        ('rng_counter', ctypes.c_void_p),
        ('act', ctypes.c_void_p),
        ('mover_rot', ctypes.c_void_p),
        ('object_pos', ctypes.c_void_p),
        ('object_vel', ctypes.c_void_p),
        ('needs_reset', ctypes.c_void_p),
        ('episode_return', ctypes.c_void_p),
        ('contact_warm', ctypes.c_void_p),
    ]


# ----------------------------------------------------------------------------------------------------------------------
# reference-rule helpers
# ----------------------------------------------------------------------------------------------------------------------
def _tile_centres(num_tiles: int, tile_wl: float) -> np.ndarray:
    """basic_envs.py:1300-1303 (``get_1D_tile_pos``), verbatim expression."""
    return np.linspace(start=tile_wl / 2, stop=(num_tiles - 1) * tile_wl + (tile_wl / 2), num=num_tiles, endpoint=True)


def _reject_out_of_scope(mover_params: dict[str, Any], render_mode, show_2D_plot=False, use_mj_passive_viewer=False):
    if render_mode is not None:
        raise NotImplementedError(
            "render_mode must be None: rendering (MuJoCo viewers / Matplotlib2DViewer) is outside the B200 step path "
            '(SURVEY.md §2 row 12). Use the reference package to visualise a trajectory.'
        )
    if show_2D_plot or use_mj_passive_viewer:
        raise NotImplementedError('show_2D_plot / use_mj_passive_viewer are visual-only and not part of the step path.')
    shape = mover_params.get('shape', 'box')
    if shape == 'mesh' or (isinstance(shape, list) and 'mesh' in shape) or 'mesh' in mover_params:
        raise NotImplementedError("mesh movers are visual assets outside the step path (SURVEY.md §2 row 13).")


def _std_noise(std_noise) -> np.ndarray:
    # basic_envs.py:184-192
    if isinstance(std_noise, (float, int)):
        return np.array([float(std_noise)] * 3)
    std_noise = np.asarray(std_noise, dtype=np.float64)
    assert std_noise.shape == (3,), 'noise standard deviation has to be a float or a numpy array of shape (3,)'
    return std_noise


def _common(
    cfg: GprConfig,
    *,
    num_envs: int,
    layout_tiles: np.ndarray,
    num_movers: int,
    tile_params: dict | None,
    mover_params: dict | None,
    std_noise,
    num_cycles: int,
    collision_params: dict | None,
    v_max: float,
    a_max: float,
    j_max: float,
    learn_jerk: bool,
    threshold_pos: float,
    cycle_time: float,
    max_episode_steps: int,
    autoreset_mode,
    max_reset_attempts: int,
    env_index_base: int,
    seed: int,
    reference_quirks: bool,
) -> dict[str, Any]:
    """Fill the fields shared by both envs; returns the derived Python-side values the env classes also need."""
    if num_envs <= 0:
        raise ValueError('num_envs must be > 0')
    cfg.struct_bytes = ctypes.sizeof(GprConfig)
    cfg.abi_version = GPR_ABI_VERSION
    cfg.num_envs = int(num_envs)
    cfg.env_index_base = int(env_index_base)
    if not (0 <= env_index_base and env_index_base + num_envs <= 2**32):
        raise ValueError('global env indices must fit 32 bits (they are one Philox counter word)')
    cfg.seed = int(seed) & (2**64 - 1)

    # --- tiles: basic_envs.py:194-204, _check_tile_config 1469-1485
    layout_tiles = np.asarray(layout_tiles)
    assert len(layout_tiles.shape) == 2, 'Unexpected tile layout shape. Expected: (num_tiles_x,num_tiles_y)'
    layout = layout_tiles.astype(np.int8)
    assert np.bitwise_or(layout == 0, layout == 1).all(), 'Use a numpy array of only 0 and 1 to specify the tile layout.'
    assert np.sum(layout) > 0, 'Number of tiles must be >0.'
    nx, ny = layout.shape
    if nx > GPR_MAX_TILES_1D or ny > GPR_MAX_TILES_1D:
        raise ValueError(f'at most {GPR_MAX_TILES_1D} tiles per axis are supported')
    tile_params = {} if tile_params is None else tile_params
    tile_size = np.asarray(tile_params.get('size', np.array([0.24 / 2, 0.24 / 2, 0.0352 / 2])), dtype=np.float64)
    assert tile_size.shape == (3,), 'Specify the size of a tile using a numpy array of shape (3,)'
    assert (tile_size > 0).all(), 'Tile size must be >0.'
    cx = _tile_centres(nx, tile_size[0] * 2)
    cy = _tile_centres(ny, tile_size[1] * 2)
    cfg.num_tiles_x, cfg.num_tiles_y = nx, ny
    flat = layout.astype(np.uint8).reshape(-1)
    for k, val in enumerate(flat):
        cfg.layout[k] = int(val)
    cfg.tile_half[0], cfg.tile_half[1] = float(tile_size[0]), float(tile_size[1])
    for i in range(nx):
        cfg.tile_cx[i] = float(cx[i])
    for j in range(ny):
        cfg.tile_cy[j] = float(cy[j])

    # --- movers: basic_envs.py:224-255
    assert num_movers > 0, 'Number of movers must be >0.'
    if num_movers > GPR_MAX_MOVERS:
        raise ValueError(f'at most {GPR_MAX_MOVERS} movers are supported (one lane per mover, one warp per env)')
    mover_params = {} if mover_params is None else mover_params
    mover_size = np.asarray(mover_params.get('size', np.array([0.155 / 2, 0.155 / 2, 0.012 / 2])), dtype=np.float64)
    if mover_size.shape == (3,):
        mover_size = np.tile(mover_size, reps=(num_movers, 1))
    assert mover_size.shape == (num_movers, 3), 'Unexpected mover size.'
    assert (mover_size > 0).all(), 'Mover size must be >0.'
    mover_mass = mover_params.get('mass', 1.24)
    mover_shape = mover_params.get('shape', 'box')

    # --- collision: basic_envs.py:257-264, _check_collision_params 1556-1603
    collision_params = {} if collision_params is None else collision_params
    c_shape = collision_params.get('shape', 'circle')
    c_size = collision_params.get('size', 0.11)
    c_off = collision_params.get('offset', 0.0)
    c_off_wall = collision_params.get('offset_wall', 0.0)
    assert c_shape in ('circle', 'box'), 'Unexpected collision shape. You can choose between circle and box.'
    assert isinstance(c_off, float) and isinstance(c_off_wall, float), 'Use single float values to specify the offsets.'
    assert c_off >= 0 and c_off_wall >= 0, 'collision offsets must be >= 0.'
    per_mover = False
    if c_shape == 'circle':
        if isinstance(c_size, np.ndarray):
            assert c_size.shape == (num_movers,), 'circle: size must be a float or an array of shape (num_movers,)'
            per_mover = True
        else:
            c_size = float(c_size)
    else:
        assert not isinstance(c_size, float), 'box: use a numpy array of shape (2,) or (num_movers,2) for the size.'
        c_size = np.asarray(c_size, dtype=np.float64)
        assert c_size.shape == (2,) or c_size.shape == (num_movers, 2)
        per_mover = c_size.ndim == 2

    def c_size_arr(val) -> np.ndarray:
        # basic_envs.py:1209-1242 get_c_size_arr
        if isinstance(val, float):
            return np.tile(np.array([[val]]), reps=(num_movers, 1))
        if c_shape == 'circle':
            return val.reshape((num_movers, 1))
        if val.shape == (2,):
            return np.tile(val, reps=(num_movers, 1))
        return val.copy()

    for s in (0, 1):
        wall = c_size_arr(c_size + c_off_wall + int(bool(s)) * c_off)  # basic_envs.py:487
        mov = c_size_arr(c_size + c_off * int(bool(s)))  # basic_envs.py:390
        for m in range(num_movers):
            for k in range(2):
                cfg.c_wall[s][m][k] = float(wall[m, min(k, wall.shape[1] - 1)])
                cfg.c_mover[s][m][k] = float(mov[m, min(k, mov.shape[1] - 1)])
    cfg.c_shape = SHAPE_CIRCLE if c_shape == 'circle' else SHAPE_BOX
    cfg.reference_quirks = int(bool(reference_quirks))

    # guard the reference's `assert mask_valid in {0,1}` (basic_envs.py:650): a shape as wide as a tile makes both
    # sides of one axis unsafe at once.  The reference crashes there; refuse at construction instead.
    worst = max(cfg.c_wall[1][m][k] for m in range(num_movers) for k in range(2))
    if c_shape == 'circle' and worst >= min(tile_size[0], tile_size[1]):
        raise ValueError(
            'collision circle (size + offsets) must be smaller than half a tile; the reference trips the assertion at '
            'basic_envs.py:650 for such shapes'
        )

    # warnings the reference emits (basic_envs.py:1585-1598)
    for m in range(num_movers):
        if c_shape == 'circle':
            if cfg.c_mover[0][m][0] < np.sqrt(mover_size[m, 0] ** 2 + mover_size[m, 1] ** 2):
                warnings.warn(f'Mover {m} is not completely included in collision shape.', stacklevel=3)
        elif (np.array([cfg.c_mover[0][m][0], cfg.c_mover[0][m][1]]) < mover_size[m, :2]).any():
            warnings.warn(f'Mover {m} is not completely included in collision shape.', stacklevel=3)

    # --- dynamics
    cfg.num_movers = int(num_movers)
    cfg.learn_jerk = int(bool(learn_jerk))
    assert num_cycles > 0
    cfg.num_cycles = int(num_cycles)
    cfg.max_episode_steps = int(max_episode_steps)
    cfg.cycle_time = float(cycle_time)
    cfg.v_max, cfg.a_max, cfg.j_max = float(v_max), float(a_max), float(j_max)
    cfg.threshold_pos = float(threshold_pos)
    sn = _std_noise(std_noise)
    for k in range(3):
        cfg.std_noise[k] = float(sn[k])
    if isinstance(autoreset_mode, str) or autoreset_mode is None:
        autoreset_mode = _AUTORESET[autoreset_mode.lower() if isinstance(autoreset_mode, str) else None]
    cfg.autoreset_mode = int(autoreset_mode)
    cfg.max_reset_attempts = int(max_reset_attempts)

    # --- spawn box: planning:262-267 == pushing:250-255 (note: tile_size/2 on a HALF size is the reference's formula)
    safety_margin = c_size + c_off_wall + c_off
    min_xy_pos = np.zeros(2) + (safety_margin if not per_mover else np.max(np.asarray(safety_margin), axis=0))
    max_xy_pos = np.array([np.max(cx) + (tile_size[0] / 2), np.max(cy) + (tile_size[1] / 2)]) - (
        safety_margin if not per_mover else np.max(np.asarray(safety_margin), axis=0)
    )
    cfg.min_xy_pos[0], cfg.min_xy_pos[1] = float(min_xy_pos[0]), float(min_xy_pos[1])
    cfg.max_xy_pos[0], cfg.max_xy_pos[1] = float(max_xy_pos[0]), float(max_xy_pos[1])
    high_goals = np.array([np.max(cx) + (tile_size[0] / 2), np.max(cy) + (tile_size[1] / 2)])

    return {
        'tile_size': tile_size,
        'x_pos_tiles': cx,
        'y_pos_tiles': cy,
        'mover_size': mover_size,
        'mover_mass': mover_mass,
        'mover_shape': mover_shape,
        'c_shape': c_shape,
        'c_size': c_size,
        'c_size_offset': c_off,
        'c_size_offset_wall': c_off_wall,
        'min_xy_pos': min_xy_pos,
        'max_xy_pos': max_xy_pos,
        'high_goals': high_goals,
        'layout_tiles': layout,
        'per_mover_sizes': per_mover,
    }


def planning_config(
    *,
    num_envs: int,
    layout_tiles: np.ndarray,
    num_movers: int,
    show_2D_plot: bool = False,
    mover_colors_2D_plot=None,
    tile_params: dict | None = None,
    mover_params: dict | None = None,
    initial_mover_zpos: float = 0.003,
    std_noise=1e-5,
    render_mode: str | None = None,
    render_every_cycle: bool = False,
    num_cycles: int = 40,
    collision_params: dict | None = None,
    v_max: float = 2.0,
    a_max: float = 10.0,
    j_max: float = 100.0,
    learn_jerk: bool = False,
    threshold_pos: float = 0.1,
    use_mj_passive_viewer: bool = False,
    # --- additions of the batched simulator (not reference kwargs)
    cycle_time: float = 0.001,
    max_episode_steps: int = 50,
    autoreset_mode='same_step',
    max_reset_attempts: int = 100000,
    env_index_base: int = 0,
    seed: int = 0,
    reference_quirks: bool = False,
    goal_output_on_change: bool = True,
    float64_outputs: bool = False,
    obstacles=None,
    extra_bodies=None,
) -> tuple[GprConfig, dict[str, Any]]:
    """kwargs of ``BenchmarkPlanningEnv`` (planning:165-185) -> ``gpr_config``.

    ``obstacles``: static obstacles, the typed form of ``_check_for_other_collisions_callback`` (basic_envs.py:1976-1986):
    an array (K, 3) of ``[x, y, radius]`` rows for the circle collision shape or (K, 4) of ``[x, y, half_x, half_y]`` rows
    (axis-aligned) for the box shape, K <= 8; two more columns ``[..., vx, vy]`` give a body a prescribed constant
    velocity.  See ``gpr_config.num_obstacles`` / ``obstacle_vel`` in include/gpr.h for the rules.

    ``extra_bodies``: the same thing as a typed list — what a custom env of the reference would add to its MuJoCo model
    through ``custom_model_xml_strings`` (basic_envs.py:134-155) and check in its collision callback, here without XML:
    ``[{'shape': 'circle'|'box', 'pos': (x, y), 'size': r | (half_x, half_y), 'vel': (vx, vy)}, ...]`` (``vel`` optional:
    static body; the shape must be the env's collision shape).  Appended to ``obstacles``."""
    del mover_colors_2D_plot, render_every_cycle, initial_mover_zpos  # visual only / z is not simulated
    _reject_out_of_scope({} if mover_params is None else mover_params, render_mode, show_2D_plot, use_mj_passive_viewer)
    cfg = GprConfig()
    cfg.env_kind = ENV_PLANNING
    d = _common(
        cfg,
        num_envs=num_envs,
        layout_tiles=layout_tiles,
        num_movers=num_movers,
        tile_params=tile_params,
        mover_params=mover_params,
        std_noise=std_noise,
        num_cycles=num_cycles,
        collision_params=collision_params,
        v_max=v_max,
        a_max=a_max,
        j_max=j_max,
        learn_jerk=learn_jerk,
        threshold_pos=threshold_pos,
        cycle_time=cycle_time,
        max_episode_steps=max_episode_steps,
        autoreset_mode=autoreset_mode,
        max_reset_attempts=max_reset_attempts,
        env_index_base=env_index_base,
        seed=seed,
        reference_quirks=reference_quirks,
    )
    # planning:270-274 minimum distance between any two goals
    if d['c_shape'] == 'circle':
        min_goal_dist = 2 * (np.max(d['c_size']) + d['c_size_offset'])
    else:
        cs = np.asarray(d['c_size'])
        cs = cs if cs.ndim == 1 else np.max(cs, axis=0)
        min_goal_dist = 2 * np.linalg.norm(cs + d['c_size_offset'], ord=2)
    cfg.min_goal_dist = float(min_goal_dist)
    d['min_goal_dist'] = float(min_goal_dist)
    want = 3 if d['c_shape'] == 'circle' else 4
    obst = np.zeros((0, want + 2)) if obstacles is None else np.asarray(obstacles, dtype=np.float64)
    if obst.size:
        if obst.ndim != 2 or obst.shape[1] not in (want, want + 2):
            raise ValueError(f"obstacles must have shape (K, {want}) or (K, {want + 2}) for collision shape '{d['c_shape']}' "
                             f"([x, y, radius] / [x, y, half_x, half_y], optionally followed by [vx, vy]), got {obst.shape}")
        if obst.shape[1] == want:
            obst = np.concatenate([obst, np.zeros((obst.shape[0], 2))], axis=1)
    else:
        obst = np.zeros((0, want + 2))
    if extra_bodies:
        rows = []
        for body in extra_bodies:
            if body.get('shape', d['c_shape']) != d['c_shape']:
                raise NotImplementedError(f"an extra body must have the env's collision shape ('{d['c_shape']}'); mixed shapes have no "
                                          'rule in the reference to restate')
            size = np.atleast_1d(np.asarray(body['size'], dtype=np.float64))
            if size.shape != (want - 2,):
                raise ValueError(f"extra body size: expected {'a radius' if want == 3 else '(half_x, half_y)'}, got {body['size']!r}")
            rows.append(np.concatenate([np.asarray(body['pos'], dtype=np.float64).reshape(2), size,
                                        np.asarray(body.get('vel', (0.0, 0.0)), dtype=np.float64).reshape(2)]))
        obst = np.concatenate([obst, np.stack(rows)], axis=0)
    if obst.shape[0]:
        if obst.shape[0] > GPR_MAX_OBSTACLES:
            raise ValueError(f'at most {GPR_MAX_OBSTACLES} obstacles / extra bodies')
        if not (np.isfinite(obst).all() and (obst[:, 2:want] > 0).all()):
            raise ValueError('obstacle sizes must be finite and > 0')
        cfg.num_obstacles = int(obst.shape[0])
        for k in range(obst.shape[0]):
            cfg.obstacle_xy[k][0], cfg.obstacle_xy[k][1] = float(obst[k, 0]), float(obst[k, 1])
            cfg.obstacle_size[k][0] = float(obst[k, 2])
            cfg.obstacle_size[k][1] = float(obst[k, 3]) if want == 4 else float(obst[k, 2])
            cfg.obstacle_vel[k][0], cfg.obstacle_vel[k][1] = float(obst[k, want]), float(obst[k, want + 1])
    d['obstacles'] = obst
    d['obs_dim'] = num_movers * (1 + int(bool(learn_jerk))) * 2
    d['goal_dim'] = num_movers * 2
    d['action_dim'] = num_movers * 2
    # the env classes hand the same output buffers to every call: desired_goal rows are rewritten only when they change
    # float64_outputs (GPR_OUT_FLOAT64): observation / goal arrays in the reference's dtype (the single-env classes set it)
    cfg.output_flags = (1 if goal_output_on_change else 0) | (2 if float64_outputs else 0)
    return cfg, d


def pushing_config(
    *,
    num_envs: int,
    mover_params: dict | None = None,
    initial_mover_zpos: float = 0.003,
    std_noise=1e-5,
    render_mode: str | None = None,
    render_every_cycle: bool = False,
    num_cycles: int = 40,
    collision_params: dict | None = None,
    v_max: float = 2.0,
    a_max: float = 10.0,
    j_max: float = 100.0,
    learn_jerk: bool = False,
    threshold_pos: float = 0.05,
    use_mj_passive_viewer: bool = False,
    # --- additions of the batched simulator
    cycle_time: float = 0.001,
    max_episode_steps: int = 50,
    autoreset_mode='same_step',
    max_reset_attempts: int = 4096,
    env_index_base: int = 0,
    seed: int = 0,
    contact_iterations: int = 3,
    contact_warm_start: bool = True,
    goal_output_on_change: bool = True,
    float64_outputs: bool = False,
) -> tuple[GprConfig, dict[str, Any]]:
    """kwargs of ``BenchmarkPushingEnv`` (pushing:154-169) -> ``gpr_config``.

    ``max_reset_attempts`` caps the object-placement loop of pushing:392-407.  The reference loops without bound, and
    NEVER terminates when the mover is drawn within a few millimetres of the layout centre (every point of the object box
    [0.22, 0.44]^2 is then closer than ``min_mo_dist`` = 0.159 m; probability ~2e-4 per reset).  Here such a reset keeps
    its last draw and is counted in ``gpr_reset_failures``."""
    del render_every_cycle, initial_mover_zpos
    _reject_out_of_scope({} if mover_params is None else mover_params, render_mode, False, use_mj_passive_viewer)
    cfg = GprConfig()
    cfg.env_kind = ENV_PUSHING
    d = _common(
        cfg,
        num_envs=num_envs,
        layout_tiles=np.ones((3, 3)),  # pushing:195
        num_movers=1,  # pushing:196
        tile_params=None,
        mover_params=mover_params,
        std_noise=std_noise,
        num_cycles=num_cycles,
        collision_params=collision_params,
        v_max=v_max,
        a_max=a_max,
        j_max=j_max,
        learn_jerk=learn_jerk,
        threshold_pos=threshold_pos,
        cycle_time=cycle_time,
        max_episode_steps=max_episode_steps,
        autoreset_mode=autoreset_mode,
        max_reset_attempts=max_reset_attempts,
        env_index_base=env_index_base,
        seed=seed,
        reference_quirks=False,
    )
    object_length_xy = 0.07 / 2  # pushing:173
    safety_margin = d['c_size'] + d['c_size_offset_wall'] + d['c_size_offset']
    object_min = d['min_xy_pos'] + safety_margin  # pushing:257
    object_max = d['max_xy_pos'] - safety_margin  # pushing:258
    for k in range(2):
        cfg.object_min_xy_pos[k] = float(object_min[k])
        cfg.object_max_xy_pos[k] = float(object_max[k])
    mover_size = d['mover_size']
    # pushing:279-288
    if d['c_shape'] == 'circle':
        min_mo_dist = max(np.linalg.norm(object_length_xy + mover_size.flatten()[:2], ord=2), d['c_size'] + d['c_size_offset'])
    else:
        min_mo_dist = max(
            np.linalg.norm(object_length_xy + mover_size.flatten()[:2], ord=2),
            np.linalg.norm(np.asarray(d['c_size']) + d['c_size_offset'], ord=2),
        )
    cfg.min_mo_dist = float(min_mo_dist)
    cfg.object_noise_xy = 1e-5  # pushing:178
    cfg.object_half_xy = object_length_xy
    cfg.object_mass = 0.01  # pushing:175
    cfg.object_damping = 0.01  # pushing:337
    cfg.mover_half[0], cfg.mover_half[1] = float(mover_size[0, 0]), float(mover_size[0, 1])
    mm = d['mover_mass']
    cfg.mover_mass = float(mm if not isinstance(mm, np.ndarray) else mm[0])
    cfg.imp_k_rot = 0.1  # pushing:266
    cfg.gravity = 9.81  # basic_envs.py:1132
    cfg.friction = 1.0  # MuJoCo default geom friction
    cfg.solref[0], cfg.solref[1] = 0.02, 1.0  # MuJoCo defaults
    for k, val in enumerate((0.9, 0.95, 0.001, 0.5, 2.0)):
        cfg.solimp[k] = val
    # projected Gauss-Seidel sweeps per substep, started from the previous substep's forces (MuJoCo's default warm start):
    # 3 warm sweeps are as close to the converged solution as 8 cold ones (tests/test_pushing_independent.py)
    cfg.contact_iterations = int(contact_iterations)
    cfg.contact_warm_start = int(bool(contact_warm_start))
    d['min_mo_dist'] = float(min_mo_dist)
    d['object_min_xy_pos'] = object_min
    d['object_max_xy_pos'] = object_max
    d['obs_dim'] = (2 + int(bool(learn_jerk))) * 2
    d['goal_dim'] = 2
    d['action_dim'] = 2
    # the env classes hand the same output buffers to every call: desired_goal rows are rewritten only when they change
    # float64_outputs (GPR_OUT_FLOAT64): observation / goal arrays in the reference's dtype (the single-env classes set it)
    cfg.output_flags = (1 if goal_output_on_change else 0) | (2 if float64_outputs else 0)
    return cfg, d
