"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/gpr.h declares, the
ctypes mirror of gpr_config matches, and — with no GPU visible — everything fails LOUDLY (no CPU fallback)."""

import ctypes
import os
import re

import numpy as np
import pytest

import gymnasium_planar_robotics_b200 as gpr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    with open(os.path.join(ROOT, 'include', 'gpr.h')) as f:
        text = f.read()
    return sorted(set(re.findall(r'GPR_API\s+[\w\s\*]+?\b(gpr_\w+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = gpr._lib.load()
    declared = _header_symbols()
    assert len(declared) >= 18
    assert set(declared) == set(gpr._lib.EXPORTED_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_config_struct_matches_header():
    lib = gpr._lib.load()
    assert lib.gpr_config_bytes() == ctypes.sizeof(gpr._config.GprConfig)
    assert lib.gpr_abi_version() == gpr._config.GPR_ABI_VERSION
    import gpr_oracle

    assert gpr_oracle.lib().gpro_config_bytes() == ctypes.sizeof(gpr._config.GprConfig)


def test_planning_config_follows_reference_numbers():
    """planning:262-274 on the 3x3 layout: spawn box [0.11, 0.55]^2 (the reference's tile_size/2 quirk), min goal
    distance 2r; obs/action dims planning:242-259."""
    cfg, d = gpr.planning_config(num_envs=3, layout_tiles=np.ones((3, 3)), num_movers=4)
    assert (cfg.min_xy_pos[0], cfg.max_xy_pos[0]) == (0.11, np.max(np.linspace(0.12, 0.6, 3)) + 0.06 - 0.11)
    assert cfg.min_goal_dist == 0.22 and cfg.num_cycles == 40 and cfg.max_episode_steps == 50
    assert d['obs_dim'] == 8 and d['goal_dim'] == 8 and d['action_dim'] == 8
    assert list(cfg.tile_cx)[:3] == list(np.linspace(0.12, 0.6, 3))
    cfg, d = gpr.planning_config(num_envs=3, layout_tiles=np.ones((5, 5)), num_movers=8, learn_jerk=True,
                                 collision_params={'shape': 'box', 'size': np.array([0.08, 0.08]), 'offset': 0.01})
    assert d['obs_dim'] == 32 and cfg.min_goal_dist == 2 * np.linalg.norm(np.array([0.08, 0.08]) + 0.01)
    assert cfg.c_wall[1][7][0] == 0.08 + 0.0 + 0.01 and cfg.c_mover[0][7][1] == 0.08


def test_pushing_config_follows_reference_numbers():
    """pushing:250-288: object box [0.22, 0.44]^2, min mover-object distance max(||(.035+.0775)*(1,1)||, r)."""
    cfg, d = gpr.pushing_config(num_envs=2)
    assert np.allclose([cfg.object_min_xy_pos[0], cfg.object_max_xy_pos[0]], [0.22, 0.44])
    assert cfg.min_mo_dist == max(np.linalg.norm(0.035 + np.array([0.0775, 0.0775])), 0.11)
    assert cfg.threshold_pos == 0.05 and d['obs_dim'] == 4 and d['goal_dim'] == 2


def test_out_of_scope_kwargs_are_rejected():
    with pytest.raises(NotImplementedError):
        gpr.planning_config(num_envs=1, layout_tiles=np.ones((3, 3)), num_movers=1, render_mode='human')
    with pytest.raises(NotImplementedError):
        gpr.planning_config(num_envs=1, layout_tiles=np.ones((3, 3)), num_movers=1, mover_params={'shape': 'mesh'})
    with pytest.raises(ValueError):  # the reference asserts at basic_envs.py:650 for such shapes
        gpr.planning_config(num_envs=1, layout_tiles=np.ones((3, 3)), num_movers=1, collision_params={'shape': 'circle', 'size': 0.12})
    with pytest.raises(AssertionError):
        gpr.planning_config(num_envs=1, layout_tiles=np.array([[1, 2]]), num_movers=1)


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip('a GPU is visible')
    lib = gpr._lib.load()
    cfg, _ = gpr.planning_config(num_envs=4, layout_tiles=np.ones((3, 3)), num_movers=2)
    h = ctypes.c_void_p()
    rc = lib.gpr_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc == -3 and not h.value  # GPR_ERR_NO_DEVICE
    assert b'no CPU fallback' in lib.gpr_last_error()
    with pytest.raises(gpr.GprError):
        gpr.BenchmarkPlanningVecEnv(4, np.ones((3, 3)), 2)


def test_create_rejects_bad_configs():
    lib = gpr._lib.load()
    h = ctypes.c_void_p()
    cfg, _ = gpr.planning_config(num_envs=4, layout_tiles=np.ones((3, 3)), num_movers=2)
    cfg.struct_bytes = 12
    assert lib.gpr_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -6  # GPR_ERR_ABI_MISMATCH
    cfg, _ = gpr.planning_config(num_envs=4, layout_tiles=np.ones((3, 3)), num_movers=2)
    cfg.num_movers = 33
    assert lib.gpr_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1
    assert b'num_movers' in lib.gpr_last_error()
    assert lib.gpr_step(None, None, None, None) == -1


def test_shard_range_partitions_exactly():
    from gymnasium_planar_robotics_b200.envs import shard_range

    for total in (1, 7, 65536, 8388608 + 3):
        for w in (1, 2, 4, 8):
            parts = [shard_range(total, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (b0, c0), (b1, _) in zip(parts, parts[1:]):
                assert b0 + c0 == b1


def test_env_ids_register_like_the_reference(monkeypatch):
    """__init__.py:21-41 of the reference: both ids, max_episode_steps=50, registered on import (here: when gymnasium is
    importable; a recording stub stands in for it)."""
    import importlib
    import sys
    import types

    calls = []
    reg = types.ModuleType('gymnasium.envs.registration')
    reg.registry = {}
    reg.register = lambda **kw: (calls.append(kw), reg.registry.__setitem__(kw['id'], kw))
    gym = types.ModuleType('gymnasium')
    envs = types.ModuleType('gymnasium.envs')
    gym.envs, envs.registration = envs, reg
    for name, mod in (('gymnasium', gym), ('gymnasium.envs', envs), ('gymnasium.envs.registration', reg)):
        monkeypatch.setitem(sys.modules, name, mod)
    assert gpr.register_gymnasium_envs() is True
    assert [c['id'] for c in calls] == ['BenchmarkPlanningEnv-v0', 'BenchmarkPushingEnv-v0']
    assert all(c['max_episode_steps'] == 50 for c in calls)
    assert calls[0]['entry_point'].endswith('envs:BenchmarkPlanningEnv') and calls[1]['vector_entry_point'].endswith('envs:BenchmarkPushingVecEnv')
    mod_name, cls_name = calls[0]['entry_point'].split(':')
    assert hasattr(importlib.import_module(mod_name), cls_name)  # the entry point resolves
    assert gpr.register_gymnasium_envs() is True and len(calls) == 2  # idempotent
