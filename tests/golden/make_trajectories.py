"""Generate tests/golden/reference_trajectories.npz by running the reference's UNMODIFIED reset() / step().

Run in the build container (needs the read-only reference mount; the GPU box only replays the committed file):

    python tests/golden/make_trajectories.py

For every case of ``ref_trajectory.PLANNING_CASES`` / ``PUSHING_CASES`` the reference env — on the closed-form MuJoCo
stand-in (tests/mujoco_standin.py), with the oracle's noise variates served through its ``rng_noise`` attribute
(ref_trajectory.OracleNoise) — is driven for K trajectories x T steps and everything it returns is stored.  The script
refuses to write a file the oracle does not reproduce bit for bit.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.dirname(HERE), ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import ref_trajectory as rt  # noqa: E402

SEED = 20261018

if __name__ == '__main__':
    out = {}
    for i, name in enumerate(rt.PLANNING_CASES):
        rec = rt.record_planning_case(name, SEED + i)
        cfg, _ = rt.planning_case_config(name)
        bad = rt.replay_planning_on_oracle(rec, cfg)
        assert not bad, (name, bad[:5])
        print(f'{name:28s} resets {int(rec["reset_before"].sum()):3d}  terminated {int(rec["terminated"].sum()):3d}  truncated '
              f'{int(rec["truncated"].sum()):2d}  success {int(rec["info"][:, :, 0].sum()):2d}  mover {int(rec["info"][:, :, 1].sum()):3d}  '
              f'wall {int(rec["info"][:, :, 2].sum()):3d}')
        for k, v in rec.items():
            out[f'planning/{name}/{k}'] = v
    for i, name in enumerate(getattr(rt, 'PUSHING_CASES', {})):
        rec = rt.record_pushing_case(name, SEED + 100 + i)
        cfg, _ = rt.pushing_case_config(name)
        bad = rt.replay_pushing_on_oracle(rec, cfg)
        assert not bad, (name, bad[:5])
        print(f'{name:28s} resets {int(rec["reset_before"].sum()):3d}  terminated {int(rec["terminated"].sum()):3d}  truncated '
              f'{int(rec["truncated"].sum()):2d}  success {int(rec["info"][:, :, 0].sum()):2d}  wall {int(rec["info"][:, :, 2].sum()):3d}')
        for k, v in rec.items():
            out[f'pushing/{name}/{k}'] = v
    path = os.path.join(HERE, 'reference_trajectories.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')
