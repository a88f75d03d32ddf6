"""bench.py's reference arm runs without a GPU (it times the CPU oracle port): check the JSON-line contract here, so a
broken line is caught before the driver runs it on the GPU box."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    d = _run('--impl', 'reference', '--steps', '2', '--warmup', '1')
    assert d['impl'] == 'reference' and d['metric'] == 'env-steps/s' and d['unit'] == 'env-steps/s' and d['higher_is_better'] is True
    assert d['steps'] == 2 and d['warmup'] == 1 and d['n_gpus'] == 1 and d['value'] > 0 and d['ms_per_step'] > 0
    assert d['config']['workload'].startswith('BenchmarkPlanningEnv-v0, 4 movers')
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    assert d['e2e'] == {'value': d['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_ignores_omp_num_threads_and_non_zero_ranks_stay_silent():
    d = _run('--impl', 'reference', '--steps', '1', '--warmup', '1', '--workload', 'pushing', env={'OMP_NUM_THREADS': '1'})
    assert d['cpu_baseline']['cores'] == len(os.sched_getaffinity(0))  # torchrun pins OMP_NUM_THREADS=1: must not shrink the baseline
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1'],
                         capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK='1', WORLD_SIZE='2'))
    assert out.returncode == 0 and out.stdout.strip() == ''
