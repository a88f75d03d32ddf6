#!/usr/bin/env python
"""bench.py — env-steps/s of the GymPR step path on B200 (metric and configs: BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload planning4|pushing|planning8box]

One "step" = one `step()` of every environment of the workload = num_cycles (40) control cycles of 1 ms each plus
observation / reward / termination / auto-reset.  Default workload = BASELINE configs[1]:
BenchmarkPlanningEnv-v0, 4 movers, 3x3 tiles, 65,536 environments PER GPU (weak scaling: env indices shard across ranks,
no traffic on the step path), reference-default kwargs (std_noise=1e-5), random actions, SAME_STEP auto-reset.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
    roofline      dominant kernel (the fused step kernel): algorithmic bytes (8N(9+3J)+17 per env-step) x
                  envs per launch / CUDA-event kernel time, against the measured HBM peak (MEASURED_PEAKS.json)
    cpu_baseline  the float64 C oracle (a port of the reference's algorithm; the MuJoCo-backed reference cannot be
                  installed offline) on this box's host cores, bounded sample
    e2e           same metric through the host-buffer API (gpr_step_host): NumPy action in, NumPy results out, H2D and
                  D2H copies inside the timed region
    value_no_noise / value_f64_exact ... extra context, see DESIGN.md
`--impl reference` times the oracle port with every host thread (rank 0 only) on the same workload/metric.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # BASELINE.json configs[1]
    'planning4': dict(kind='planning', num_envs=65536, env_id='BenchmarkPlanningEnv-v0',
                      kwargs=dict(layout_tiles=np.ones((3, 3)), num_movers=4),
                      desc='BenchmarkPlanningEnv-v0, 4 movers, 3x3 tiles, circle r=0.11, acc actions, 65,536 envs/GPU'),
    # BASELINE.json configs[3]
    'planning8box': dict(kind='planning', num_envs=262144, env_id='BenchmarkPlanningEnv-v0',
                         kwargs=dict(layout_tiles=np.ones((5, 5)), num_movers=8, learn_jerk=True,
                                     collision_params={'shape': 'box', 'size': np.array([0.08, 0.08])}),
                         desc='BenchmarkPlanningEnv-v0, 8 movers, 5x5 tiles, box 0.08x0.08, jerk actions, 262,144 envs/GPU'),
    # BASELINE.json configs[2]
    'pushing': dict(kind='pushing', num_envs=65536, env_id='BenchmarkPushingEnv-v0', kwargs=dict(),
                    desc='BenchmarkPushingEnv-v0, 1 mover + box object, 65,536 envs/GPU'),
}


def algorithmic_bytes_per_env_step(kind: str, N: int, J: int) -> int:
    """float32 I/O, SoA state. planning 8N(9+3J)+17; pushing 161 (+24 with jerk)."""
    return 8 * N * (9 + 3 * J) + 17 if kind == 'planning' else 161 + 24 * J


def measured_hbm_peak() -> tuple[float, str]:
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)['hbm_gbs']), 'measured'
        except Exception:
            pass
    return 6650.0, 'fallback'


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line), sampled through NVML every
    few milliseconds (the timed region of the default run is ~60 ms: `nvidia-smi` itself takes longer than that per query;
    it remains the fallback when pynvml is missing)."""

    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None
        self.source = 'nvidia-smi'
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the torch device
            import torch

            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), 'pci_bus_id') else None
            self._h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hh).bus == bus:
                        self._h = hh
                        break
            if self._h is None:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = pynvml
            self.source = 'nvml'
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n, h = self._nvml, self._h
        sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
        mx = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        return [sm, mx] + ['Active' if (r & b) else 'Not Active' for b in bits]

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.004 if self._nvml is not None else 0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for n, v in zip(self.NAMES, s[2:6]):
                    if str(v).lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons), 'samples': len(sm),
                'source': self.source}


def make_cfg(wl, num_envs, env_index_base=0, **over):
    import gymnasium_planar_robotics_b200 as gpr

    kw = dict(wl['kwargs'])
    kw.update(over)
    fn = gpr.planning_config if wl['kind'] == 'planning' else gpr.pushing_config
    return fn(num_envs=num_envs, env_index_base=env_index_base, **kw)


def host_threads() -> int:
    """Host cores available to this process (torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU baseline)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def time_oracle(wl, num_envs: int, steps: int, warmup: int, threads: int, seed: int = 0) -> float:
    """env-steps/s of the CPU oracle (float64 C restatement of the reference's algorithm) on `threads` host threads."""
    import gpr_oracle

    cfg, d = make_cfg(wl, num_envs, seed=seed)
    ora = gpr_oracle.OracleEnv(cfg, nthreads=threads)
    ora.reset(seed=seed)
    rng = np.random.default_rng(seed)
    lim = cfg.j_max if cfg.learn_jerk else cfg.a_max
    acts = [rng.uniform(-lim, lim, (num_envs, ora.action_dim)).astype(np.float32) for _ in range(4)]
    for i in range(warmup):
        ora.step(acts[i % 4])
    t0 = time.perf_counter()
    for i in range(steps):
        ora.step(acts[i % 4])
    return num_envs * steps / (time.perf_counter() - t0)


def reference_numpy_share():
    """The reference's UNMODIFIED NumPy share of one env-step, measured where its source is mounted
    (the build container) by tools/reference_numpy_share.py and committed as profiles/reference_numpy_share.json: the
    GPU box has no reference mount, so the bench line carries the committed measurement with its provenance."""
    path = os.path.join(ROOT, 'profiles', 'reference_numpy_share.json')
    if not os.path.exists(path):
        return None
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def run_reference(args, wl):
    """Reference arm: the reference's CPU implementation of the path.  The real reference needs the MuJoCo wheel, which
    cannot be installed offline, so this runs the oracle port with every host thread (kind 'port') on the SAME config as
    the B200 arm: the workload's full batch (65,536 envs for the headline), `--steps` steps of it."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    threads = host_threads()
    B = args.num_envs or wl['num_envs']
    v = time_oracle(wl, B, args.steps, max(args.warmup, 1), threads)
    cfg, d = make_cfg(wl, B)
    line = {
        'impl': 'reference', 'metric': 'env-steps/s', 'value': v, 'unit': 'env-steps/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * B / v, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': wl['desc'], 'env_id': wl['env_id'], 'envs_per_gpu': B, 'num_cycles': int(cfg.num_cycles), 'std_noise': cfg.std_noise[0],
                   'autoreset': 'same_step', 'sample': f'the full batch: {B} envs per step'},
        'cpu_baseline': {'value': v, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{B} envs x {args.steps} steps (the B200 arm\'s batch), OpenMP over envs; MuJoCo-backed reference not installable offline'},
        'e2e': {'value': v, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'reference_numpy_share': reference_numpy_share(),
    }
    print(json.dumps(line), flush=True)


KERNEL_NAMES = {'planning': ('planning_step_kernel', 'autoreset', 'planning_autoreset_kernel'),
                'pushing': ('pushing_step_kernel', 'contact', 'pushing_contact_kernel')}


def run_b200(args, wl):
    import torch
    import torch.distributed as dist

    import gymnasium_planar_robotics_b200 as gpr

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    # one process per GPU: run on (and first-touch pinned memory from) the cores of the GPU's own NUMA node
    full_affinity = os.sched_getaffinity(0) if hasattr(os, 'sched_getaffinity') else None
    numa_cpus = None if args.no_numa_bind else gpr.bind_to_gpu_numa(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    peak, peak_src = measured_hbm_peak()

    def build(w, B, **over):
        cls = gpr.BenchmarkPlanningVecEnv if w['kind'] == 'planning' else gpr.BenchmarkPushingVecEnv
        kw = dict(w['kwargs'])
        kw.update(over)
        # weak scaling: B envs per GPU, global index base = rank * B (results independent of the split)
        return cls(B, device=dev, env_index_base=rank * B, seed=args.seed, **kw)

    def timed(env, B, steps, warmup, per_kernel=False, reset=True):
        """Device time of `steps` steps: CUDA events around each gpr_step on the launching stream, L2 flushed in between.
        per_kernel: also record CUDA events around each kernel inside gpr_step (gpr_kernel_times).  An event between the
        step kernel and the kernel that follows it keeps the two from overlapping, so that pass is a separate one: it
        gives the kernels' own launch durations (roofline), the plain pass gives the step time (value)."""
        lim = env.j_max if env.learn_jerk else env.a_max
        acts = [(torch.rand((B, env.core.action_dim), device=dev, generator=gen) * 2 - 1) * lim for _ in range(8)]
        if reset:
            env.reset(seed=args.seed)
        for i in range(warmup):
            env.step(acts[i % 8])
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        n0 = env.core.launch_count
        env.core.kernel_times(per_kernel)  # per-kernel CUDA events inside gpr_step, on the launching stream
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            flush.zero_()
            ev[i][0].record()
            env.step(acts[i % 8])
            ev[i][1].record()
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        kt = env.core.kernel_times(False)
        return ms, env.core.launch_count - n0, wall, kt

    def measure(w, B, steps, warmup, repeats, clocks_into=None):
        """The device-resident leg of one workload: best of `repeats` timed blocks of `steps` steps (every block: max over
        ranks), then a per-kernel pass for the roofline of the dominant kernel."""
        env = build(w, B)
        blocks, launches, wall = [], 0, 0.0
        ctx = ClockSampler(local)
        with ctx:
            for r in range(repeats):
                ms, launches, wall, _ = timed(env, B, steps, warmup if r == 0 else 1, reset=(r == 0))
                blocks.append(allmax(ms))
            _, _, _, kt = timed(env, B, steps, 3, per_kernel=True, reset=False)  # same steps again, each kernel timed alone
        best = min(blocks)
        N, J = int(env.cfg.num_movers), int(env.cfg.learn_jerk)
        abytes = algorithmic_bytes_per_env_step(w['kind'], N, J)
        k_main, k_other_key, k_other = KERNEL_NAMES[w['kind']]
        # the dominant kernel: planning -> the fused step kernel; pushing -> the step is split by population into the free
        # kernel and the contact kernel (each env runs its 40 cycles in one or both), so both launches count
        dom_ms = kt['step_kernel_ms'] + (kt['autoreset_kernel_ms'] if w['kind'] == 'pushing' else 0.0)
        if w['kind'] == 'pushing':
            k_main = 'pushing_step_kernel + pushing_contact_kernel'
        achieved = abytes * B / (dom_ms * 1e-3) / 1e9
        res = {
            'env': env, 'value': world * B * steps / (best * 1e-3), 'ms_per_step': best / steps, 'blocks_ms_per_step': [b / steps for b in blocks],
            'launches': launches, 'wall': wall, 'clocks': ctx.summary(), 'abytes': abytes,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'peak_source': peak_src,
                         'algorithmic_bytes_per_env_step': abytes, 'algorithmic_bytes_per_launch': abytes * B, 'kernel': k_main,
                         'kernel_ms': dom_ms, 'kernels_ms': {'step': kt['step_kernel_ms'], k_other_key: kt['autoreset_kernel_ms']},
                         'other_kernels_ms': {} if w['kind'] == 'pushing' else {k_other_key: kt['autoreset_kernel_ms']},
                         'whole_step_frac': abytes * B / (best / steps * 1e-3) / 1e9 / peak},
        }
        return res

    B = args.num_envs or wl['num_envs']
    m = measure(wl, B, args.steps, args.warmup, args.repeats)
    env = m['env']
    stats = env.episode_stats(reset=True, all_reduce=world > 1)
    fails = env.core.reset_failures()

    # ---- end to end through the host-buffer API (NumPy in / NumPy out), H2D + D2H inside the timed region
    def e2e_leg(env, B, steps, repeats):
        lim = env.j_max if env.learn_jerk else env.a_max
        rng = np.random.default_rng(5 + rank)
        hacts = []
        for _ in range(4):  # page-locked host arrays (the contract's "pinned host memory"); filled on the host
            buf = env.core.pinned_action_buffer()
            buf[:] = rng.uniform(-lim, lim, (B, env.core.action_dim)).astype(np.float32)
            hacts.append(buf)
        for i in range(3):
            env.step_host(hacts[i % 4])
        blocks = []
        for r in range(repeats):
            barrier()
            t0 = time.perf_counter()
            acc = 0.0
            for i in range(steps):
                out = env.step_host(hacts[i % 4])
                acc += float(out[1][0])  # the host reads the step's result (reward of env 0) before issuing the next step
            barrier()
            blocks.append(allmax(time.perf_counter() - t0))
        best = min(blocks)
        env.core.kernel_times(True)  # device time of the same kernels when their I/O lives in pinned host memory (own pass)
        done_rows = 0
        for i in range(steps):
            out = env.step_host(hacts[i % 4])
            done_rows += int(np.count_nonzero(out[2] | out[3]))
        kt = env.core.kernel_times(False)
        # bytes a step actually moves device -> host: dense rows every step, final_* and desired_goal rows only for the
        # environments that finished (GPR_OUT_GOAL_ON_CHANGE; final rows are written for finished envs only)
        od, gd = env.core.obs_dim, env.core.goal_dim
        dense = 4 * (od + gd) + 4 + sum(1 for k in env.core._host if env.core._host[k].dtype == np.uint8)
        per_done = 4 * (od + gd + gd) + 4 * gd
        return {'value': world * B * steps / best, 'seconds': best, 'blocks': blocks, 'steps': steps, 'kt': kt,
                'h2d': B * env.core.action_dim * 4, 'd2h_written': int(B * dense + per_done * done_rows / steps),
                'd2h_buffers': int(sum(v.nbytes for v in env.core._host.values()))}

    e2e_steps = max(10, args.steps)  # (as many as the device-resident leg: a single host hiccup must not dominate)
    e2e = e2e_leg(env, B, e2e_steps, args.repeats)
    e2e_blocks, e2e_s, e2e_kt, e2e_value = e2e['blocks'], e2e['seconds'], e2e['kt'], e2e['value']
    h2d, d2h_written, d2h_buffers = e2e['h2d'], e2e['d2h_written'], e2e['d2h_buffers']
    env.close()

    # context number: the same workload without sensor noise (the reference tests' parity setting).  EVERY rank runs it:
    # timed() holds barriers, so it must never sit inside rank-conditional code.
    extra = {}
    if not args.quick:
        e0 = build(wl, B, std_noise=0.0)
        n0 = max(20, args.steps // 2)
        ms0, _, _, _ = timed(e0, B, n0, 3)
        extra['value_no_noise'] = world * B * n0 / (allmax(ms0) * 1e-3)
        e0.close()

    # ---- the other BASELINE configs, measured by the same method outside the headline's timed region (one GPU only: the
    #      scaling runs stay short).  Reported under `extra_workloads` so that every config is in the driver's record.
    extra_workloads = {}
    if world == 1 and not args.quick and not args.no_extra and args.workload == 'planning4':
        for name, w, nb, st in (('pushing', WORKLOADS['pushing'], 65536, 20), ('planning8box', WORKLOADS['planning8box'], 262144, 10),
                                ('planning4_1M', WORKLOADS['planning4'], 1048576, 10), ('pushing_1M', WORKLOADS['pushing'], 1048576, 10)):
            try:
                r = measure(w, nb, st, 3, 3)
                x = e2e_leg(r['env'], nb, st, 3)  # the same leg end to end (host buffers in / out)
                r['env'].close()
                extra_workloads[name] = {'workload': w['desc'].replace(f"{w['num_envs']:,} envs/GPU", f'{nb:,} envs/GPU'), 'envs_per_gpu': nb,
                                         'value': r['value'], 'unit': 'env-steps/s', 'ms_per_step': r['ms_per_step'], 'steps': st, 'best_of': 3,
                                         'e2e': {'value': x['value'], 'unit': 'env-steps/s', 'ms_per_step': 1e3 * x['seconds'] / st,
                                                 'h2d_bytes_per_step': x['h2d'], 'd2h_bytes_per_step': x['d2h_written']},
                                         'roofline': r['roofline'], 'clocks': r['clocks']}
            except Exception as exc:  # an extra leg must never take the headline line down with it
                extra_workloads[name] = {'error': repr(exc)[:300]}

    if rank == 0:
        traffic, traffic_l2 = None, None
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    tj = json.load(f)
                traffic = tj.get(args.workload)
                traffic_l2 = tj.get(args.workload + '_lts_t_bytes')
            except Exception:
                traffic = None
        # CPU baseline on this box's host cores (rank 0, N=1 only; bounded sample: the FULL batch, 10 steps)
        cpu = None
        if world == 1 and not args.no_cpu:
            if full_affinity is not None:
                os.sched_setaffinity(0, full_affinity)  # the CPU baseline may use every host core again
            threads = host_threads()
            cs = 10
            v1 = time_oracle(wl, B, cs, 2, threads)
            cpu = {'value': v1, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
                   'sample': f'{B} envs x {cs} steps of the same workload (the full batch; float64 C oracle, OpenMP over envs); the '
                             f'MuJoCo-backed reference is not installable offline',
                   'reference_numpy_share': reference_numpy_share()}
        roof = dict(m['roofline'])
        roof['traffic'] = traffic
        roof['traffic_l2_lts_t_bytes'] = traffic_l2
        roof['kernel_timing'] = 'CUDA events around this kernel alone, a second pass over the same steps (the timed pass overlaps the following kernel with its tail)'
        roof['note'] = 'issue-bound kernel (40-cycle float64 loop per env): see profiles/ for issue-slot and stall breakdown'
        line = {
            'metric': 'env-steps/s', 'value': m['value'], 'unit': 'env-steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': m['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': wl['desc'], 'env_id': wl['env_id'], 'envs_per_gpu': B, 'num_cycles': int(env.cfg.num_cycles),
                       'substeps_per_s': m['value'] * int(env.cfg.num_cycles), 'std_noise': env.cfg.std_noise[0], 'autoreset': 'same_step',
                       'actions': 'uniform(-max,max), 8 pre-generated device tensors cycled',
                       'l2': 'flushed between timed steps (256 MiB memset, outside the event pairs)', 'parallelism': f'env-shard x{world}',
                       'timing': f'best of {args.repeats} blocks of {args.steps} steps, each block max over ranks',
                       'numa_bind': f'{len(numa_cpus)} cores local to the GPU' if numa_cpus else 'none'},
            'timed_blocks_ms_per_step': m['blocks_ms_per_step'],
            'roofline': roof,
            'cpu_baseline': cpu,
            'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h_written,
                    'd2h_buffer_bytes': d2h_buffers, 'steps': e2e_steps, 'best_of': args.repeats, 'blocks_ms_per_step': [1e3 * b / e2e_steps for b in e2e_blocks],
                    'ms_per_step': 1e3 * e2e_s / e2e_steps, 'kernel_ms_with_host_io': {'step': e2e_kt['step_kernel_ms'], KERNEL_NAMES[wl['kind']][1]: e2e_kt['autoreset_kernel_ms']},
                    'd2h_note': 'd2h_bytes_per_step = bytes the kernels actually write per step (dense observation / achieved_goal / reward / flags rows + final_* and desired_goal rows of finished envs only); d2h_buffer_bytes = size of all result buffers',
                    'api': f'{type(env).__name__}.step_host -> gpr_step_host: NumPy views of page-locked host arrays in/out, read and written in place by the kernels over PCIe (zero-copy), stream sync before returning'},
            'gpu_launches': int(m['launches']), 'clocks': m['clocks'],
            'episode_stats': stats, 'reset_failures': fails, 'wall_s_timed_region': m['wall'],
            'extra_workloads': extra_workloads,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='planning4', choices=sorted(WORKLOADS))
    ap.add_argument('--num-envs', type=int, default=0, help='envs per GPU (default: the workload\'s)')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--repeats', type=int, default=5, help='timed blocks of --steps steps; the best block is reported (SURVEY §8d)')
    ap.add_argument('--quick', action='store_true', help='skip the extra context measurements')
    ap.add_argument('--no-extra', action='store_true', help='skip the extra_workloads legs (pushing, planning8box, 1M-env points)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline')
    ap.add_argument('--no-numa-bind', action='store_true', help='do not pin the rank to the cores local to its GPU')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == '__main__':
    main()
