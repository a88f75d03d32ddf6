"""Summarise ncu outputs (run here, no GPU needed):
    python tools/ncu_summary.py launches gpurun_out/launches.csv
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep
"""
import collections, csv, subprocess, sys, io

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        k = r[ki][:90]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f'{"total ms":>10} {"n":>5} {"avg us":>9} {"share":>6}  kernel')
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'{t/1e6:10.3f} {n:5d} {t/n/1e3:9.1f} {100*t/tot:5.1f}%  {k}')

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum']

def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for i, h in enumerate(hdr):
        if h == 'Kernel Name' or h in KEYS or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
            vals = [r[i] for r in rows[2:]]
            print(f'{h:95s} {units[i]:12s} {vals}')

if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
