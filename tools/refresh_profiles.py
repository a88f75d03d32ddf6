"""Turn what tools/gpu_job.sh left in gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU needed):
    python tools/refresh_profiles.py r2 "final round-2 state, commit <hash>"
"""
import csv, io, json, os, subprocess, sys

tag, note = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def run(*a):
    return subprocess.run(a, capture_output=True, text=True, cwd=ROOT).stdout


def json_line(path):
    for l in open(path):
        if l.startswith('{'):
            return l
    raise SystemExit(f'no JSON line in {path}')


for src, dst in (('bench_default', 'planning4'), ('bench_pushing', 'pushing'), ('bench_planning8box', 'planning8box'),
                 ('bench_reference', 'reference_arm')):
    open(os.path.join(P, f'{tag}_bench_{dst}.json'), 'w').write(json_line(os.path.join(G, src + '.log')))

cmd = 'python bench.py --steps 10 --warmup 3 --quick --no-cpu --no-extra --repeats 1'
with open(os.path.join(P, f'{tag}_launches_planning4.txt'), 'w') as f:
    f.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none -c 200 : {cmd}   ({note})\n')
    f.write('# per-launch times are cold-cache and serialised: compare SHARES\n')
    f.write(run(sys.executable, 'tools/ncu_summary.py', 'launches', os.path.join(G, 'launches.csv')))

traffic = {'per_kernel': {}, 'note': f'{note}; bytes per launch from ncu --set full: dram__bytes_read/write.sum and lts__t_sectors.sum x 32 B '
                                     '(first captured launch of each kernel); the entry named after the workload is its dominant kernel'}
DOMINANT = {'planning4': 'planning_step_kernel', 'planning8box': 'planning_step_kernel', 'pushing': None}
for wl in ('planning4', 'planning8box', 'pushing'):
    rep = os.path.join(G, f'prof_{wl}.ncu-rep')
    with open(os.path.join(P, f'{tag}_full_{wl}.txt'), 'w') as f:
        f.write(f'# ncu --set full --clock-control none --import-source on : bench.py workload {wl} ({note}); columns = kernels in capture order\n')
        f.write(run(sys.executable, 'tools/ncu_summary.py', 'full', rep))
    rows = list(csv.reader(io.StringIO(run('ncu', '-i', rep, '--page', 'raw', '--csv'))))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, want):
        v, u = float(r[col[name]].replace(',', '')), units[col[name]]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1, 'ms': 1e3, 'ns': 1e-3, 'sector': 32, 'Ksector': 32e3,
                 'Msector': 32e6}[u]
        return v * scale

    per = {}
    for r in rows[2:]:
        k = r[col['Kernel Name']].replace('void ', '').split('(')[0]
        if k in per:
            continue
        per[k] = {'dram_bytes_read': val(r, 'dram__bytes_read.sum', 'byte'), 'dram_bytes_write': val(r, 'dram__bytes_write.sum', 'byte'),
                  'lts_t_bytes': val(r, 'lts__t_sectors.sum', 'byte'), 'time_us': val(r, 'gpu__time_duration.sum', 'us')}
    traffic['per_kernel'][wl] = per
    dom = [k for k in per if DOMINANT[wl] is None or k.startswith(DOMINANT[wl])]
    traffic[wl] = sum(per[k]['dram_bytes_read'] + per[k]['dram_bytes_write'] for k in dom)
    traffic[wl + '_lts_t_bytes'] = sum(per[k]['lts_t_bytes'] for k in dom)
    for kern in per:
        short = kern.split('<')[0]
        mix = run('ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '--kernel-name', f'regex:{short}')
        tmp = os.path.join(G, f'_mix_{wl}_{short}.csv')
        open(tmp, 'w').write(mix)
        with open(os.path.join(P, f'{tag}_lines_{wl}_{short}.txt'), 'w') as f:
            f.write(f'# {kern}: stall samples / warp instructions by source line (ncu --page source --print-source cuda,sass; {note})\n')
            f.write('# percentages are of the sums over the line rows; inlined code can be listed under more than one line\n')
            f.write(run(sys.executable, 'tools/lines_by_source.py', tmp, '40'))
json.dump(traffic, open(os.path.join(P, 'traffic.json'), 'w'), indent=1)
print(json.dumps({k: v for k, v in traffic.items() if k not in ('per_kernel', 'note')}, indent=1))
