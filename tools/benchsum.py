"""Summarise bench.py JSON lines: python tools/benchsum.py gpurun_out/bench_a.log [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        l = [x for x in open(f) if x.startswith('{')][-1]
        d = json.loads(l)
        r = d.get('roofline') or {}
        print(f.split('/')[-1], 'value %.1fM' % (d['value'] / 1e6), 'ms %.4f' % d['ms_per_step'], 'e2e %.1fM' % (d['e2e']['value'] / 1e6),
              'k', round(r.get('kernel_ms', 0), 4), r.get('other_kernels_ms'), 'fails', d.get('reset_failures'))
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-600:])
