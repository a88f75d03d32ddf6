"""Opcode histogram (weighted by executed count) of one kernel from `ncu --page source --csv`.
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:<k> --launch-count 1 > k.csv ; python tools/sass_hist.py k.csv [top]
"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
si, ii, sm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops, samp, tot, tots = collections.Counter(), collections.Counter(), 0, 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] in ('Address', 'Kernel Name'):
        break
    n, s = int(r[ii] or 0), int(r[sm] or 0)
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si].strip())
    op = m.group(2).split('.')[0] if m else r[si].strip()
    ops[op] += n
    samp[op] += s
    tot += n
    tots += s
print(f'warp instructions executed {tot:,}; stall samples {tots:,}')
for op, n in ops.most_common(top):
    print(f'{op:10s} {100 * n / tot:6.2f}% inst   {100 * samp[op] / max(tots, 1):6.2f}% samples')
