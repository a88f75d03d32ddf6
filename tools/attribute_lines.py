"""Attribute ncu per-SASS-instruction counts to source lines.
    python tools/attribute_lines.py <nvdisasm -g output of ONE kernel> <ncu --page source --csv file> [top]
"""
import csv, re, sys, collections
sass, ncu_csv = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
addr2line = {}
cur = None
chain = []
MODE = sys.argv[4] if len(sys.argv) > 4 else 'outer'   # outer: outermost (kernel body) line; planning: deepest line in gpr_planning.cuh
for ln in open(sass, errors='replace'):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        chain.append((m.group(1).split('/')[-1], int(m.group(2)), False))
        if MODE == 'outer':
            cur = chain[-1]
        elif MODE == 'inner':
            cur = chain[0]
        else:
            cur = next((c for c in chain if c[0] == 'gpr_planning.cuh'), chain[-1])
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
        chain = []
rows = list(csv.reader(open(ncu_csv)))
# first kernel only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
ai, ii, si = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
first = None
inst = collections.Counter(); samp = collections.Counter(); n=0
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ('Kernel Name', 'Address'):
        break
    a = int(r[ai], 16) if r[ai].startswith('0x') or re.fullmatch(r'[0-9a-f]+', r[ai]) else None
    if a is None: continue
    if first is None: first = a
    off = a - first
    key = addr2line.get(off, ((None, 0, False), '?'))[0]
    key = (key[0], key[1]) if key else (None, 0)
    inst[key] += int(r[ii] or 0); samp[key] += int(r[si] or 0); n+=1
ti, ts = sum(inst.values()), sum(samp.values())
print(f'instructions {ti:,}  samples {ts:,}  sass rows {n}')
print(f'{"inst%":>6} {"samp%":>6}  line')
for k, v in inst.most_common(top):
    print(f'{100*v/ti:6.2f} {100*samp[k]/max(ts,1):6.2f}  {k[0]}:{k[1]}')
