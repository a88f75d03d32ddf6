"""Stall samples / instructions by SOURCE LINE from an ncu capture.
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:NAME > mix.csv
    python tools/lines_by_source.py mix.csv [top]
"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 50
f=None; res=[]
for r in rows:
    if not r: continue
    if r[0]=='File Path': f=r[1].split('/')[-1]; continue
    if r[0] in ('Function Name','Line No'):
        if r[0]=='Line No': hdr=r
        continue
    if r[0]!='' and r[0].isdigit():
        si=hdr.index('# Samples'); ii=hdr.index('Instructions Executed'); ti=hdr.index('Thread Instructions Executed')
        try: res.append((f,int(r[0]),int(r[si]),int(r[ii]),int(r[ti]),r[1].strip()[:100]))
        except: pass
ts=sum(x[2] for x in res); tin=sum(x[3] for x in res)
print('samples',ts,'inst',tin)
for x in sorted(res,key=lambda x:-x[3])[:top]:
    print(f'{100*x[2]/ts:5.2f}% samp {100*x[3]/tin:5.2f}% inst thr/inst {x[4]/max(x[3],1):4.1f} {x[0]}:{x[1]}  {x[5]}')
