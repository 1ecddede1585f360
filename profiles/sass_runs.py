"""Compress the SASS page of an ncu report (ncu -i rep --page source --csv --print-source sass > x.csv)
into runs of instructions with (nearly) equal execution counts: the phases of the kernel.
usage: sass_runs.py x.csv"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=rows[2:]
tot=sum(int(r[ix["Instructions Executed"]]) for r in data)
print("total",tot)
runs=[]
for k,r in enumerate(data):
    c=int(r[ix["Instructions Executed"]]); t=int(r[ix["Thread Instructions Executed"]]); s=int(r[ix["# Samples"]]); op=r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1]
    if runs and abs(runs[-1][2]-c)<=0.02*max(c,1):
        runs[-1][1]=k; runs[-1][3]+=c; runs[-1][4]+=t; runs[-1][5]+=s; runs[-1][6].append(op)
    else: runs.append([k,k,c,c,t,s,[op]])
cum=0
for a,b,c,ci,t,s,ops in runs:
    cum+=ci
    if ci/tot<0.002: continue
    import collections
    oc=collections.Counter(ops).most_common(6)
    print(f"{a:5d}-{b:5d} n={b-a+1:4d} exec/instr={c/1e6:8.3f}M share={100*ci/tot:5.2f}% cum={100*cum/tot:5.1f}% thr={t/max(ci,1):4.1f} smp={s:6d} {' '.join(f'{o}:{n}' for o,n in oc)}")
