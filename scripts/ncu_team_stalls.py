import csv, subprocess, sys
rep=sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = next(i for i, r in enumerate(rows) if len(r) > 3 and 'Instructions Executed' in r)
hdr = rows[h]
keys=[k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
idx={k:hdr.index(k) for k in keys}; iS=hdr.index('# Samples'); iI=hdr.index('Instructions Executed')
lines=open('localexpstereo_b200/csrc/lexp_kernels.cuh').read().split('\n')
def find(s): return next(i+1 for i,l in enumerate(lines) if s in l)
marks=[('pro',0),('A',find('team A')),('H',find('team H')),('C',find('team C')),('E',find('team E')),('end',find('K0: one-time'))]
agg={m[0]:{k:0 for k in keys} for m in marks}; ins={m[0]:0 for m in marks}; smp={m[0]:0 for m in marks}
cur=None
def team(l):
    t='pro'
    for n,a in marks:
        if l>=a: t=n
    return t
for r in rows[h+1:]:
    if r and r[0].isdigit(): cur=int(r[0]); continue
    if len(r)>iS and r[2].startswith('0x') and cur:
        t=team(cur) if cur>=marks[1][1] or cur>=101 else 'pro'
        if cur<101: t='helpers'
        agg.setdefault(t,{k:0 for k in keys}); ins.setdefault(t,0); smp.setdefault(t,0)
        try:
            ins[t]+=int(r[iI]); smp[t]+=int(r[iS])
            for k in keys: agg[t][k]+=int(r[idx[k]])
        except: pass
for t in agg:
    if ins[t]==0: continue
    top=sorted(agg[t].items(),key=lambda kv:-kv[1])[:7]
    print(t,'instr',ins[t],'samples',smp[t],' '.join(f"{k[6:]}={v}" for k,v in top))
