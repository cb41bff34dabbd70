import csv, subprocess, sys
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 22
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = next(i for i, r in enumerate(rows) if len(r) > 3 and 'Instructions Executed' in r)
hdr = rows[h]
idx={k:hdr.index(k) for k in ['# Samples','stall_long_sb','stall_barrier','stall_short_sb','stall_wait','Instructions Executed','stall_mio','stall_lg','stall_math','stall_not_selected','stall_selected','stall_dispatch','stall_no_inst','stall_branch_resolving']}
cur=None; out=[]
for r in rows[h+1:]:
    if r and r[0].isdigit(): cur=int(r[0]); continue
    if len(r)>idx['stall_long_sb'] and r[2].startswith('0x'):
        try: out.append((int(r[idx['# Samples']]),{k:int(r[v]) for k,v in idx.items()},cur,r[3].strip()[:60]))
        except: pass
tot=sum(o[0] for o in out)
print('samples',tot,{k:sum(o[1][k] for o in out) for k in idx if k.startswith('stall')})
for o in sorted(out,key=lambda o:-o[0])[:topn]:
    d=o[1]; print(f"samp {o[0]:5d} long {d['stall_long_sb']:5d} bar {d['stall_barrier']:5d} short {d['stall_short_sb']:4d} wait {d['stall_wait']:4d} L{o[2]} exec {d['Instructions Executed']:8d} {o[3]}")
