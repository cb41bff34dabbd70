#!/usr/bin/env python
"""Per-team stall samples of lexp_fused_kernel from the SASS page of an .ncu-rep (no source mapping needed): the code of
a team is one contiguous address range that ends with the team's EXIT; ranges are labelled by the thread count of the
named barriers they use (0xa0 = link A-H, 0x80 = H-C / C-H, 0x60 = H-E).
usage: ncu -i rep --page source --csv --print-source sass --launch-count 1 > x.csv; python scripts/ncu_sass_teams.py x.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
st = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
segs, cur = [], []
for r in rows[h + 1:]:
    if len(r) < len(hdr): continue
    cur.append(r)
    s = r[ix["Source"]].strip()
    if s.startswith("EXIT") or s.startswith("@") and " EXIT" in s and False:
        segs.append(cur); cur = []
if cur: segs.append(cur)
print(f"{len(segs)} segments")
for k, seg in enumerate(segs):
    ins = sum(int(r[ix["Instructions Executed"]]) for r in seg)
    smp = sum(int(r[ix["# Samples"]]) for r in seg)
    agg = collections.Counter()
    for r in seg:
        for s in st: agg[s[6:]] += int(r[ix[s]])
    bars = collections.Counter(r[ix["Source"]].split(",")[-1].strip() for r in seg if "BAR." in r[ix["Source"]])
    wf = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in seg)
    tags = sum(int(r[ix["L1 Tag Requests Global"]] or 0) for r in seg)
    print(f"seg {k}: {len(seg)} sass, warp-instr {ins}, samples {smp}, smem wavefronts {wf}, global tag requests {tags}, barriers {dict(bars)}")
    print("     ", " ".join(f"{a}={b}" for a, b in agg.most_common(8)))
    if len(sys.argv) > 2 and int(sys.argv[2]) == k:
        for r in seg:
            n = int(r[ix["# Samples"]])
            top = sorted(((int(r[ix[s]]), s[6:]) for s in st), reverse=True)[:2]
            print(f"{n:5d} {r[ix['Instructions Executed']]:>8} {r[ix['Source']].strip()[:70]:70s} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
