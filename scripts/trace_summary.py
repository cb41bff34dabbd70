#!/usr/bin/env python
"""Summarises gpurun_out/trace.txt (written by a -DLEXP_TRACE=1 build, see scripts/gpu_trace.sh): per launch size class
(layer), the average cycles per 2-row chunk every team spends busy (and, inside that, stalled on its own global loads) / waiting
for input / waiting for an output buffer.
The team with the smallest waits is the one the pipeline is waiting for."""
import collections
import re
import sys

rows = collections.defaultdict(list)
for line in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace.txt"):
    m = re.match(r"launch items=(\d+) chunks/item=([\d.]+)", line)
    if not m:
        continue
    items, chunks = int(m.group(1)), float(m.group(2))
    teams = {}
    for t in re.finditer(r"\| (\w+) total (\d+) wait_in (\d+) wait_out (\d+) busy (-?\d+) wait_ld (\d+)", line):
        teams[t.group(1)] = [float(x) / chunks for x in t.groups()[1:]]
    key = "layer0 (>=300 items)" if items >= 300 else "layer1 (40..299 items)" if items >= 40 else "layer2 (<40 items)"
    rows[key].append((items, chunks, teams))
for key in sorted(rows):
    r = rows[key]
    print(f"{key}: {len(r)} launches, {sum(x[0] for x in r) / len(r):.0f} items, {sum(x[1] for x in r) / len(r):.1f} chunks per item")
    for team in ("A", "H1", "C", "H2", "E"):
        v = [x[2][team] for x in r if team in x[2]]
        if not v:
            continue
        avg = [sum(c[i] for c in v) / len(v) for i in range(5)]
        print(f"   team {team:2s}: per chunk {avg[0]:7.0f} cycles = busy {avg[3]:7.0f} (of which waiting for its own global loads {avg[4]:6.0f})"
              f" + wait_in {avg[1]:7.0f} + wait_out {avg[2]:7.0f}")
