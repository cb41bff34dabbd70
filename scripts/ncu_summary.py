#!/usr/bin/env python
"""Summarise an .ncu-rep (raw metrics + instructions by CUDA source line) -> stdout / markdown."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers',
 'launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','sm__inst_executed.avg.per_cycle_elapsed','smsp__inst_executed.sum',
 'sm__cycles_elapsed.avg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'lts__t_bytes.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('##', d.get('Kernel Name','')[:70], 'grid', d.get('launch__grid_size'))
    for k in keys:
        if k in d: print(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
if len(sys.argv) > 2 and sys.argv[2] == 'lines':
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = next(i for i, r in enumerate(rows) if len(r) > 3 and 'Instructions Executed' in r)
    hdr = rows[h]; iI = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
    tot = 0; L = []
    for r in rows[h+1:]:
        if r and r[0].isdigit():
            try: n = int(r[iI]); s = int(r[iS])
            except Exception: continue
            L.append((n, s, int(r[0]), r[1].strip()[:100])); tot += n
    stot = sum(x[1] for x in L)
    print('total warp instr', tot, 'samples', stot)
    for n, s, l, t in sorted(L, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
        print(f"{n:>9} {100*n/tot:5.1f}%  samp {100*s/max(stot,1):5.1f}%  L{l}: {t}")
