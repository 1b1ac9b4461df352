#!/usr/bin/env python
"""Prints the key numbers of an .ncu-rep (run where ncu is installed): tools/ncu_summary.py file.ncu-rep [kernel-regex]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('----')
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"{w} = {r[i][:70]} {units[i]}")
    items = []
    for i, h in enumerate(hdr):
        if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try: items.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError: pass
    print('stalls per issue: ' + ' '.join(f"{h}={v:.2f}" for v, h in sorted(items, reverse=True)[:9]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    byop, exop, top, tot = collections.Counter(), collections.Counter(), [], 0
    for r in rows[2:]:
        if len(r) <= iex: continue
        s = int(r[isamp] or 0); e = int(r[iex] or 0)
        toks = r[ia].split()
        op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
        byop[op] += s; exop[op] += e; tot += s
        top.append((s, e, r[ia].strip()[:60]))
    print('total samples', tot, 'instructions', sum(exop.values()))
    for op, s in byop.most_common(16):
        print(f"{op:10s} samples {100*s/tot:5.1f}%  executed {100*exop[op]/sum(exop.values()):5.1f}%")
    for t in sorted(top, reverse=True)[:14]: print(t)
