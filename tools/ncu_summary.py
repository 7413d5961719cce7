"""Summarise an ncu raw CSV (ncu -i rep --page raw --csv) per kernel launch; with a second
argument (the --page source CSV) also the sample share of the hottest SASS regions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('====', r[hdr.index('Kernel Name')][:80])
    for k in keys:
        if k in hdr:
            print(f"  {k:82s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
            try:
                st.append((float(r[i].replace(',', '')), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except ValueError:
                pass
    print('  stalls (warps per issue):', ', '.join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    h = rows[1]
    si, ai, xi = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
    data = [(i, int(r[xi] or 0), int(r[si] or 0), r[ai].strip(), r) for i, r in enumerate(rows[2:])
            if len(r) > si and r[xi].isdigit()]
    tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
    print(f"  source page: {tot} warp instructions, {ts} samples; regions (SASS lines, executions per "
          "instruction, share of instructions, share of samples):")
    reg, cur = [], None
    for i, x, s, src, r in data:
        if cur and abs(x - cur['x']) <= 0.25 * max(x, cur['x'], 1):
            cur['n'] += 1; cur['ex'] += x; cur['s'] += s; cur['end'] = i
        else:
            if cur: reg.append(cur)
            cur = {'start': i, 'end': i, 'x': x, 'n': 1, 'ex': x, 's': s}
    reg.append(cur)
    for r in reg:
        if r['ex'] > 0.02 * tot or r['s'] > 0.02 * ts:
            print(f"    lines {r['start']:5d}-{r['end']:5d}  n={r['n']:4d}  exec={r['x']:9d}  instr {100*r['ex']/tot:5.1f} %  samples {100*r['s']/ts:5.1f} %")
    names = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
    tots = {n: 0 for n in names}
    for i, x, s, src, r in data:
        for n in names:
            v = r[h.index(n)]
            if v.isdigit(): tots[n] += int(v)
    print('  stall samples:', ', '.join(f"{k[6:]}={v}" for k, v in sorted(tots.items(), key=lambda kv: -kv[1])[:8]))
