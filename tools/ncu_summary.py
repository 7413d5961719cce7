"""Summarise an ncu raw CSV (ncu -i rep --page raw --csv) per kernel launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('====', r[hdr.index('Kernel Name')][:60])
    for k in keys:
        if k in hdr:
            print(f"  {k:72s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
            try:
                st.append((float(r[i].replace(',', '')), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except ValueError:
                pass
    print('  stalls (warps per issue):', ', '.join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
