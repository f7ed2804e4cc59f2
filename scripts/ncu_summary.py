"""Summarise an .ncu-rep: headline metrics + stall breakdown + per-region sample shares.
   python scripts/ncu_summary.py REPORT [KERNEL_SUBSTR]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ksub = sys.argv[2] if len(sys.argv) > 2 else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_static',
        'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if ksub not in d['Kernel Name']:
        continue
    print('##', d['Kernel Name'][:70])
    for k in KEYS:
        if k in d:
            print('  %-84s %s %s' % (k, d[k], rows[1][hdr.index(k)]))
    st = [(float(d[k]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
          for k in hdr if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and d[k] not in ('', 'nan')]
    st.sort(reverse=True)
    print('  stalls (warps per issue):', ', '.join('%s %.2f' % (n, v) for v, n in st[:8]))
