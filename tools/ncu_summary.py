"""Condense an .ncu-rep (one kernel, `ncu --set full --import-source on`) into a
small text summary for profiles/: headline counters + SASS opcode mix per
32-ray surface event.   python tools/ncu_summary.py REPORT.ncu-rep EVENTS_PER_LAUNCH > profiles/x.txt"""
import collections
import csv
import io
import subprocess
import sys

rep, events = sys.argv[1], float(sys.argv[2])
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print(f'# kernel: {name}')
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f'{k:88s} {vals[i]:>16s} {units[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if 'Source' in r and 'Instructions Executed' in r)
i_s, i_e, i_n = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
mix, samples, total = collections.Counter(), collections.Counter(), 0
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= i_e or not r[i_s].split():
        continue
    tok = r[i_s].split()
    op = (tok[1] if tok[0].startswith('@') else tok[0]).split('.')[0]
    mix[op] += int(r[i_e]); samples[op] += int(r[i_n]); total += int(r[i_e])
per = events / 32.0
print(f'\n# SASS mix: warp-instructions per 32-ray surface event ({events:.0f} events per launch)')
print(f'{"TOTAL":10s} {total / per:8.2f}')
for op, n in mix.most_common(24):
    print(f'{op:10s} {n / per:8.2f}   stall samples {samples[op]}')
