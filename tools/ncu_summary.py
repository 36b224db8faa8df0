"""Markdown summary of an ncu --set full report: per kernel, the metrics DESIGN.md quotes.
usage: python tools/ncu_summary.py report.ncu-rep "title" > profiles/xxx.md"""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
want = ['launch__grid_size', 'launch__cluster_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_wait',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_selected',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle']
ix = {k: i for i, k in enumerate(h)}
print(f'# {title}\n')
for r in rows[2:]:
    name = r[ix['Kernel Name']].split('(')[0]
    print(f'### `{name}`\n\n|metric|value|unit|\n|---|---|---|')
    for k in want:
        if k in ix and r[ix[k]] != '':
            print(f'|{k}|{r[ix[k]]}|{units[ix[k]]}|')
    print()
