"""Turns an .ncu-rep into the transposed raw-page summary kept under profiles/ (metric, unit, one column per launch).
Usage: python tools/ncu_summary.py in.ncu-rep out.csv"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(r'^(Kernel Name|Block Size|Grid Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum($|\.per_second)|'
                  r'dram__throughput\.avg\.pct|sm__pipe_tensor\w*cycles_active\.avg\.pct|sm__throughput\.avg\.pct|'
                  r'sm__inst_executed\.sum$|sm__inst_executed_pipe_\w+\.sum$|sm__warps_active\.avg\.pct|launch__\w+|'
                  r'smsp__issue_active\.avg\.pct|smsp__inst_executed\.sum$|l1tex__throughput\.avg\.pct|'
                  r'lts__throughput\.avg\.pct|lts__t_bytes\.sum$|lts__t_sector_hit_rate\.pct|sm__cycles_elapsed\.(avg|max)$|'
                  r'sm__cycles_active\.avg$|smsp__cycles_active\.avg$|sm__pipe_(fp64|fma|alu|fmaheavy|xu)_cycles_active\.avg\.pct\w*elapsed|'
                  r'smsp__average_warps?_issue_stalled_\w+_per_issue_active\.ratio|l1tex__data_bank_conflicts\w*\.sum$|'
                  r'smsp__sass_thread_inst_executed_op_\w+_pred_on\.sum$|sm__sass_inst_executed_op_shared\w*\.sum$|'
                  r'sm__ctas_launched\.sum|gpc__cycles_elapsed\.max|sm__maximum_warps_per_active_cycle_pct|sm__occupancy\w*)')

rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
with open(out, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit'] + [f'launch{i}' for i in range(len(data))])
    for c, name in enumerate(hdr):
        if name in ('ID', 'Process ID', 'Process Name', 'Host Name', 'Context', 'Stream', 'Device', 'CC'):
            continue
        if not KEEP.match(name):
            continue
        w.writerow([name, units[c]] + [d[c] if c < len(d) else '' for d in data])
print('wrote', out, len(data), 'launches')
