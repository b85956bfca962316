"""Reduces an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:conv_` launch
list of tools/profile_step.py (eager forward steps) to the per-launch DRAM traffic of the conv kernels over the LAST
step, and writes profiles/conv_traffic.json (read by bench.py for roofline.traffic).
Usage: python tools/conv_traffic.py gpurun_out/conv_traffic.csv <launches_per_step>"""
import collections
import csv
import json
import os
import sys

path, per_step = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ik, im, iu, iv, iid = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value', 'ID'))
per = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(',', ''))
    u = r[iu].lower()
    scale = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'usecond': 1e-6,
             'nsecond': 1e-9, 'msecond': 1e-3, 'second': 1}.get(u, 1)
    per.setdefault(r[iid], {'name': r[ik]})[r[im]] = v * scale
launches = list(per.values())[-per_step:]
rd = sum(x['dram__bytes_read.sum'] for x in launches)
wr = sum(x['dram__bytes_write.sum'] for x in launches)
t = sum(x['gpu__time_duration.sum'] for x in launches)
out = {'source': f'ncu dram__bytes_read.sum + dram__bytes_write.sum over the {len(launches)} conv launches of one eager '
                 f'B=64 forward ({os.path.basename(path)}), averaged per launch',
       'launches': len(launches), 'dram_bytes_per_launch': (rd + wr) / len(launches),
       'dram_read_bytes_per_step': rd, 'dram_write_bytes_per_step': wr, 'ncu_time_s_per_step': t}
try:      # the commit the capture was taken from (the working tree gpurun snapshotted)
    import subprocess
    out['commit'] = subprocess.check_output(['git', 'rev-parse', '--short', 'HEAD'], text=True,
                                            cwd=os.path.dirname(os.path.abspath(__file__))).strip()
except Exception:
    pass
print(json.dumps(out, indent=1))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, 'profiles', 'conv_traffic.json'), 'w'), indent=1)
