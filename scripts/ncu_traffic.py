#!/usr/bin/env python
"""dram__bytes_read.sum + dram__bytes_write.sum of the grid_kernel launch in an `ncu --set full` report ->
profiles/grid_kernel_traffic.json (read by bench.py's roofline.traffic).
usage: python scripts/ncu_traffic.py X.ncu-rep <streams> <frames per launch> [out.json]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, S, F = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "grid_kernel_traffic.json")
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    if 'grid_kernel' not in d.get('Kernel Name', ''):
        continue
    def get(name):
        return float(d[name].replace(',', '')) * UNIT[units[hdr.index(name)]]
    rd, wr = get('dram__bytes_read.sum'), get('dram__bytes_write.sum')
    tu = units[hdr.index('gpu__time_duration.sum')]
    us = float(d['gpu__time_duration.sum'].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(tu, 1.0)
    res = {"kernel": "grid_kernel", "report": os.path.basename(rep), "streams": S, "frames_per_launch": F,
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "dram_bytes_per_frame_stream": (rd + wr) / (S * F), "gpu_time_us_under_ncu": us,
           "note": "one launch under ncu --set full --clock-control none; algorithmic bytes per frame and stream = 16*W*H"}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))
    break
