#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (ncu --set full). usage: python scripts/ncu_keys.py X.ncu-rep [more metrics]"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print('== %s  (id %s)' % (d.get('Kernel Name', '?')[:60], d.get('ID')))
        for w in WANT + extra:
            if w in d:
                print('  %-72s %16s %s' % (w, d[w], units[hdr.index(w)]))
        st = [(float(d[h].replace(',', '')), h) for h in hdr if h.startswith('smsp__average_warp') and h.endswith('_per_issue_active.ratio') or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'))]
        for v, h in sorted(st, reverse=True)[:8]:
            print('  stall %-66s %16.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
main()
