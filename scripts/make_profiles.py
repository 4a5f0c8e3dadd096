#!/usr/bin/env python
"""Summaries of one GPU pass (scripts/gpu_check.sh + scripts/gpu_profile.sh, tag TAG) -> profiles/ (tracked):
  bench_TAG.json, bench_ref_TAG.json      the bench lines of the pass
  launches_TAG.csv / launches_TAG.md      ncu launch list of the bench command and its per-kernel summary
  ncu_keys_TAG.txt                        key metrics of one steady-state `ncu --set full` capture per hot kernel
  ncu_hot_lines_TAG.txt                   hottest source lines of the same captures
  kernel_traffic_r2.json                  dram__bytes_read + dram__bytes_write per launch of every captured kernel and per step of
                                          every stage (read by bench.py for roofline.traffic)
  sass_TAG_<kernel>.txt                   SASS of the hot kernels of the build the captures were taken from
usage: python scripts/make_profiles.py TAG [frames per step = 16] [propagation chains = 3: a launch covers 1/chains of the streams]"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1]
F = int(sys.argv[2]) if len(sys.argv) > 2 else 16
G = int(sys.argv[3]) if len(sys.argv) > 3 else 3
# kernel -> (translation unit, stage, launches per step)
KERNELS = {"cand_lane_kernel": ("extract", "extract", F * G), "birth_lane_kernel": ("extract", "extract", F * G), "finalize_kernel": ("extract", "extract", F * G),
           "track_poses2_kernel": ("pose", "pose", F // 4), "tp_prep_kernel": ("pose", "pose", F // 4), "grid_kernel": ("grid", "grid", 1)}
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw).stdout


def main():
    for name in ("bench_%s.json" % TAG, "bench_ref_%s.json" % TAG, "launches_%s.csv" % TAG):
        if os.path.exists(os.path.join(OUT, name)):
            shutil.copy(os.path.join(OUT, name), os.path.join(PROF, name))
    if os.path.exists(os.path.join(OUT, "launches_%s.csv" % TAG)):
        open(os.path.join(PROF, "launches_%s.md" % TAG), "w").write(
            run([sys.executable, os.path.join(ROOT, "scripts", "launch_summary.py"), os.path.join(OUT, "launches_%s.csv" % TAG)]))
    keys, lines, traffic = [], [], {"tag": TAG, "kernels": {}, "note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE steady-state launch "
                                    "(ncu --set full --clock-control none; scripts/gpu_profile.sh), and per step of a stage = sum over its kernels "
                                    "x launches per step"}
    for k, (unit, stage, per_step) in KERNELS.items():
        rep = os.path.join(OUT, "%s_%s.ncu-rep" % (k, TAG))
        if not os.path.exists(rep):
            continue
        keys.append(run([sys.executable, os.path.join(ROOT, "scripts", "ncu_keys.py"), rep]))
        lines.append("==== %s\n" % k + run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, unit, k, "25"]))
        rows = list(csv.reader(run(['ncu', '-i', rep, '--page', 'raw', '--csv']).splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = dict(zip(hdr, vals))

        def get(name):
            return float(d[name].replace(',', '')) * UNIT[units[hdr.index(name)]]
        rd, wr = get('dram__bytes_read.sum'), get('dram__bytes_write.sum')
        traffic["kernels"][k] = {"stage": stage, "launches_per_step": per_step, "dram_bytes_read": rd, "dram_bytes_write": wr,
                                 "dram_bytes_per_launch": rd + wr, "gpu_time_under_ncu": d['gpu__time_duration.sum'] + " " + units[hdr.index('gpu__time_duration.sum')],
                                 "grid": d.get('launch__grid_size'), "warp_instructions": d.get('smsp__inst_executed.sum')}
        traffic[stage] = traffic.get(stage, 0.0) + (rd + wr) * per_step
        # SASS of the kernel (the 1024-pitch instantiation of the templated ones), without the encoding comments
        rx = k + ("ILi1024" if k in ("cand_lane_kernel", "birth_lane_kernel") else "ILi1" if k == "grid_kernel" else "")
        if stage == "extract":   # the three kernels of the dominant stage
            sass = run(["bash", os.path.join(ROOT, "scripts", "sass.sh"), unit, rx])
            open(os.path.join(PROF, "sass_%s_%s.txt" % (TAG, k)), "w").write(sass)
    open(os.path.join(PROF, "ncu_keys_%s.txt" % TAG), "w").write("\n".join(keys))
    open(os.path.join(PROF, "ncu_hot_lines_%s.txt" % TAG), "w").write("\n".join(lines))
    json.dump(traffic, open(os.path.join(PROF, "kernel_traffic_r2.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in traffic.items() if k != "kernels"}))


if __name__ == "__main__":
    main()
