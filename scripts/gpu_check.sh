#!/bin/bash
# One GPU-box pass: parity tests, the bench (both arms), the ncu launch list of the bench command and one
# `ncu --set full` capture per hot kernel. Run through gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh r1z'
# Everything lands in gpurun_out/ (scratch); summaries worth judging are copied to profiles/ afterwards
# (scripts/launch_summary.py, scripts/ncu_keys.py, scripts/ncu_traffic.py).
TAG=${1:-dev}
OUT=gpurun_out
mkdir -p $OUT
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
tail -3 $OUT/bench_$TAG.err; cat $OUT/bench_$TAG.json
if [ "${SKIP_REF:-0}" != "1" ]; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
  cat $OUT/bench_ref_$TAG.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  # launch list of the bench command (cold-cache, serialised: shares only); the plain run above exited 0 first
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 1 > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
  for K in ${NCU_KERNELS:-grid_kernel cand_kernel finalize_kernel track_poses_kernel}; do
    STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip ${NCU_SKIP:-1} -c 1 \
        -o $OUT/${K}_$TAG -f python scripts/bench_probe.py > $OUT/ncu_${K}_$TAG.log 2>&1; echo "ncu $K rc=$?"
  done
fi
