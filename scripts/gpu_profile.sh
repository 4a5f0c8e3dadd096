#!/bin/bash
# Profiling pass on the GPU box (run through gpurun from the repo root): development sweeps of the library's switches on the
# bench workload, the ncu launch list of the bench command and one `ncu --set full` capture per hot kernel IN STEADY STATE
# (a launch of a timed-equivalent window: the skip counts below step over the map pre-pass, the seeding window and two
# warm-up windows of `bench.py --steps 2 --warmup 2`).   gpurun --timeout 1800 -- 'bash scripts/gpu_profile.sh r2c'
TAG=${1:-dev}
OUT=gpurun_out
mkdir -p $OUT
set -x
export BENCH_QUICK=1
if [ "${SKIP_SWEEP:-0}" != "1" ]; then
  for V in "X=1" "MOVFE_EXTRACT_GROUPS=2" "MOVFE_EXTRACT_GROUPS=4" "MOVFE_PDL=1" "MOVFE_PDL=2" "BENCH_GRID_MODE=1" "BENCH_NO_POSE=1" ${EXTRA_SWEEP}; do
    echo "== $V" >> $OUT/sweep_$TAG.log
    env $V timeout 600 python bench.py --steps 6 --warmup 3 >> $OUT/sweep_$TAG.log 2>> $OUT/sweep_$TAG.err
  done
  cat $OUT/sweep_$TAG.log
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 2 > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
  for KS in ${NCU_KERNELS:-cand_lane_kernel:136 birth_lane_kernel:136 finalize_kernel:136 track_poses2_kernel:14 tp_prep_kernel:14 grid_kernel:8}; do
    K=${KS%%:*}; SKIP=${KS##*:}
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $SKIP -c 1 \
        -o $OUT/${K}_$TAG -f python bench.py --steps 2 --warmup 2 > $OUT/ncu_${K}_$TAG.log 2>&1; echo "ncu $K rc=$?"
  done
fi
