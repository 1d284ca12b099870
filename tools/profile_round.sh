#!/bin/bash
# ncu evidence for the bench's kernels (run under gpurun; ~2 GPU-minutes).  Usage: bash tools/profile_round.sh <tag>
# 1) plain bench runs must exit 0; 2) launch list of the same command; 3) `--set full` of one whole step per workload.
TAG=${1:-r01d}
mkdir -p gpurun_out
for W in cfg2 cfg3; do
  N=$([ $W = cfg2 ] && echo 16 || echo 9)   # launches of this library per step (incl. score_finalize)
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload $W > gpurun_out/plain_${TAG}_$W.log 2>&1 || { echo "plain $W failed"; tail -5 gpurun_out/plain_${TAG}_$W.log; continue; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_$W.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload $W > gpurun_out/ncu_${TAG}_a_$W.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_|convt|convlstm|score_finalize" -s $((3 * N)) -c $N -f -o gpurun_out/prof_${TAG}_$W \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload $W > gpurun_out/ncu_${TAG}_b_$W.log 2>&1
  tail -2 gpurun_out/ncu_${TAG}_b_$W.log | cut -c1-200
done
ls -la gpurun_out | grep ${TAG}
