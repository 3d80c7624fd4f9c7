#!/bin/sh
# Run on the GPU box (gpurun): launch list + one --set full capture of every kernel of one step of the default bench.
# Usage: sh tools/profile_step.sh <tag>
set -x
TAG=${1:-r1}
CMD="python bench.py --no-cpu --no-hamming --no-fundamental --steps 2 --warmup 1"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 15 -c 45 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 15 -c 15 -o gpurun_out/step_$TAG -f $CMD > gpurun_out/ncu_step_$TAG.log 2>&1
tail -2 gpurun_out/ncu_step_$TAG.log | cut -c1-200
