#!/bin/sh
# Run on the GPU box (gpurun): round-2 ncu evidence.  sh tools/profile_step_r2.sh
#   launches_r2.csv      every launch of one default bench step with its device time
#   step_r2.ncu-rep      --set full of every kernel of one step
#   fast_natural_r2      --set full of k_fast on camera-like frames
#   tc_r2                --set full of the tensor-core matcher (75776 x 125000)
set -x
CMD="python bench.py --no-cpu --no-hamming --no-fundamental --no-loop --no-natural --no-cfg3 --no-single --no-sustained --steps 2 --warmup 1"
$CMD > gpurun_out/plain_r2.log 2>&1 || exit 1
python tools/fast_ncu.py natural > gpurun_out/plain_fast_r2.log 2>&1 || exit 1
python tools/tc_ncu.py > gpurun_out/plain_tc_r2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 15 -c 45 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_launches_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 15 -c 15 -o gpurun_out/step_r2 -f $CMD > gpurun_out/ncu_step_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fast -s 2 -c 1 -o gpurun_out/fast_natural_r2 -f python tools/fast_ncu.py natural > gpurun_out/ncu_fast_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hamming_tc -s 3 -c 1 -o gpurun_out/tc_r2 -f python tools/tc_ncu.py > gpurun_out/ncu_tc_r2.log 2>&1
tail -2 gpurun_out/ncu_step_r2.log gpurun_out/ncu_fast_r2.log gpurun_out/ncu_tc_r2.log | cut -c1-200
