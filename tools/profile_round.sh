#!/bin/bash
# Run on the GPU box (gpurun): plain runs first, then the ncu passes of the SAME commands.
set -u
TAG=${1:-r01}
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > $OUT/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prompt_mass|score_select|quant_gather" -s 192 -c 192 \
    --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launches_${TAG}.log 2>&1
$CMD > $OUT/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"prompt_mass|score_select|quant_gather" -s 288 -c 6 \
    -o $OUT/prof_${TAG} $CMD > $OUT/ncu_full_${TAG}.log 2>&1
QK="python tools/bench_kernels.py c3 --only k1b"
$QK > $OUT/plain_qk_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"prompt_mass_qk" -s 10 -c 2 \
    -o $OUT/prof_qk_${TAG} $QK > $OUT/ncu_qk_${TAG}.log 2>&1
EX="python tools/profile_ex.py"
$EX > $OUT/plain_ex_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"quant_" -s 9 -c 9 \
    -o $OUT/prof_ex_${TAG} $EX > $OUT/ncu_ex_${TAG}.log 2>&1
for f in $OUT/ncu_launches_${TAG}.log $OUT/ncu_full_${TAG}.log $OUT/ncu_qk_${TAG}.log; do tail -n 2 $f | cut -c1-200; done
