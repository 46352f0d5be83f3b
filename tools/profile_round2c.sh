#!/bin/bash
# Round-2 (c) ncu evidence for the two tensor-core kernels that changed after r02b (one GPU).  usage: bash tools/profile_round2c.sh r02c
TAG=${1:-r02c}
F="python tools/profile_flash.py"
$F > gpurun_out/plain_flash_$TAG.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flash_prefill -s 2 -c 1 -o gpurun_out/prof_flash_$TAG $F > gpurun_out/ncu4.log 2>&1
K="python tools/bench_kernels.py c3 --only k1b"
$K > gpurun_out/plain_k1b_$TAG.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:prompt_mass_qk -s 3 -c 1 -o gpurun_out/prof_k1b_$TAG $K > gpurun_out/ncu5.log 2>&1
tail -2 gpurun_out/ncu4.log gpurun_out/ncu5.log
cat gpurun_out/plain_k1b_$TAG.txt | tail -5
python tools/mma_rate.py > gpurun_out/mma_rate_$TAG.txt 2>&1; tail -8 gpurun_out/mma_rate_$TAG.txt
python tools/profile_flash_phases.py 8192 0 > gpurun_out/flash_phases_v0_$TAG.json 2>&1
python tools/profile_flash_phases.py 8192 4 > gpurun_out/flash_phases_v4_$TAG.json 2>&1
python tools/bench_flash.py > gpurun_out/bench_flash_$TAG.txt 2>&1; tail -5 gpurun_out/bench_flash_$TAG.txt | cut -c1-400
ls -la gpurun_out/*$TAG*
