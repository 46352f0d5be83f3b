#!/bin/bash
# Round-2 ncu evidence (one GPU).  usage: bash tools/profile_round2.sh r02a
TAG=${1:-r02}
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --layers 8"
$B > gpurun_out/plain_bench_$TAG.json 2> gpurun_out/plain_bench_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prompt_mass|score_select|quant_gather|log_table" -s 72 -c 240 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu1.log 2>&1
$B > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'quant_gather_vec|prompt_mass_kernel|score_select' -s 72 -c 3 -o gpurun_out/prof_path_$TAG $B > gpurun_out/ncu2.log 2>&1
S="env WORLDS=8 python tools/bench_shard_stages.py"
$S > gpurun_out/plain_shard_$TAG.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'shard_stats_staged|quant_gather_flat|shard_combine' -s 8 -c 3 -o gpurun_out/prof_shard_$TAG $S > gpurun_out/ncu3.log 2>&1
F="python tools/profile_flash.py"
$F > gpurun_out/plain_flash_$TAG.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flash_prefill -s 2 -c 1 -o gpurun_out/prof_flash_$TAG $F > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log gpurun_out/ncu4.log
ls -la gpurun_out/*$TAG*
