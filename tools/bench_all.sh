#!/bin/bash
# All BASELINE.json workloads that fit one GPU, device-resident timing only (gpurun).  One JSON summary line per run.
sumline() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1', 'ms/step', round(d['ms_per_step'],3), 'tok/s', round(d['value']), 'K3 in-situ', round(r['frac'],3), 'whole', round(r['whole_path']['frac'],3), 'rho', round(r['retained_fraction'],3), 'layers', d['config']['layers'])"; }
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
$B --workload c1 2>/dev/null | tail -1 | sumline c1
$B --workload c3 --layers 8 2>/dev/null | tail -1 | sumline c3
for bits in 8,8,8 4,4,4 2,2,2 2,4,8; do $B --workload c4 --layers 10 --bits $bits 2>/dev/null | tail -1 | sumline c4_$bits; done
$B --workload c5 --layers 8 2>/dev/null | tail -1 | sumline c5
$B --workload c5 --layers 8 --ablation quant_only 2>/dev/null | tail -1 | sumline c5_quant_only
