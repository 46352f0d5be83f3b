#!/bin/bash
# quick head-sharded timing on N GPUs (8 layers): fused scheme only
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
  bench.py --gpus $N --steps 20 --warmup 5 --no-secondary --layers 8 $EXTRA > gpurun_out/q${N}.json 2> gpurun_out/q${N}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/q${N}.json").read().strip().splitlines()[-1])
    print("N=${N} fused: ms/step", round(d["ms_per_step"], 4), "us/layer", round(d["ms_per_step"] * 1e3 / 8, 1), "xerr", d.get("exchange_error"), "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/q${N}.err").read()[-1500:])
PY
