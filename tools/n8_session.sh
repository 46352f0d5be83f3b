#!/bin/bash
# One multi-GPU session: head-sharded bench variants on N GPUs (8 layers each: per-layer times are what matter)
N=${1:-8}
run() { # name, extra env, extra args
  env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus $N --steps 20 --warmup 5 --no-secondary --layers 8 $3 > gpurun_out/n${N}_$1.json 2> gpurun_out/n${N}_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/n${N}_$1.json").read().strip().splitlines()[-1])
    print("$1", "ms/step", round(d["ms_per_step"], 4), "us/layer", round(d["ms_per_step"] * 1e3 / 8, 1), "xerr", d.get("exchange_error"))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/n${N}_$1.err").read()[-1500:])
PY
}
run fused "X=1" ""
run nostores "RTKV_SHARD_VARIANT=3" ""
run generic "RTKV_SHARD_VARIANT=1" ""
run collective "X=1" "--shard-mode collective"
