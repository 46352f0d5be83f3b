for p in 8 2 16 32; do echo "pieces $p"; RTKV_HOST_PIECES=$p RTKV_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/trace_$p.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['e2e']['ms_per_step'])"; grep "rtkv host" gpurun_out/trace_$p.log | tail -3; done
