#!/bin/bash
# Round 2, 8-GPU check of the shipped build (gpurun --gpus 8), launched as the driver launches it.
O=gpurun_out/r2mg8_final; mkdir -p $O
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_n8.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("N=8", "GDOF/s=%.2f"%(d["value"]/1e9), "e2e=%.2f"%(d["e2e"]["value"]/1e9), "stage_ms=%.4f"%r.get("stage_ms",0), "bitwise", d.get("multi_gpu_bitwise"), d.get("multi_gpu_check"), d["clocks"])
except Exception as e:
    print("N=8 FAILED", e); print(open("$O/bench_n8.err").read()[-1500:])
PY
echo "total $(( $(date +%s) - t0 )) s"
