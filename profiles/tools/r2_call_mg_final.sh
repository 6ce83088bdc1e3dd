#!/bin/bash
# Round 2, multi-GPU check of the shipped build (gpurun --gpus 2): bitwise partition tests, bench at
# N=2 as the driver launches it (multi_gpu_bitwise self-check, e2e), reference arm under torchrun.
O=gpurun_out/r2mg_final; mkdir -p $O
t0=$(date +%s)
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -s > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?"; grep -E "multigpu\]|passed|failed" $O/pytest_multigpu.log | tail -12
echo "t=$(( $(date +%s) - t0 )) s"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_n2.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("N=2", "GDOF/s=%.2f"%(d["value"]/1e9), "e2e=%.2f"%(d["e2e"]["value"]/1e9), "stage_ms=%.4f"%r.get("stage_ms",0), "bitwise", d.get("multi_gpu_bitwise"), d.get("multi_gpu_check"), d["clocks"])
except Exception as e:
    print("N=2 FAILED", e); print(open("$O/bench_n2.err").read()[-1500:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_reference_n2.json 2>$O/bench_reference_n2.err; cut -c1-260 $O/bench_reference_n2.json
echo "total $(( $(date +%s) - t0 )) s"
