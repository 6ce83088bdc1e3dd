#!/bin/bash
# Round 2, 8-GPU check (gpurun --gpus 8): graph-captured halo exchange at N=8 with the bitwise
# self-check, against direct launches.
O=gpurun_out/r2mg8; mkdir -p $O
t0=$(date +%s)
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -s > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?"; grep -E "multigpu\]|passed|failed" $O/pytest_multigpu.log | tail -8
echo "t=$(( $(date +%s) - t0 )) s"
run() {  # name env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > $O/bench_n8_$name.json 2> $O/bench_n8_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_n8_$name.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("$name", "GDOF/s=%.2f"%(d["value"]/1e9), "stage_ms=%.4f"%r.get("stage_ms",0), "bitwise", d.get("multi_gpu_bitwise"), d.get("multi_gpu_check"), d["clocks"])
except Exception as e:
    print("$name FAILED", e); print(open("$O/bench_n8_$name.err").read()[-1500:])
PY
}
run graph FLOU_B200_MG_GRAPH=1
echo "t=$(( $(date +%s) - t0 )) s"
run direct FLOU_B200_MG_GRAPH=0
echo "total $(( $(date +%s) - t0 )) s"
