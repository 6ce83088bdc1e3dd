#!/bin/bash
# Round 2, multi-GPU check (gpurun --gpus 2): bitwise partition test, graph-captured halo exchange,
# bench at N=2 with the multi_gpu_bitwise self-check, A/B against direct launches.
O=gpurun_out/r2mg; mkdir -p $O
t0=$(date +%s)
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -s > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?"; grep -E "multigpu\]|passed|failed" $O/pytest_multigpu.log | tail -12
echo "t=$(( $(date +%s) - t0 )) s"
run() {  # name env...
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $O/bench_n2_$name.json 2> $O/bench_n2_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_n2_$name.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("$name", "GDOF/s=%.2f"%(d["value"]/1e9), "stage_ms=%.4f"%r.get("stage_ms",0), "bitwise", d.get("multi_gpu_bitwise"), d.get("multi_gpu_check"), d["clocks"])
except Exception as e:
    print("$name FAILED", e); print(open("$O/bench_n2_$name.err").read()[-1500:])
PY
}
run graph FLOU_B200_MG_GRAPH=1
run direct FLOU_B200_MG_GRAPH=0
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/bench_n1.json 2> $O/bench_n1.err; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('n1 GDOF/s=%.2f'%(d['value']/1e9))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>$O/bench_reference.err; cut -c1-200 $O/bench_reference.json
echo "total $(( $(date +%s) - t0 )) s"
