#!/bin/bash
# Round 1, third GPU session, call 1: GPU test-suite, then A/B runs (run from the repo root on the GPU box).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -5 gpurun_out/pytest_gpu.log
run() {   # tag workload [env assignments...]
  tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/ab_${wl}_$tag.json 2> gpurun_out/ab_${wl}_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${wl}_$tag.json")); r=d["roofline"]
    print("$wl $tag", "GDOF/s=%.2f"%(d["value"]/1e9), "stage_ms=%.3f"%r.get("stage_ms",0), r.get("kernels_per_stage"), d["clocks"])
except Exception as e:
    print("$wl $tag FAILED", e); print(open("gpurun_out/ab_${wl}_$tag.err").read()[-600:])
PY
}
L=$PWD/flou.jl_b200/flou_b200
run main    cfg4s FLOU_B200_LIB=$L/libflou_b200.so
run fp      cfg4s FLOU_B200_LIB=$L/libflou_b200_x_fp.so
run l2f32   cfg4s FLOU_B200_LIB=$L/libflou_b200.so FLOU_B200_L2_FETCH=32
run l2f128  cfg4s FLOU_B200_LIB=$L/libflou_b200.so FLOU_B200_L2_FETCH=128
run main2   cfg4s FLOU_B200_LIB=$L/libflou_b200.so
run main    cfg4  FLOU_B200_LIB=$L/libflou_b200.so
run l2f32   cfg4  FLOU_B200_LIB=$L/libflou_b200.so FLOU_B200_L2_FETCH=32
echo "total $(( $(date +%s) - t0 )) s"
