#!/bin/bash
# A/B of variant libraries on the GPU box: ./profiles/tools/ab2.sh WORKLOAD name1 name2 ...  ("main" = product library)
# each variant: quick parity (3-D p=4 cases vs the oracle) + bench without the CPU / e2e legs
wl=$1; shift
mkdir -p gpurun_out
L=$PWD/flou.jl_b200/flou_b200
for n in "$@"; do
  lib=$L/libflou_b200_x_$n.so; [ "$n" = main ] && lib=$L/libflou_b200.so
  echo "== $n"
  FLOU_B200_LIB=$lib timeout 300 python profiles/tools/quick_parity.py 2>&1 | tail -3
  FLOU_B200_LIB=$lib timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/ab_${wl}_$n.json 2> gpurun_out/ab_${wl}_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${wl}_$n.json")); r=d["roofline"]
    print("$wl $n", "GDOF/s=%.2f"%(d["value"]/1e9), "stage_ms=%.3f"%r.get("stage_ms",0), r.get("kernels_per_stage"), d["clocks"])
except Exception as e:
    print("$wl $n FAILED", e); print(open("gpurun_out/ab_${wl}_$n.err").read()[-600:])
PY
done
