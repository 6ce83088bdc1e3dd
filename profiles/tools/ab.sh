#!/bin/bash
# A/B helper (run on the GPU box): ./profiles/tools/ab.sh WORKLOAD name1 name2 ...
# times bench.py with flou.jl_b200/flou_b200/libflou_b200_x_<name>.so ("main" = product library)
wl=$1; shift
for n in "$@"; do
  lib=flou.jl_b200/flou_b200/libflou_b200_x_$n.so
  [ "$n" = main ] && lib=flou.jl_b200/flou_b200/libflou_b200.so
  [ "$n" = node ] && { lib=flou.jl_b200/flou_b200/libflou_b200.so; export FLOU_B200_NODE_KERNEL=1; } || unset FLOU_B200_NODE_KERNEL
  FLOU_B200_LIB=$PWD/$lib python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_${wl}_$n.json 2> gpurun_out/ab_${wl}_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${wl}_$n.json")); r=d["roofline"]
    print("$wl $n", "GDOF/s=%.2f"%(d["value"]/1e9), "stage_ms=%.3f"%r.get("stage_ms",0), r.get("kernels_per_stage"), d["config"]["launch"])
except Exception as e:
    print("$wl $n FAILED", e); print(open("gpurun_out/ab_${wl}_$n.err").read()[-600:])
PY
done
