#!/bin/bash
# Round 2, tenth GPU call: face slots in blocks of 256 master elements (x-, y-, z-faces of a block
# side by side) and the face kernel walking them from the last to the first, against round-1 order.
O=gpurun_out/r2j; mkdir -p $O
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
echo "t=$(( $(date +%s) - t0 )) s"
bench() {  # workload name env...
  wl=$1; name=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-check > $O/ab_${wl}_$name.json 2> $O/ab_${wl}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/ab_${wl}_$name.json")); r=d["roofline"]
    print("$wl $name", "GDOF/s=%.2f"%(d["value"]/1e9), "stage_ms=%.4f"%r.get("stage_ms",0), "stage_frac=%.3f"%r["stage_frac"], r.get("kernels_per_stage"), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$wl $name FAILED", e); print(open("$O/ab_${wl}_$name.err").read()[-800:])
PY
}
for wl in cfg3 cfg3b cfg2 cfg5b cfg4; do
  bench $wl old FLOU_B200_FACE_CHUNK=0 FLOU_B200_FACE_REVERSE=0
  bench $wl chunk FLOU_B200_FACE_CHUNK=256 FLOU_B200_FACE_REVERSE=0
  bench $wl both FLOU_B200_FACE_CHUNK=256 FLOU_B200_FACE_REVERSE=1
  bench $wl rev FLOU_B200_FACE_CHUNK=0 FLOU_B200_FACE_REVERSE=1
done
echo "total $(( $(date +%s) - t0 )) s"
