#!/bin/bash
# Round 2, call 21: no x-face trace array TOGETHER with face slots blocked by master element
# (x-, y-, z-faces of a block of elements evaluated close in time: the y / z node layers are then L2
# hits behind the x-face reads of the same elements).
O=gpurun_out/r2u; mkdir -p $O
t0=$(date +%s)
bench() {  # workload name env...
  wl=$1; name=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-check > $O/ab_${wl}_$name.json 2> $O/ab_${wl}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/ab_${wl}_$name.json")); r=d["roofline"]
    print("$wl $name", "GDOF/s=%.2f"%(d["value"]/1e9), "stage_ms=%.3f"%r["stage_ms"], "stage_frac=%.3f"%r["stage_frac"], r.get("kernels_per_stage"), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$wl $name FAILED", e); print(open("$O/ab_${wl}_$name.err").read()[-800:])
PY
}
bench cfg4 base X=1
for c in 32 128 256 1024 4096 16384; do
  bench cfg4 nox_chunk$c FLOU_B200_XTRACE=0 FLOU_B200_FACE_CHUNK=$c
done
bench cfg4 nox_chunk256_fwd FLOU_B200_XTRACE=0 FLOU_B200_FACE_CHUNK=256 FLOU_B200_FACE_REVERSE=0
bench cfg4 base2 X=1
echo "t=$(( $(date +%s) - t0 )) s"
FLOU_B200_XTRACE=0 FLOU_B200_FACE_CHUNK=256 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:face_flux_kernel --launch-skip 12 --launch-count 1 --csv --log-file $O/ff_nox_chunk256.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_ff.log 2>&1
grep -E "dram__bytes|gpu__time" $O/ff_nox_chunk256.csv | cut -d, -f 5,12- | head -5
echo "total $(( $(date +%s) - t0 )) s"
