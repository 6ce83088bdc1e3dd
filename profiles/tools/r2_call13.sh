#!/bin/bash
# Round 2, call 13: update warp relieved -- copies issued by one elected lane (straight-line
# uniform-datapath code), freeP released and tmp of the next group requested before the trace pass,
# no x-face trace array with collocated nodes (FLOU_B200_XTRACE=1 keeps it) -- against the r2g build.
O=gpurun_out/r2m; mkdir -p $O
t0=$(date +%s)
L=$PWD/flou.jl_b200/flou_b200
python profiles/tools/mid_parity.py 8 5 2>&1 | tail -1
python profiles/tools/mid_parity.py 9 4 2>&1 | tail -1
timeout 900 python -m pytest tests/test_parity_production_gpu.py tests/test_unstructured.py tests/test_source_bc_gpu.py tests/test_multi_gpu.py -m gpu -q -x > $O/pytest_subset.log 2>&1; echo "pytest subset rc=$?"; tail -3 $O/pytest_subset.log
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
for wl in cfg4 cfg3 cfg2; do
  bench $wl base FLOU_B200_LIB=$L/libflou_b200_base.so
  bench $wl new_xtr FLOU_B200_XTRACE=1
  bench $wl new FLOU_B200_XTRACE=0
done
bench cfg4 base2 FLOU_B200_LIB=$L/libflou_b200_base.so
bench cfg4 new2 FLOU_B200_XTRACE=0
echo "t=$(( $(date +%s) - t0 )) s"
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python profiles/tools/mid_parity.py 8 5 > $O/sanitizer_racecheck_np5.log 2>&1; grep -E "RACECHECK SUMMARY|OK|FAIL" $O/sanitizer_racecheck_np5.log | tail -3
ncu --set full --clock-control none --import-source on -k regex:line_kernel_ws --launch-skip 12 --launch-count 1 -f -o $O/lk_cfg4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_lk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:face_flux_kernel --launch-skip 12 --launch-count 1 -f -o $O/ff_cfg4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_ff.log 2>&1
echo "total $(( $(date +%s) - t0 )) s"
