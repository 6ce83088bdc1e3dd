#!/bin/bash
# Round 2, ninth GPU call: TMA load issue spread over the line warps (variant li) against main
O=gpurun_out/r2i; mkdir -p $O
L=$PWD/flou.jl_b200/flou_b200
t0=$(date +%s)
for a in "12 5" "11 5" "12 4"; do FLOU_B200_LIB=$L/libflou_b200_x_li.so timeout 180 python profiles/tools/mid_parity.py $a 2>&1 | tail -1; done
echo "t=$(( $(date +%s) - t0 )) s"
bench() {  # workload name lib
  FLOU_B200_LIB=$3 timeout 600 python bench.py --workload $1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-check > $O/ab_$1_$2.json 2> $O/ab_$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open("$O/ab_$1_$2.json")); r=d["roofline"]
    print("$1 $2", "GDOF/s=%.2f"%(d["value"]/1e9), "stage_ms=%.4f"%r.get("stage_ms",0), "stage_frac=%.3f"%r["stage_frac"], r.get("kernels_per_stage"), d["clocks"], d["config"].get("launch"))
except Exception as e:
    print("$1 $2 FAILED", e); print(open("$O/ab_$1_$2.err").read()[-800:])
PY
}
bench cfg4 li $L/libflou_b200_x_li.so
bench cfg4 main $L/libflou_b200.so
bench cfg4 li2 $L/libflou_b200_x_li.so
bench cfg3 li $L/libflou_b200_x_li.so
bench cfg3 main $L/libflou_b200.so
echo "t=$(( $(date +%s) - t0 )) s"
FLOU_B200_LIB=$L/libflou_b200_x_li.so ncu --set full --clock-control none --import-source on -k regex:line_kernel_ws --launch-skip 12 --launch-count 1 -f -o $O/lk_cfg4_li python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_lk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:line_kernel_ws --launch-skip 12 --launch-count 1 -f -o $O/lk_cfg4_main python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_lk_main.log 2>&1
ls -la $O | head; echo "total $(( $(date +%s) - t0 )) s"
