#!/bin/bash
# Round 2, first GPU call: full GPU suite on the new build (TMA flux blocks), measured fp64 peak,
# A/B against the round-1 kernels, cfg3 group-size / occupancy variants, launch list.
O=gpurun_out/r2a; mkdir -p $O
L=$PWD/flou.jl_b200/flou_b200
t0=$(date +%s)
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_peak profiles/tools/dfma_peak.cu && /tmp/dfma_peak > $O/fp64_peak.json; cat $O/fp64_peak.json
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 $O/pytest_gpu.log
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
bench cfg4 main $L/libflou_b200.so
bench cfg4 r1 $L/libflou_b200_x_r1.so
echo "t=$(( $(date +%s) - t0 )) s"
for v in c3a c3b c3c c3d; do FLOU_B200_LIB=$L/libflou_b200_x_$v.so timeout 300 python profiles/tools/mid_parity.py 12 4 2>&1 | tail -1; done
bench cfg3 main $L/libflou_b200.so
bench cfg3 r1 $L/libflou_b200_x_r1.so
for v in c3a c3b c3c c3d; do bench cfg3 $v $L/libflou_b200_x_$v.so; done
for v in main c3b c3c; do lib=$L/libflou_b200_x_$v.so; [ $v = main ] && lib=$L/libflou_b200.so; bench cfg3b $v $lib; done
echo "t=$(( $(date +%s) - t0 )) s"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg4_bench_steps2_warmup1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:line_kernel_ws --launch-skip 12 --launch-count 1 -f -o $O/lk_cfg4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_lk.log 2>&1
ls -la $O | head -40; echo "total $(( $(date +%s) - t0 )) s"
