#!/bin/bash
# Round 2, call 15: slab-wise stage (FLOU_B200_SLAB = elements per slab): face kernel and element
# kernel alternate slab by slab so that a slab's flux blocks are still in L2 when its elements run.
O=gpurun_out/r2o; mkdir -p $O
t0=$(date +%s)
FLOU_B200_SLAB=64 python profiles/tools/mid_parity.py 8 5 2>&1 | tail -1
FLOU_B200_SLAB=162 python profiles/tools/mid_parity.py 9 4 2>&1 | tail -1
FLOU_B200_SLAB=144 timeout 600 python -m pytest tests/test_parity_production_gpu.py -m gpu -q -x -k "12, 12, 12 or 24, 24, 24 or 23, 23, 25" > $O/pytest_slab.log 2>&1; echo "pytest slab rc=$?"; tail -2 $O/pytest_slab.log
echo "t=$(( $(date +%s) - t0 )) s"
bench() {  # workload name env...
  wl=$1; name=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-check > $O/ab_${wl}_$name.json 2> $O/ab_${wl}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$O/ab_${wl}_$name.json")); r=d["roofline"]
    print("$wl $name", "GDOF/s=%.2f"%(d["value"]/1e9), "ms/step=%.3f"%d["ms_per_step"], "stage_frac=%.3f"%r["stage_frac"], "launches", d["gpu_launches"], d["clocks"]["sm_mhz"], d["clocks"].get("power_w"))
except Exception as e:
    print("$wl $name FAILED", e); print(open("$O/ab_${wl}_$name.err").read()[-800:])
PY
}
bench cfg4 slab0 FLOU_B200_SLAB=0
bench cfg4 slab16384 FLOU_B200_SLAB=16384
bench cfg4 slab8192 FLOU_B200_SLAB=8192
bench cfg4 slab32768 FLOU_B200_SLAB=32768
bench cfg4 slab65536 FLOU_B200_SLAB=65536
bench cfg4 slab262144 FLOU_B200_SLAB=262144
bench cfg4 slab0b FLOU_B200_SLAB=0
bench cfg4 slab16384_nox FLOU_B200_SLAB=16384 FLOU_B200_XTRACE=0
echo "t=$(( $(date +%s) - t0 )) s"
# DRAM bytes of one RK stage, default against slab-wise (ncu, one pass, 3 metrics)
for sl in 0 16384; do
  n=$(( sl == 0 ? 2 : 256 ))
  FLOU_B200_SLAB=$sl ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --launch-skip $(( 3 + 6 * n )) --launch-count $n --csv --log-file $O/stage_dram_slab$sl.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_slab$sl.log 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("$O/stage_dram_slab$sl.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]; hdr=rows[h]
kn,mn,mu,mv=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Unit"),hdr.index("Metric Value")
tot={}
scale={"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9,"ns":1e-6,"us":1e-3,"ms":1,"nsecond":1e-6,"usecond":1e-3,"msecond":1}
for r in rows[h+1:]:
    if len(r)<=mv: continue
    k=r[kn].split("<")[0].split()[-1]+":"+r[mn]
    tot[k]=tot.get(k,0)+float(r[mv].replace(",",""))*scale.get(r[mu],1)
print("slab=$sl", {k:(round(v/1e9,2) if "bytes" in k else round(v,3)) for k,v in sorted(tot.items())})
PY
done
echo "total $(( $(date +%s) - t0 )) s"
