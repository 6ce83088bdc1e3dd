#!/bin/bash
# Round 2, evidence run of the committed build on one B200 (repo root on the GPU box)
O=gpurun_out/r2_final2; mkdir -p $O
t0=$(date +%s)
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
echo "t=$(( $(date +%s) - t0 )) s"
python bench.py > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$O/bench_cfg4.json")); r=d["roofline"]
print("cfg4 GDOF/s=%.2f stage_ms=%.3f stage_frac=%.3f frac=%.3f fp64_frac=%.3f e2e=%.2f cpu=%.1fM parity_err=%.1e launches=%d"%(d["value"]/1e9, r["stage_ms"], r["stage_frac"], r["frac"], r.get("fp64_frac",0), d["e2e"]["value"]/1e9, d["cpu_baseline"]["value"]/1e6, d.get("parity_err",-1), d["gpu_launches"]), r["kernels_per_stage"], d["clocks"])
PY
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-160 $O/bench_reference.json
echo "t=$(( $(date +%s) - t0 )) s"
for wl in cfg1 cfg2 cfg3 cfg3b cfg4s cfg5 cfg5b; do python bench.py --workload $wl --no-cpu-baseline --no-check > $O/bench_$wl.json 2> $O/bench_$wl.err; python -c "
import json; d=json.load(open('$O/bench_$wl.json')); r=d['roofline']; print('$wl', 'GDOF/s=%.2f'%(d['value']/1e9), 'stage_frac=%.3f'%r['stage_frac'], 'e2e=%.2f'%(d['e2e']['value']/1e9), r.get('kernels_per_stage'))"; done
echo "t=$(( $(date +%s) - t0 )) s"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg4_bench_steps2_warmup1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:line_kernel_ws --launch-skip 12 --launch-count 1 -f -o $O/lk_cfg4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_lk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:face_flux_kernel --launch-skip 12 --launch-count 1 -f -o $O/ff_cfg4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_ff.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:line_kernel_ws --launch-skip 12 --launch-count 1 -f -o $O/lk_cfg3 python bench.py --workload cfg3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $O/ncu_lk3.log 2>&1
ls -la $O | head -40; echo "total $(( $(date +%s) - t0 )) s"
