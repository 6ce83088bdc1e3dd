#!/bin/bash
# Round 1, final single-GPU evidence run (repo root on the GPU box)
mkdir -p gpurun_out/r1c
O=gpurun_out/r1c
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench rc=$?"; cut -c1-400 $O/bench_cfg4.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
for wl in cfg1 cfg2 cfg3 cfg4s cfg5; do python bench.py --workload $wl --no-cpu-baseline > $O/bench_$wl.json 2> $O/bench_$wl.err; python -c "
import json; d=json.load(open('$O/bench_$wl.json')); print('$wl', '%.3g'%d['value'], d['roofline'].get('stage_frac'), d['e2e']['value'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg4_bench_steps2_warmup1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:line_kernel_ws --launch-skip 12 --launch-count 1 -f -o $O/lk_cfg4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_lk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:face_flux_kernel --launch-skip 12 --launch-count 1 -f -o $O/ff_cfg4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_ff.log 2>&1
ls -la $O; echo "total $(( $(date +%s) - t0 )) s"
