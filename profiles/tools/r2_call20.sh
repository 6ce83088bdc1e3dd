#!/bin/bash
# Round 2, last call: the committed build as the driver will run it -- GPU suite, smoke, default bench,
# reference arm.
O=gpurun_out/r2t; mkdir -p $O
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
echo "t=$(( $(date +%s) - t0 )) s"
python bench.py > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/bench_cfg4.json')); r=d['roofline']; print('cfg4', 'GDOF/s=%.2f'%(d['value']/1e9), 'stage_frac=%.3f'%r['stage_frac'], 'frac=%.3f'%r['frac'], 'fp64_frac=%.3f'%r['fp64_frac'], 'e2e=%.2f'%(d['e2e']['value']/1e9), 'cpu=%.1fM'%(d['cpu_baseline']['value']/1e6), 'parity', d['parity']['ok'], d['parity']['rhs_err'], 'launches', d['gpu_launches'], d['clocks'])"
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-200 $O/bench_reference.json
echo "total $(( $(date +%s) - t0 )) s"
