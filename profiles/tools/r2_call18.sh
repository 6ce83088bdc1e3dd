#!/bin/bash
# Round 2, call 18: CUDA graphs of both ping-pong parities built (and uploaded) at the first use, from
# two steps on: no graph instantiation inside a later (timed) call.  Full GPU suite + every workload.
O=gpurun_out/r2r; mkdir -p $O
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
echo "t=$(( $(date +%s) - t0 )) s"
python bench.py > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench rc=$?"
for wl in cfg4 cfg1 cfg2 cfg3 cfg3b cfg4s cfg5 cfg5b; do [ $wl = cfg4 ] || python bench.py --workload $wl --no-cpu-baseline --no-check > $O/bench_$wl.json 2> $O/bench_$wl.err; python -c "
import json; d=json.load(open('$O/bench_$wl.json')); r=d['roofline']; print('$wl', 'GDOF/s=%.2f'%(d['value']/1e9), 'ms/step=%.4f'%d['ms_per_step'], 'stage_frac=%.3f'%r['stage_frac'], 'e2e=%.2f'%(d['e2e']['value']/1e9), 'launches', d['gpu_launches'], r.get('kernels_per_stage'), d['clocks'].get('sm_mhz'))"; done
echo "total $(( $(date +%s) - t0 )) s"
