#!/bin/bash
# Round 2, call 16: size-based default of the x-face trace array (none while the state fits in L2):
# full GPU suite (both branches in the parity tests), smoke, the L2-resident workloads.
O=gpurun_out/r2p; mkdir -p $O
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --durations=3 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
echo "t=$(( $(date +%s) - t0 )) s"
for wl in cfg1 cfg2 cfg3 cfg5 cfg5b; do for x in default 1; do
  if [ $x = default ]; then unset FLOU_B200_XTRACE; else export FLOU_B200_XTRACE=1; fi
  python bench.py --workload $wl --no-cpu-baseline --no-check > $O/bench_${wl}_x$x.json 2> $O/bench_${wl}_x$x.err; python -c "
import json; d=json.load(open('$O/bench_${wl}_x$x.json')); r=d['roofline']; print('$wl xtrace=$x', 'GDOF/s=%.2f'%(d['value']/1e9), 'stage_frac=%.3f'%r['stage_frac'], 'e2e=%.2f'%(d['e2e']['value']/1e9), r.get('kernels_per_stage'))"
done; done
unset FLOU_B200_XTRACE
echo "total $(( $(date +%s) - t0 )) s"
