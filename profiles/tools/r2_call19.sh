#!/bin/bash
# Round 2, call 19: programmatic dependent launch of the stage kernels (FLOU_B200_PDL=1) against
# normal launches, inside the captured graphs; parity suite with it on.
O=gpurun_out/r2s; mkdir -p $O
t0=$(date +%s)
FLOU_B200_PDL=1 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_production_gpu.py tests/test_monitors_gpu.py tests/test_source_bc_gpu.py -m gpu -q -x > $O/pytest_pdl.log 2>&1; echo "pytest PDL rc=$?"; tail -2 $O/pytest_pdl.log
echo "t=$(( $(date +%s) - t0 )) s"
for wl in cfg1 cfg2 cfg3 cfg5 cfg5b cfg3b cfg4; do for x in 0 1 0 1; do
  FLOU_B200_PDL=$x python bench.py --workload $wl --no-cpu-baseline --no-check --no-e2e > $O/bench_${wl}_pdl$x.json 2> $O/bench_${wl}_pdl$x.err; python -c "
import json; d=json.load(open('$O/bench_${wl}_pdl$x.json')); r=d['roofline']; print('$wl pdl=$x', 'GDOF/s=%.2f'%(d['value']/1e9), 'ms/step=%.4f'%d['ms_per_step'], 'stage_frac=%.3f'%r['stage_frac'], d['clocks'].get('sm_mhz'))"
done; done
echo "total $(( $(date +%s) - t0 )) s"
