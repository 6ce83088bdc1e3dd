#!/bin/bash
# Round 2, call 22: the new default of large single-GPU states (no x-face trace array + face slots in
# blocks of 4096 master elements) -- full-size parity test and the default bench on the committed build.
O=gpurun_out/r2v; mkdir -p $O
t0=$(date +%s)
FLOU_B200_XTRACE=0 FLOU_B200_FACE_CHUNK=16 python profiles/tools/mid_parity.py 8 5 2>&1 | tail -1
timeout 400 python -m pytest tests/test_parity_fullsize_gpu.py -m gpu -q -x > $O/pytest_fullsize.log 2>&1; echo "pytest fullsize rc=$?"; tail -1 $O/pytest_fullsize.log
echo "t=$(( $(date +%s) - t0 )) s"
python bench.py --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/bench_cfg4.json')); r=d['roofline']; print('cfg4', 'GDOF/s=%.2f'%(d['value']/1e9), 'stage_ms=%.3f'%r['stage_ms'], 'stage_frac=%.3f'%r['stage_frac'], 'frac=%.3f'%r['frac'], 'e2e=%.2f'%(d['e2e']['value']/1e9), 'parity', d['parity']['ok'], r['kernels_per_stage'], d['clocks'])"
echo "total $(( $(date +%s) - t0 )) s"
