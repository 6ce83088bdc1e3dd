#!/bin/bash
# Round 2, call 17: compute-sanitizer on the shipped build, including the paths added late in the
# round (elected-lane copies, stage without x-face trace array, slab-wise stage, snapshot projection).
O=gpurun_out/r2q; mkdir -p $O
t0=$(date +%s)
run() {  # name tool env... -- command
  name=$1; tool=$2; shift; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 compute-sanitizer --tool $tool --print-limit 10 "$@" > $O/sanitizer_${tool}_$name.log 2>&1
  echo "$tool $name rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/sanitizer_${tool}_$name.log | tail -1) $(grep -E 'OK$|passed|failed' $O/sanitizer_${tool}_$name.log | tail -1)"
}
run np5 memcheck X=1 -- python profiles/tools/mid_parity.py 8 5
run np5 racecheck X=1 -- python profiles/tools/mid_parity.py 8 5
run np5 synccheck X=1 -- python profiles/tools/mid_parity.py 8 5
run np5_noxtrace memcheck FLOU_B200_XTRACE=0 -- python profiles/tools/mid_parity.py 8 5
run np5_noxtrace racecheck FLOU_B200_XTRACE=0 -- python profiles/tools/mid_parity.py 8 5
run np5_xtrace racecheck FLOU_B200_XTRACE=1 -- python profiles/tools/mid_parity.py 8 5
run np4_slab memcheck FLOU_B200_SLAB=162 -- python profiles/tools/mid_parity.py 9 4
run io memcheck X=1 -- python -m pytest tests/test_io.py -m gpu -q -x
run unstructured_source memcheck X=1 -- python -m pytest tests/test_unstructured.py tests/test_source_bc_gpu.py -m gpu -q -x
echo "total $(( $(date +%s) - t0 )) s"
