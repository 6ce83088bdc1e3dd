#!/bin/bash
# Round 2: compute-sanitizer on the production path (TMA / mbarrier / cp.async hand-offs)
O=gpurun_out/r2l; mkdir -p $O
t0=$(date +%s)
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/mid_parity.py 8 5 > $O/sanitizer_${tool}_np5.log 2>&1; echo "$tool np5 rc=$?"; grep -E "ERROR SUMMARY|OK|FAIL|hazard|Invalid" $O/sanitizer_${tool}_np5.log | tail -4
  echo "t=$(( $(date +%s) - t0 )) s"
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/tools/mid_parity.py 9 4 > $O/sanitizer_memcheck_np4.log 2>&1; echo "memcheck np4 rc=$?"; grep -E "ERROR SUMMARY|OK|FAIL" $O/sanitizer_memcheck_np4.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_unstructured.py tests/test_source_bc_gpu.py -m gpu -q -x > $O/sanitizer_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitizer_memcheck_tests.log | tail -3
echo "total $(( $(date +%s) - t0 )) s"
