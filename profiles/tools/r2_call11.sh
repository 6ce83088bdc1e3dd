#!/bin/bash
# Round 2: config 4 at full size against the oracle (periodic tiling), timing of the new tests
O=gpurun_out/r2k; mkdir -p $O
t0=$(date +%s)
timeout 1500 python -m pytest tests/test_parity_fullsize_gpu.py -m gpu -q --durations=5 -x > $O/pytest_fullsize.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_fullsize.log
free -g | head -2
echo "total $(( $(date +%s) - t0 )) s"
