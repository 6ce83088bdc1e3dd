#!/bin/bash
# Round 2, call 23: the example script end to end (monitor + save callback + stage limiter).
O=gpurun_out/r2w; mkdir -p $O
timeout 100 python examples/tgv3d.py 8 3 10 $O/tgv > $O/example.log 2>&1; echo "example rc=$?"; tail -8 $O/example.log; ls $O | head
