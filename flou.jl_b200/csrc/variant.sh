#!/bin/bash
# Development helper: build a kernel variant next to the product library,
#   ./variant.sh NAME "PAIRS" "EXTRA nvcc flags"   ->  ../flou_b200/libflou_b200_x_NAME.so
# select it at run time with FLOU_B200_LIB=<path> (see flou_b200/_lib.py).
set -e
name=$1; pairs=${2:-3_5}; extra=$3
make -j8 BUILD=build/x_$name OUT=../flou_b200/libflou_b200_x_$name.so PAIRS="$pairs" EXTRA="$extra" 2>&1 \
  | grep -E "error|line_kernel.*LCfgILi3ELi5ELi1ELi2ELb1" -A2 | grep -E "error|registers|spill" || true
