#!/bin/bash
# A/B of two builds of libcldrd.so inside one GPU session (alternating, to cancel thermal drift)
A=${A:-cl-drd_b200/cldrd/libcldrd_prev.so}
B=${B:-cl-drd_b200/cldrd/libcldrd.so}
for i in 1 2 3; do
  for L in $A $B; do
    echo -n "$(basename $L): "
    CLDRD_LIB_PATH=$PWD/$L timeout 300 python tools/tune_scan.py --settings ${SETTINGS:-0:0:0:1} 2>&1 | tail -1 | cut -c1-200
  done
done
