#!/bin/bash
# same-box A/B of two libcldrd builds on the 128-query loop (tools/diag_b128.py), alternating
A=${A:-cl-drd_b200/cldrd/libcldrd_prev.so}
B=${B:-cl-drd_b200/cldrd/libcldrd.so}
for i in 1 2 3; do
  for L in $A $B; do
    echo -n "$(basename $L): "
    CLDRD_LIB_PATH=$PWD/$L timeout 300 python tools/diag_b128.py 2>&1 | tail -1 | cut -c1-260
  done
done
