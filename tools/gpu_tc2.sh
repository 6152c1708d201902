#!/bin/bash
# Validate the 2-CTA scan: quick diag, the GPU suite with CLDRD_TC2=1, then an A/B against the 1-CTA kernel.
mkdir -p gpurun_out
CLDRD_TC2=1 timeout 300 python tools/gpu_diag.py --scan f16 > gpurun_out/diag_tc2_f16.log 2>&1
echo "== diag tc2 f16 exit $?"; tail -n 9 gpurun_out/diag_tc2_f16.log | cut -c1-420
CLDRD_TC2=1 timeout 300 python tools/gpu_diag.py --scan tf32 > gpurun_out/diag_tc2_tf32.log 2>&1
echo "== diag tc2 tf32 exit $?"; tail -n 3 gpurun_out/diag_tc2_tf32.log | cut -c1-420
CLDRD_TC2=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_tc2.log 2>&1
echo "== pytest tc2 exit $?"; tail -n 12 gpurun_out/pytest_tc2.log | cut -c1-300
timeout 600 python tools/tune_scan.py 2>&1 | tail -5 | cut -c1-400
