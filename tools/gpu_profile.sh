#!/bin/bash
# ncu evidence for the scan kernel: (1) launch list of a short bench run, (2) --set full on the two
# biggest scan launches of a warm search.  Numbers printed under ncu are never bench values.
SCAN=${SCAN:-f16}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$SCAN.csv \
    python bench.py --steps 1 --warmup 3 --scan $SCAN --no-cpu-baseline > gpurun_out/bench_under_ncu_$SCAN.log 2>&1
echo "== launch list exit $?"; tail -n 2 gpurun_out/bench_under_ncu_$SCAN.log | cut -c1-400
# second search = launches 9.. of the scan kernel (9 chunks per search); take the last two (largest)
ncu --set full --clock-control none --import-source on -k regex:scan_tc -s ${SKIP:-16} -c 2 \
    -o gpurun_out/prof_$SCAN -f python tools/profile_scan.py --scan $SCAN > gpurun_out/profile_$SCAN.log 2>&1
echo "== ncu full exit $?"; tail -n 5 gpurun_out/profile_$SCAN.log | cut -c1-600
ls -la gpurun_out
