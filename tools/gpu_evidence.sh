#!/bin/bash
# One-GPU evidence call:  gpurun --timeout 2400 -- 'bash tools/gpu_evidence.sh'
#   1. compute-sanitizer memcheck + racecheck on the small config (tools/sanitize.py)
#   2. ncu launch list of the bench command (per-launch durations; cold-cache, serialised: shares, not absolutes)
#   3. ncu --set full capture of the dominant launch (full-index filter scan) and of the re-score kernel
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for tool in memcheck racecheck; do
  SAN_ROWS=${SAN_ROWS:-6000} timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== compute-sanitizer $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run ok|ok=" gpurun_out/sanitizer_$tool.log | tail -n 8
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu launch list exit $?"; wc -l gpurun_out/r02_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"scan_tc2_kernel|rescore_sort_kernel|select_merge_kernel" \
  --launch-skip 9 --launch-count 4 -o gpurun_out/r02_prof_main -f python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline \
  > gpurun_out/profile_main.log 2>&1
echo "== ncu full exit $?"; ls -la gpurun_out/r02_prof_main.ncu-rep
