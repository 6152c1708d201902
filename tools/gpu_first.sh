#!/bin/bash
# First GPU bring-up: diagnostics per scan mode (each in its own process: a trapped kernel kills
# only that context), the GPU test-suite, one bench line.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
for s in simt f16 bf16 tf32; do
  timeout 400 python tools/gpu_diag.py --scan $s > gpurun_out/diag_$s.log 2>&1
  echo "== diag $s exit $?"; tail -n 12 gpurun_out/diag_$s.log | cut -c1-600
done
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest exit $?"; tail -n 40 gpurun_out/pytest_gpu.log | cut -c1-400
for s in ${BENCH_SCANS:-f16}; do
  timeout 900 python bench.py --steps 3 --warmup 3 --scan $s > gpurun_out/bench_$s.log 2>&1
  echo "== bench $s exit $?"; tail -n 3 gpurun_out/bench_$s.log | cut -c1-3000
done
