#!/bin/bash
# One-GPU validation call:  gpurun --timeout 1500 -- 'bash tools/gpu_1gpu.sh'
# the whole -m gpu suite (incl. the node-wide protocol on one GPU: 3 in-process shards, 2 processes over gloo + CUDA IPC),
# then the default bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 -p no:cacheprovider ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest exit $?"; tail -n 5 gpurun_out/pytest_gpu.log | cut -c1-400
if [ -z "$SKIP_BENCH" ]; then
  timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/bench_n1.log 2>&1
  echo "== bench exit $?"
  tail -n 1 gpurun_out/bench_n1.log | cut -c1-3000
fi
