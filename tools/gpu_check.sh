#!/bin/bash
# GPU regression + bench: the GPU test-suite, then one bench line per scan mode in $BENCH_SCANS.
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest exit $?"; tail -n ${PYTEST_TAIL:-15} gpurun_out/pytest_gpu.log | cut -c1-300
fi
for s in ${BENCH_SCANS:-f16}; do
  timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 --scan $s ${BENCH_ARGS} > gpurun_out/bench_$s.log 2>&1
  echo "== bench $s exit $?"; tail -n 2 gpurun_out/bench_$s.log | cut -c1-2600
done
