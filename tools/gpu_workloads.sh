#!/bin/bash
# The other BASELINE.json configs through bench.py on one GPU:  gpurun --timeout 1800 -- 'bash tools/gpu_workloads.sh'
mkdir -p gpurun_out
for w in ${WORKLOADS:-curriculum encoder bf16}; do
  steps=3; [ "$w" = curriculum ] && steps=2
  timeout 900 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_n1.log 2>&1
  echo "== $w exit $?"
  tail -n 1 gpurun_out/bench_${w}_n1.log | python -c "
import sys, json
l = json.loads(sys.stdin.read())
print(' value', round(l['value']), 'ms', round(l['ms_per_step'], 2), 'e2e', round(l['e2e']['value']), 'parity', l['parity']['ok'], l['parity']['overlap'],
      'step_frac', round(l['roofline']['step_frac'], 3), 'extras', {k: l.get(k) for k in ('e2e_with_run_file',)})" || tail -n 20 gpurun_out/bench_${w}_n1.log
done
