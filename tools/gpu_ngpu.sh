#!/bin/bash
# Multi-GPU validation call:  gpurun --gpus N --timeout 1200 -- 'N=2 bash tools/gpu_ngpu.sh'
# 1. the multi-GPU tests (2-rank NCCL worker: peer-memory exchange and NCCL transport, shared host block; the CLI under
#    torchrun; in-process shards over two devices)   2. bench at N GPUs (parity block inside, phase times in the line)
N=${N:-2}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -q -k "torchrun or multi_gpu or gloo" --timeout 800 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1
  echo "== pytest multi exit $?"; tail -n 4 gpurun_out/pytest_multi.log | cut -c1-300
fi
for n in ${BENCH_NS:-$N}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_n$n.log 2>&1
  echo "== bench n=$n exit $?"
  tail -n 1 gpurun_out/bench_n$n.log | python -c "
import sys, json
l = json.loads(sys.stdin.read())
print(l['n_gpus'], 'value', round(l['value']), 'ms', round(l['ms_per_step'], 2), 'e2e', round(l['e2e']['value']), 'e2e ms', round(l['e2e']['ms_per_step'], 2))
print(' roofline', {k: round(v, 3) for k, v in l['roofline'].items() if k in ('frac', 'step_frac', 'e2e_frac', 'scan_ms_per_step', 'scan_share_of_step')})
print(' phases', l.get('phase_ms_last_batch'))
for r, p in enumerate(l.get('phase_ms_mean_by_rank') or []): print('   rank', r, p)
print(' parity', l['parity'])
print(' extras', {k: l.get(k) for k in ('writer', 'index_load', 'e2e_inprocess', 'e2e_with_run_file')})" || tail -n 30 gpurun_out/bench_n$n.log
done
