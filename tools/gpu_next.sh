#!/bin/bash
# First GPU call of the next round: what the last round could not measure any more.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_next.sh'
# 1. the 2-GPU torchrun parity test (both transports + the shared host result block)
# 2. bench at 8 GPUs with phase times: `e2e` now goes through ShardedSearcher.search_host (every rank copies its
#    slice into one shared page-locked block); expected ~11.9 ms per search end to end instead of 13.4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -q -k "torchrun or multi_gpu" --timeout 800 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1
echo "== pytest multi exit $?"; tail -n 4 gpurun_out/pytest_multi.log | cut -c1-300
for n in ${BENCH_NS:-8}; do
  CLDRD_DIST_PROFILE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.log 2>&1
  echo "== bench n=$n exit $?"
  tail -n 1 gpurun_out/bench_n$n.log | python -c "
import sys, json
l = json.loads(sys.stdin.read())
print(l['n_gpus'], 'value', round(l['value']), 'ms', round(l['ms_per_step'], 2), 'e2e', round(l['e2e']['value']), 'e2e ms', round(l['e2e']['ms_per_step'], 2), l.get('phase_ms_last_step'))"
done
