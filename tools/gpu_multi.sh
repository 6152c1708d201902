#!/bin/bash
# Multi-GPU validation: NCCL test (2 ranks) + bench at N GPUs.  Run with: gpurun --gpus N -- 'N=... bash tools/gpu_multi.sh'
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus.txt
timeout 1200 python -m pytest tests/test_gpu_search.py -m gpu -q -k "torchrun or multi_gpu" --timeout 900 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1
echo "== pytest multi exit $?"; tail -n 8 gpurun_out/pytest_multi.log | cut -c1-300
for n in ${BENCH_NS:-$N}; do
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_n$n.log 2>&1
  fi
  echo "== bench n=$n exit $?"; tail -n 2 gpurun_out/bench_n$n.log | cut -c1-2400
done
