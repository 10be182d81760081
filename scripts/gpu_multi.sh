#!/bin/bash
# Multi-GPU pass on one box: the partitioned-solve tests, then the benchmark under torchrun for each N given.
# Usage: bash scripts/gpu_multi.sh <tag> <levels> N [N ...]
TAG=$1; LEVELS=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 1200 python bench.py --levels $LEVELS --skip-native --skip-cpu > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus $N --levels $LEVELS > $OUT/${TAG}_n$N.json 2> $OUT/${TAG}_n$N.err
  fi
  echo "N=$N exit $?"; tail -4 $OUT/${TAG}_n$N.err; cat $OUT/${TAG}_n$N.json | cut -c1-1800
done
