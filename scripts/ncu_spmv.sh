#!/bin/bash
# ncu --set full capture of one SpMV launch for each kernel variant given (on the X mesh, level 2)
OUT=gpurun_out
mkdir -p $OUT
TAG=$1; shift
for K in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_spmv_(stream|window|jds)' -s 5 -c 1 -f -o $OUT/${TAG}_k${K} \
      env SWEEP_NATIVE=0 python scripts/spmv_sweep.py 2 $K 0 > $OUT/${TAG}_k${K}.log 2>&1; echo "ncu k$K exit $?"
  tail -2 $OUT/${TAG}_k${K}.log
done
