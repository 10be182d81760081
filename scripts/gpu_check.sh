#!/bin/bash
# One GPU-box pass: parity tests, the benchmark, the ncu launch list and one full capture of the SpMV kernel.
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc >> $OUT/${TAG}_smi.txt; lscpu | grep "Model name" >> $OUT/${TAG}_smi.txt
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -3 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 0 --skip-native --skip-cpu > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 4 -c 2 -f -o $OUT/${TAG}_spmv \
    python bench.py --steps 1 --warmup 0 --skip-native --skip-cpu > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT
fi
