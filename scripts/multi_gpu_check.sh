#!/bin/bash
# N-GPU pass of the partitioned CG: parity tests, then the X bench in peer-mapped and NCCL mode.
# Usage (under gpurun --gpus N): bash scripts/multi_gpu_check.sh N tag
N=${1:-2}; TAG=${2:-r02}; OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q > $OUT/${TAG}_n${N}_pytest.log 2>&1; tail -3 $OUT/${TAG}_n${N}_pytest.log
for MODE in 1 0; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 2 --warmup 2 --skip-cpu --skip-native --cg-p2p $MODE > $OUT/${TAG}_n${N}_p2p${MODE}_dev.json 2> $OUT/${TAG}_n${N}_p2p${MODE}_dev.err
  echo "bench N=$N p2p=$MODE exit $?"; tail -2 $OUT/${TAG}_n${N}_p2p${MODE}_dev.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_n${N}_p2p${MODE}_dev.json").read().strip().splitlines()[-1])
    print("N=$N p2p=$MODE value %.2f GDoF/s ms/step %.1f its %s mode %s iter_ms %.4f" % (d["value"], d["ms_per_step"], d["sizes"]["cg_iterations_per_step"], d["sizes"].get("comm_mode"), d["roofline"]["cg_iteration"]["ms"]))
except Exception as e: print("no line", e)
PY
done
