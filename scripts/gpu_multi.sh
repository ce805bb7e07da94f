#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): peer-exchange parity under torchrun, then c2 with both exchange modes.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SVSB_XCHG_TIMEOUT_MS=10000 timeout 300 $TR --master-port 29611 scripts/sharded_check.py 2>&1 | tail -5
for ex in peer collective; do
  SVSB_XCHG_TIMEOUT_MS=10000 timeout 300 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 3 --exchange $ex > gpurun_out/bench_c2_n${N}_$ex.json 2> gpurun_out/bench_c2_n${N}_$ex.err; echo "c2 n$N $ex rc=$?"
  python - "$N" "$ex" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/bench_c2_n{sys.argv[1]}_{sys.argv[2]}.json"))
    print(sys.argv[2], "q/s", round(j["value"], 1), "ms/q", round(j["ms_per_query"], 4), "gemv GB/s", round(j["roofline"]["achieved"]), "e2e", round(j["e2e"]["value"], 1), "launches", j["gpu_launches"])
except Exception as ex:
    print("no result", ex)
PY
done
