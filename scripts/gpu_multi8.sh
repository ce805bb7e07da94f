#!/bin/bash
# 8-GPU visit: peer-exchange parity, c2 (both exchange modes) and c4 under torchrun.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export SVSB_XCHG_TIMEOUT_MS=10000
timeout 300 $TR --master-port 29611 scripts/sharded_check.py 2>&1 | tail -3
show() {
  python - "$1" <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1], "q/s", round(j["value"], 1), "ms/q", round(j["ms_per_query"], 4), "gemv GB/s", round(j["roofline"]["achieved"]), "e2e", round(j["e2e"]["value"], 1), "launches", j["gpu_launches"], j["clocks"]["sm_mhz"])
except Exception as ex:
    print(sys.argv[1], "no result", ex)
PY
}
for ex in ${EXCHANGES:-peer collective}; do
  timeout 300 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 3 --exchange $ex > gpurun_out/bench_c2_n${N}_$ex.json 2> gpurun_out/bench_c2_n${N}_$ex.err; echo "c2 n$N $ex rc=$?"
  show gpurun_out/bench_c2_n${N}_$ex.json
done
[ "${LIGHT:-0}" = "1" ] && exit 0
timeout 400 $TR --master-port 29613 bench.py --gpus $N --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4_n${N}_peer.json 2> gpurun_out/bench_c4_n${N}_peer.err; echo "c4 rc=$?"
show gpurun_out/bench_c4_n${N}_peer.json
timeout 300 $TR --master-port 29614 bench.py --gpus $N --workload c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_n${N}.json 2> gpurun_out/bench_c3_n${N}.err; echo "c3 rc=$?"
show gpurun_out/bench_c3_n${N}.json
