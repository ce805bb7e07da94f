#!/bin/bash
# N-GPU visit for the batched workload only: c3 under torchrun with the global-threshold path and, for comparison, with
# per-rank thresholds (SVSB_BATCH_GLOBAL=0).
set -u
N=${1:-2}
mkdir -p gpurun_out
export SVSB_XCHG_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29631 bench.py --gpus $N --workload c3 --only --steps 20 --warmup 3 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; echo "c3 global rc=$?"; tail -c 600 gpurun_out/r2_bench_c3_n$N.err
SVSB_BATCH_GLOBAL=0 timeout 600 $TR --master-port 29632 bench.py --gpus $N --workload c3 --only --steps 20 --warmup 3 > gpurun_out/r2_bench_c3_n${N}_local.json 2> gpurun_out/r2_bench_c3_n${N}_local.err; echo "c3 local rc=$?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for tag in ("", "_local"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/r2_bench_c3_n{n}{tag}.json") if l.startswith("{")][-1])
        print(tag or "global", round(j["value"]), "q/s", round(j["ms_per_step"] * 1e3, 1), "us/batch; e2e", round(j["e2e"]["value"]), "parity", j["parity"], j.get("batch_stats"))
    except Exception as ex:
        print("no line", tag, ex)
PY
