#!/bin/bash
# N-GPU visit for the batched workload only: per-phase timing of a sharded batch (global thresholds, then per-rank
# thresholds with SVSB_BATCH_GLOBAL=0) and the c3 bench leg under torchrun.
set -u
N=${1:-2}
mkdir -p gpurun_out
export SVSB_XCHG_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
F='^\*\|OMP_NUM\|^$'
timeout 300 $TR --master-port 29641 scripts/c3_phases.py 2>&1 | grep -v "$F" | tee gpurun_out/r2_c3_phases_n$N.txt
SVSB_BATCH_GLOBAL=0 timeout 300 $TR --master-port 29642 scripts/c3_phases.py 2>&1 | grep -v "$F" | tee -a gpurun_out/r2_c3_phases_n$N.txt
timeout 600 $TR --master-port 29631 bench.py --gpus $N --workload c3 --only --steps 20 --warmup 3 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; echo "c3 bench rc=$?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads([l for l in open(f"gpurun_out/r2_bench_c3_n{n}.json") if l.startswith("{")][-1])
    print(round(j["value"]), "q/s", round(j["ms_per_step"] * 1e3, 1), "us/batch; e2e", round(j["e2e"]["value"]), "parity", j["parity"]["checked"], j["parity"]["exact"], j.get("batch_stats"))
except Exception as ex:
    print("no line", ex)
PY
