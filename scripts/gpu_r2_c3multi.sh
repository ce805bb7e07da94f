#!/bin/bash
# N-GPU visit for the batched workload only: the c3 bench leg under torchrun -- fused peer exchange (default), then the
# collective form of the same protocol (exchange=collective), then per-rank thresholds (SVSB_BATCH_GLOBAL=0).
set -u
N=${1:-2}
mkdir -p gpurun_out
export SVSB_XCHG_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # tag, extra env / args
  tag=$1; shift
  env "$@" timeout 600 $TR --master-port 29631 bench.py --gpus $N --workload c3 --only --steps 20 --warmup 3 ${EXTRA:-} > gpurun_out/r2_bench_c3_n${N}$tag.json 2> gpurun_out/r2_bench_c3_n${N}$tag.err; echo "c3$tag rc=$?"
}
run "" SVSB_X=1
EXTRA="--exchange collective" run _collective SVSB_X=1
[ "${LOCAL:-0}" = "1" ] && run _local SVSB_BATCH_GLOBAL=0
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for tag in ("", "_collective", "_local"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/r2_bench_c3_n{n}{tag}.json") if l.startswith("{")][-1])
        print(tag or "peer", round(j["value"]), "q/s", round(j["ms_per_step"] * 1e3, 1), "us/batch; e2e", round(j["e2e"]["value"]), "parity", j["parity"]["checked"], j["parity"]["exact"], j.get("batch_stats"))
    except Exception as ex:
        print("no line", tag, ex)
PY
tail -c 800 gpurun_out/r2_bench_c3_n$N.err
