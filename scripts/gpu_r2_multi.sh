#!/bin/bash
# Round-2 N-GPU visit (default 2): the GPU suite with N real devices visible (tests/test_gpu_multi.py then uses them),
# the torchrun parity check (log kept), and the one-line bench at N (SPMD legs + the in-process leg).
set -u
N=${1:-2}
mkdir -p gpurun_out
export SVSB_XCHG_TIMEOUT_MS=10000
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_gpu_n$N.log
fi
timeout 600 $TR --master-port 29611 scripts/sharded_check.py > gpurun_out/r2_sharded_check_n$N.log 2>&1; echo "sharded_check rc=$?"; tail -3 gpurun_out/r2_sharded_check_n$N.log
if [ "${SOAK:-0}" != "0" ]; then
  timeout 900 $TR --master-port 29621 scripts/peer_soak.py $SOAK > gpurun_out/r2_peer_soak_n$N.log 2>&1; echo "soak rc=$?"; tail -2 gpurun_out/r2_peer_soak_n$N.log
fi
timeout 1200 $TR --master-port 29612 bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc=$?"; tail -c 1500 gpurun_out/r2_bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads([l for l in open(f"gpurun_out/r2_bench_n{n}.json") if l.startswith("{")][-1])
    def show(tag, v):
        print(tag, v.get("error") or (round(v["value"], 1), "e2e", round(v["e2e"]["value"], 1), "sync", round(v["e2e"].get("one_query_in_flight", 0), 1),
              "frac", round(v["roofline"]["frac"], 3), "parity", v["parity"].get("checked"), v["parity"].get("exact"), v["parity"].get("max_rel_err")))
    show("headline", j)
    if "inprocess" in j: show("inprocess", j["inprocess"])
    for c, v in j.get("configs", {}).items(): show(c, v)
except Exception as ex:
    print("no bench line", ex)
PY
