#!/bin/bash
# GPU visit: other workloads on one GPU (c1, c5, c4) and the launch list of a c3 step.
set -u
mkdir -p gpurun_out
for w in c1 c5 c4; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
  python - "$w" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
    print(sys.argv[1], {k: j[k] for k in ("value", "ms_per_query")}, "gemv GB/s", j["roofline"]["achieved"], "frac", j["roofline"]["frac"], "e2e", j["e2e"]["value"], "lat", j.get("latency_ms"))
except Exception as ex:
    print("no result", ex)
PY
done
CMD="python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_c3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 12 --csv --log-file gpurun_out/launches_c3.csv $CMD > gpurun_out/ncu_list_c3.log 2>&1
echo "ncu list c3 rc=$?"
