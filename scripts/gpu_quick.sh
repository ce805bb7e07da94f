#!/bin/bash
# One-GPU visit: full GPU test suite, then short bench lines for c2 / c1 / c3 / c5.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in c2 c1 c3 c5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err; echo "$w rc=$?"
  python - "$w" <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(f"gpurun_out/q_{sys.argv[1]}.json") if l.startswith("{")][-1])
    print(sys.argv[1], "q/s", round(j["value"], 1), "ms/q", round(j["ms_per_query"], 5), "roofline", round(j["roofline"]["achieved"], 1), round(j["roofline"]["frac"], 4),
          "e2e", round(j["e2e"]["value"], 1), "lat", j.get("latency_ms"), j.get("batch_stats"), j["clocks"]["sm_mhz"])
except Exception as ex:
    print("no result", ex)
PY
done
