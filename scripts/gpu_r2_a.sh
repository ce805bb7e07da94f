#!/bin/bash
# Round-2 visit A (1 GPU): parity tests, then the one-line bench with every BASELINE config attached.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2a_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
python - <<'PY'
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2a_bench.json") if l.startswith("{")][-1])
    print("c2", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), "parity", j["parity"]["checked"], j["parity"]["exact"], j["parity"]["max_rel_err"])
    for c, v in j.get("configs", {}).items():
        print(c, v.get("error") or (round(v["value"],1), "e2e", round(v["e2e"]["value"],1), "frac", round(v["roofline"]["frac"],3), "parity", v["parity"]["checked"], v["parity"]["exact"], v["parity"]["max_rel_err"]))
except Exception as ex:
    print("no bench line", ex)
PY
