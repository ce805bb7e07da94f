#!/bin/bash
# GPU visit for the batched path: parity tests, c3 bench, optional ncu capture of the coarse kernel (NCU=1).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_c3.json"))
print({k: j[k] for k in ("value", "ms_per_step")}, j["roofline"]["achieved"], j["roofline"]["frac"], j["roofline"]["coarse_ms_per_batch"], j["e2e"]["value"], j["batch_stats"], j["clocks"])
PY
if [ "${NCU:-0}" = "1" ]; then
  CMD="python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline"
  timeout 300 $CMD > gpurun_out/plain_c3.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:coarse_gemm_kernel -s 6 -c 2 -f -o gpurun_out/prof_coarse $CMD > gpurun_out/ncu_coarse.log 2>&1
  echo "ncu coarse rc=$?"
fi
