#!/bin/bash
# GPU visit: batched-path parity tests, then c3 with the threshold mode / sample size varied (env knobs of engine.cu).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_sharded.py tests/test_gpu_pairs.py -x -q 2>&1 | tail -4
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c3_$label.json 2> gpurun_out/c3_$label.err
  python - "$label" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/c3_{sys.argv[1]}.json"))
    print(sys.argv[1], "q/s", round(j["value"]), "ms/batch", round(j["ms_per_step"], 4), "coarse ms", round(j["roofline"]["coarse_ms_per_batch"], 4),
          "TF/s", round(j["roofline"]["achieved"], 1), "e2e", round(j["e2e"]["value"]), j["batch_stats"], j["clocks"]["sm_mhz"])
except Exception as ex:
    print(sys.argv[1], "no result", ex)
PY
}
run guaranteed SVSB_BATCH_GUARANTEED=1
run stat_default SVSB_BATCH_GUARANTEED=0
run stat_s8k SVSB_BATCH_SAMPLE_ROWS=8192
run stat_s16k SVSB_BATCH_SAMPLE_ROWS=16384
run stat_s64k SVSB_BATCH_SAMPLE_ROWS=65536
run stat_default2 SVSB_BATCH_GUARANTEED=0
