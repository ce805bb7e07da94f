#!/bin/bash
# Selection-kernel phase times and synchronous latency as a function of the rows-per-group knob.
for sh in 6 7 8 9; do
  echo "== SVSB_GROUP_SHIFT_MIN=$sh (rows per group $((1 << sh)))"
  SVSB_GROUP_SHIFT_MIN=$sh timeout 100 python scripts/select_phases.py 1000000 1536 100 2>&1 | tail -2
  SVSB_GROUP_SHIFT_MIN=$sh timeout 100 python scripts/select_phases.py 125000 1536 100 2>&1 | tail -1
  SVSB_GROUP_SHIFT_MIN=$sh timeout 100 python scripts/select_phases.py 1000000 3072 1000 2>&1 | tail -1
done
SVSB_GROUP_SHIFT_MIN=8 timeout 300 python -m pytest tests/test_gpu_select.py tests/test_gpu_query.py -x -q 2>&1 | tail -2
