#!/bin/bash
# Batched pipeline per shard size, new (split refine, fast order statistic) against old (SVSB_REFINE_FUSED / SVSB_SAMPLE_GENERIC).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -5
{
for n in 125000 1000000; do
  timeout 120 python scripts/c3_breakdown.py $n 768 100 1024 40
  SVSB_REFINE_FUSED=1 SVSB_SAMPLE_GENERIC=1 timeout 120 python scripts/c3_breakdown.py $n 768 100 1024 40
done
} > gpurun_out/c3_breakdown.txt 2>&1
for sp in 2 4; do
SVSB_RESCORE_SPLIT=$sp timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_bd_launches_$sp.csv python scripts/c3_breakdown.py 1000000 768 100 1024 1 > gpurun_out/c3_bd_ncu.log 2>&1
done
echo rc=$?
cat gpurun_out/c3_breakdown.txt
