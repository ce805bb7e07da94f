set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_batch.py -x -q 2>&1 | tail -4
{
timeout 200 python scripts/c3_virtual_ranks.py
SVSB_REFINE=split timeout 200 python scripts/c3_virtual_ranks.py
for n in 125000 1000000; do
  timeout 120 python scripts/c3_breakdown.py $n 768 100 1024 40
  SVSB_REFINE=split timeout 120 python scripts/c3_breakdown.py $n 768 100 1024 40
done
timeout 120 python scripts/c3_breakdown.py 1000000 3072 1000 256 10
SVSB_REFINE=split timeout 120 python scripts/c3_breakdown.py 1000000 3072 1000 256 10
} > gpurun_out/c3_virtual.txt 2>&1; cat gpurun_out/c3_virtual.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_virtual_launches.csv python scripts/c3_virtual_ranks.py 1000000 768 100 1024 8 1 > gpurun_out/c3_virtual_ncu.log 2>&1; echo ncu rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_bd_launches.csv python scripts/c3_breakdown.py 1000000 768 100 1024 1 > gpurun_out/c3_bd_ncu.log 2>&1; echo ncu rc=$?
