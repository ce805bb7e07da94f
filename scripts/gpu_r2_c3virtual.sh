set -u
mkdir -p gpurun_out
timeout 200 python scripts/c3_virtual_ranks.py > gpurun_out/c3_virtual.txt 2>&1; echo rc=$?; tail -3 gpurun_out/c3_virtual.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_virtual_launches.csv python scripts/c3_virtual_ranks.py 1000000 768 100 1024 8 1 > gpurun_out/c3_virtual_ncu.log 2>&1; echo ncu rc=$?
