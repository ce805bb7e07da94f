#!/bin/bash
# One GPU-box visit: parity tests, bench (ours + reference arm), selection phases, ncu launch list and full captures.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
cat gpurun_out/bench_c2.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 120 python scripts/select_phases.py > gpurun_out/select_phases.log 2>&1; tail -4 gpurun_out/select_phases.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 256 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_tma -s 200 -c 2 -f -o gpurun_out/prof_gemv $CMD > gpurun_out/ncu_gemv.log 2>&1
echo "ncu gemv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_topk -s 200 -c 2 -f -o gpurun_out/prof_select $CMD > gpurun_out/ncu_select.log 2>&1
echo "ncu select rc=$?"
