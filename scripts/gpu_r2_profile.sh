#!/bin/bash
# Round-2 profile visit (1 GPU): bench lines (ours with every config attached + the reference arm), drop-in benches, then the
# ncu launch lists and `--set full` captures.  Every ncu pass runs only after the same command has exited 0 without ncu.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
timeout 300 python scripts/c1_dropin_bench.py > gpurun_out/r2_c1_dropin.json 2> gpurun_out/r2_c1_dropin.err; echo "c1 dropin rc=$?"; tail -c 400 gpurun_out/r2_c1_dropin.json
timeout 120 python scripts/first_query_probe.py > gpurun_out/r2_first_query_probe.txt 2>&1; tail -4 gpurun_out/r2_first_query_probe.txt
timeout 120 python scripts/latency_ab.py > gpurun_out/r2_latency_ab.txt 2>&1; tail -3 gpurun_out/r2_latency_ab.txt
CMD="python bench.py --only --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 256 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_tma_kernel -s 200 -c 2 -f -o gpurun_out/prof_gemv $CMD > gpurun_out/ncu_gemv.log 2>&1; echo "ncu gemv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_topk -s 200 -c 2 -f -o gpurun_out/prof_select $CMD > gpurun_out/ncu_select.log 2>&1; echo "ncu select rc=$?"
CMDM="python scripts/mq_profile.py"
timeout 300 $CMDM > gpurun_out/plain_mq.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemv_tma_mq|select_topk" -s 4 -c 4 -f -o gpurun_out/prof_mq $CMDM > gpurun_out/ncu_mq.log 2>&1
echo "ncu mq rc=$?"
CMD3="python bench.py --workload c3 --only --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD3 > gpurun_out/plain_c3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 12 --csv --log-file gpurun_out/launches_c3.csv $CMD3 > gpurun_out/ncu_list_c3.log 2>&1
echo "ncu list c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coarse_gemm_kernel -s 6 -c 2 -f -o gpurun_out/prof_coarse $CMD3 > gpurun_out/ncu_coarse.log 2>&1; echo "ncu coarse rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"refine_kernel|sample_threshold" -s 6 -c 2 -f -o gpurun_out/prof_refine $CMD3 > gpurun_out/ncu_refine.log 2>&1; echo "ncu refine rc=$?"
