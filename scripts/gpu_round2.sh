#!/bin/bash
# One GPU-box visit (1 GPU): parity tests, bench lines (ours + reference arm), selection phases, ncu launch lists and
# full captures of every hot kernel.  Each ncu pass runs only after the same command has exited 0 without ncu.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
for w in c1 c5 c4; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
timeout 300 python scripts/c1_dropin_bench.py > gpurun_out/c1_dropin.json 2> gpurun_out/c1_dropin.err; echo "c1 dropin rc=$?"
timeout 120 python scripts/latency_ab.py > gpurun_out/latency_ab.txt 2>&1; tail -2 gpurun_out/latency_ab.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 120 python scripts/select_phases.py > gpurun_out/select_phases.log 2>&1; tail -2 gpurun_out/select_phases.log
timeout 200 python scripts/shard_probe.py > gpurun_out/shard_probe.txt 2>&1; tail -3 gpurun_out/shard_probe.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 256 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_tma -s 200 -c 2 -f -o gpurun_out/prof_gemv $CMD > gpurun_out/ncu_gemv.log 2>&1; echo "ncu gemv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_topk -s 200 -c 2 -f -o gpurun_out/prof_select $CMD > gpurun_out/ncu_select.log 2>&1; echo "ncu select rc=$?"
CMD3="python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD3 > gpurun_out/plain_c3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 12 --csv --log-file gpurun_out/launches_c3.csv $CMD3 > gpurun_out/ncu_list_c3.log 2>&1
echo "ncu list c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coarse_gemm_kernel -s 6 -c 2 -f -o gpurun_out/prof_coarse $CMD3 > gpurun_out/ncu_coarse.log 2>&1; echo "ncu coarse rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"refine_kernel|sample_threshold" -s 6 -c 2 -f -o gpurun_out/prof_refine $CMD3 > gpurun_out/ncu_refine.log 2>&1; echo "ncu refine rc=$?"
CMDP="python scripts/peer_profile.py"
timeout 300 $CMDP > gpurun_out/plain_peer.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"merge_window|select_topk" -s 40 -c 4 -f -o gpurun_out/prof_peer $CMDP > gpurun_out/ncu_peer.log 2>&1
echo "ncu peer rc=$?"
