#!/bin/bash
# Final 1-GPU visit of round 2: whole GPU suite, smoke, the one-line bench (every BASELINE config attached), the reference
# arm, and the batched-path ncu captures (launch list + `--set full` of the kernels that changed in this round's second half).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2f_smoke.log
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2f_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"
CMD3="python bench.py --workload c3 --only --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD3 > gpurun_out/plain_c3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 12 --csv --log-file gpurun_out/launches_c3.csv $CMD3 > gpurun_out/ncu_list_c3.log 2>&1
echo "ncu list c3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"refine_lean|sample_order" -s 6 -c 2 -f -o gpurun_out/prof_refine $CMD3 > gpurun_out/ncu_refine.log 2>&1; echo "ncu refine rc=$?"
CMDV="python scripts/c3_virtual_ranks.py 1000000 768 100 1024 8 1"
timeout 300 $CMDV > gpurun_out/plain_virtual.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3_virtual_n8.csv $CMDV > gpurun_out/ncu_list_virtual.log 2>&1
echo "ncu list virtual rc=$?"
{
for n in 125000 1000000; do
  timeout 120 python scripts/c3_breakdown.py $n 768 100 1024 40
  SVSB_REFINE=split timeout 120 python scripts/c3_breakdown.py $n 768 100 1024 40
  SVSB_REFINE=fused SVSB_SAMPLE_GENERIC=1 timeout 120 python scripts/c3_breakdown.py $n 768 100 1024 40
done
timeout 200 python scripts/c3_virtual_ranks.py
} > gpurun_out/r2f_c3_breakdown.txt 2>&1
python - <<'PY'
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2f_bench_n1.json") if l.startswith("{")][-1])
    def show(tag, v):
        print(tag, v.get("error") or (round(v["value"], 1), "e2e", round(v["e2e"]["value"], 1), "frac", round(v["roofline"]["frac"], 3),
              "parity", v["parity"].get("checked"), v["parity"].get("exact")))
    show("headline", j)
    for c, v in j.get("configs", {}).items(): show(c, v)
    print("incremental", j.get("incremental_update"))
except Exception as ex:
    print("no bench line", ex)
PY
