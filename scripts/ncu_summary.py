"""Turn gpurun_out/*.ncu-rep and launches.csv into the small text summaries committed under profiles/.

    python scripts/ncu_summary.py <round-tag>        # e.g. r01

Reads with `ncu -i ... --page raw --csv` (works without a GPU).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
]


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit)
    return float(val.replace(",", "")) * mult if mult else None


def summarise(rep, title, tag, only=None):
    hdr, units, rows = raw_rows(rep)
    lines = [f"# {title}", "", f"Source: `gpurun_out/{os.path.basename(rep)}` (`ncu --set full --clock-control none --import-source on`), "
             "read with `ncu -i ... --page raw --csv`.  Times under ncu are serialised/cold-cache: use the shares, not the absolutes.", ""]
    traffic = []
    for r in rows:
        name = r[hdr.index("Kernel Name")]
        lines += [f"## {name[:110]}", "", "| metric | value | unit |", "|---|---|---|"]
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append(f"| `{m}` | {r[i]} | {units[i]} |")
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        if rd is not None and wr is not None:
            if only is None or only in name:
                traffic.append(rd + wr)
            lines.append(f"| **traffic = dram read + write** | {rd + wr:.0f} | byte |")
        lines.append("")
    path = os.path.join(PROF, f"{tag}.md")
    open(path, "w").write("\n".join(lines))
    print("wrote", path)
    return traffic


def launches(tag, src=None):
    src = src or os.path.join(OUT, "launches.csv")
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[ki].split("(")[0], []).append(float(r[vi]))
    tot = sum(sum(v) for v in d.values())
    lines = [f"# {tag}: launch list of the bench step (ncu --metrics gpu__time_duration.sum --clock-control none)", "",
             "Per-launch times are cold-cache and serialised: the SHARE of the step is what must agree with bench.py.", "",
             "| kernel | launches | mean ns | min ns | max ns | share of summed time |", "|---|---|---|---|---|---|"]
    for k, v in d.items():
        lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.0f} | {min(v):.0f} | {max(v):.0f} | {sum(v) / tot:.4f} |")
    path = os.path.join(PROF, f"{tag}.md")
    open(path, "w").write("\n".join(lines) + "\n")
    # keep the raw list too (small)
    with open(os.path.join(PROF, f"{tag}.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "block", "grid", "ns"])
        for r in rows[1:]:
            w.writerow([r[0], r[ki].split("(")[0], r[hdr.index("Block Size")], r[hdr.index("Grid Size")], r[vi]])
    print("wrote", path)


if __name__ == "__main__":
    rtag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    out = {}
    for rep, title, key, only in (
            ("prof_gemv.ncu-rep", "similarity kernel (gemv_tma_kernel), workload c2 = 1M x 1536 fp32", "c2", None),
            ("prof_select.ncu-rep", "selection kernel (select_topk_kernel), workload c2, k = 100", None, None),
            ("prof_coarse.ncu-rep", "batched coarse contraction (coarse_gemm_kernel<1> sample pass, <0> filter pass), workload c3",
             "c3", "coarse_gemm_kernel<0"),
            ("prof_refine.ncu-rep", "batched path: sample_order_kernel (order statistic of the sample) and refine_lean_kernel (select + exact "
             "re-score + sort, one CTA per query), workload c3", None, None),
            ("prof_mq.ncu-rep", "small exact batches (scripts/mq_profile.py): gemv_tma_mq_kernel (4 and 8 queries per pass) and the batched "
             "selection (one CTA per query), 1M x 1536, k = 100", "c2_mq", "gemv_tma_mq"),
            ("prof_peer.ncu-rep", "peer exchange at world size 1 (scripts/peer_profile.py): select_topk_kernel with the fused push, "
             "merge_window_kernel; 1M x 1536 shard, k = 100", None, None)):
        p = os.path.join(OUT, rep)
        if os.path.exists(p):
            t = summarise(p, title, f"{rtag}_{rep.split('.')[0][5:]}_ncu", only)
            if key and t:
                out[key] = sum(t) / len(t)
    if os.path.exists(os.path.join(OUT, "launches.csv")):
        launches(f"{rtag}_launches")
    if os.path.exists(os.path.join(OUT, "launches_c3.csv")):
        launches(f"{rtag}_launches_c3", os.path.join(OUT, "launches_c3.csv"))
    if os.path.exists(os.path.join(OUT, "launches_c3_virtual_n8.csv")):   # scripts/c3_virtual_ranks.py: one rank's batch of an 8-rank run
        launches(f"{rtag}_launches_c3_virtual_n8", os.path.join(OUT, "launches_c3_virtual_n8.csv"))
    tf = os.path.join(PROF, f"{rtag}_traffic.json")
    old = json.load(open(tf)) if os.path.exists(tf) else {}
    old.update(out)
    json.dump(old, open(tf, "w"), indent=1)
    print("wrote", tf, old)
