#!/usr/bin/env python
"""bench.py -- retrieve queries/sec on B200 for the SVS hot path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c5|c4] [--impl reference]

A "step" is one batch of QUERIES_PER_STEP single-query retrieves (distinct query vectors) against the
resident matrix: similarity (fp32 GEMV over all rows) + exact top-k, i.e. the body of the reference's
`superheavy()` (src/svs/kb.py:1622-1627).  Workload at N=1: BASELINE.json configs[1] = 1M x 1536 fp32,
top-100 ("c2").  With N>1 (torchrun, one rank per GPU) the SAME matrix is row-sharded over the ranks
(strong scaling): every rank computes its local top-k, the k-candidate lists are all-gathered over NCCL
and merged on every rank by one kernel.

Printed JSON line (rank 0): see the contract in the task description.  Extra keys: `roofline`
(dominant kernel = the similarity kernel, algorithmic bytes n*d*4 per launch over its in-loop CUDA-event
duration, against MEASURED_PEAKS.json's HBM figure), `cpu_baseline` (the reference's NumPy path timed on
this host), `e2e` (the same metric through svsb_query with HOST buffers: H2D of the query and D2H of the
result inside every call), `latency_ms` (one query in flight).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The reference arm must run on all the host threads OpenBLAS can use, but torchrun exports OMP_NUM_THREADS=1 to
# every rank: undo that BEFORE NumPy loads its BLAS (rank 0 alone runs the arm).
if "--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"]:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QUERIES_PER_STEP = 64
WORKLOADS = {
    # name: (rows, dims, k, description)
    "c1": (10_548, 1536, 10, "10,548 x 1536 fp32 top-10 (dad-jokes shape, synthetic unit rows)"),
    "c2": (1_000_000, 1536, 100, "1M x 1536 fp32 top-100 (README 'One Million Documents' shape)"),
    "c3": (1_000_000, 768, 100, "1M x 768 fp32, batches of 1024 queries, top-100 (batched tensor-core path)"),
    "c4": (10_000_000, 1536, 100, "10M x 1536 fp32 top-100, row-sharded"),
    "c5": (1_000_000, 3072, 1000, "1M x 3072 fp32 top-1000"),
}


BATCH = 1024          # queries per step of the batched workload (c3)


def measured_peak_tflops():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            j = json.load(f)
        return float(j["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained, cuBLAS 16-bit dense, kernel timed inside a long step)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md ~1.4 PFLOP/s sustained)"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# reference / CPU arm: the reference's own NumPy path on the host cores
# ---------------------------------------------------------------------------------------------------
def _reference_get_top_k():
    """svs.util.get_top_k from the byte-compiled reference (oracle/_ref) if present, else the oracle port."""
    ref = os.path.join(ROOT, "oracle", "_ref", "svs_ref.bin")    # zip of the byte-compiled reference
    if os.path.isfile(ref):
        try:
            sys.path.insert(0, ref)
            from svs.util import get_top_k          # the reference's own code
            return get_top_k, "reference"
        except Exception:
            sys.path.remove(ref)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from svs_oracle import get_top_k                # CPU restatement (checker), timed only as the baseline
    return get_top_k, "port"


def blas_threads() -> int:
    """Threads the BLAS behind np.dot actually uses (what `cores` must report)."""
    try:
        from threadpoolctl import threadpool_info
        n = [int(p.get("num_threads", 0)) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n and max(n) > 0:
            return max(n)
    except Exception:
        pass
    return os.cpu_count() or 1


def cpu_matrix(n: int, d: int, seed: int = 0) -> np.ndarray:
    """The notebook's recipe (uniform [0,1) rows / L2 norm) in chunks, float32."""
    rng = np.random.default_rng(seed)
    m = np.empty((n, d), dtype=np.float32)
    step = 65536
    for a in range(0, n, step):
        b = min(n, a + step)
        blk = rng.random((b - a, d), dtype=np.float32)
        blk /= np.sqrt(np.einsum("ij,ij->i", blk, blk))[:, None]
        m[a:b] = blk
    return m


def cpu_arm(n: int, d: int, k: int, budget_s: float = 20.0):
    """Returns (matrix rows used, function running one query, description)."""
    get_top_k, kind = _reference_get_top_k()
    try:
        avail = int(next(l for l in open("/proc/meminfo") if l.startswith("MemAvailable")).split()[1]) * 1024
    except Exception:
        avail = 8 << 30
    rows = n
    while rows * d * 4 > avail * 0.4 and rows > 50_000:
        rows //= 2
    m = cpu_matrix(rows, d, 0)
    ids = np.arange(1, rows + 1, dtype=np.int64)

    def one_query(qv):
        x = np.dot(m, qv)                                      # src/svs/kb.py:1623
        return [(s, int(ids[i])) for s, i in get_top_k(x, k)]  # src/svs/kb.py:1625-1626
    return rows, one_query, kind


def run_reference_arm(args, n, d, k, desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                                  # other ranks exit 0 without work
    rows, one_query, kind = cpu_arm(n, d, k)
    rng = np.random.default_rng(1)
    qs = rng.random((8, d), dtype=np.float32)
    qs /= np.sqrt((qs * qs).sum(axis=1))[:, None]
    per_step = 2                                                # bounded sample: 2 queries per step
    for w in range(args.warmup):
        one_query(qs[w % len(qs)])
    t0 = time.perf_counter()
    for s in range(args.steps):
        for j in range(per_step):
            one_query(qs[(s * per_step + j) % len(qs)])
    dt = time.perf_counter() - t0
    scale = rows / n                                            # < 1 only if the host lacks RAM for n rows
    qps = args.steps * per_step / dt * scale
    sample = (f"{args.steps} steps x {per_step} queries of np.dot + get_top_k on {rows} x {d} host rows"
              + ("" if rows == n else f", scaled x{scale:.3f} to {n} rows (host RAM bound)"))
    line = {
        "impl": "reference", "metric": "retrieve_queries_per_sec", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "rows": n, "dims": d, "k": k},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": blas_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm, batched workload (c3): one step = one batch of BATCH queries through svsb_query_batch's pipeline
# ---------------------------------------------------------------------------------------------------
def run_batch_arm(args, n, d, k, desc):
    import svs_b200
    from svs_b200.engine import Engine
    rng = np.random.default_rng(2)
    queries = rng.random((BATCH, d), dtype=np.float32)
    queries /= np.sqrt((queries * queries).sum(axis=1))[:, None]
    peak, peak_src = measured_peak_tflops()
    eng = Engine([0])
    t_load = time.perf_counter()
    eng.load_synthetic(n, d, seed=0, id0=1, id_step=1)
    load_s = time.perf_counter() - t_load
    eng.bench_set_queries(queries)
    for _ in range(args.warmup):                                # a batch is ~2 ms: 10 per warm-up step lets the clocks ramp
        eng.bench_run_batch(k, 10)
    sampler = ClockSampler(0)
    sampler.start()
    total_ms = coarse_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        r = eng.bench_run_batch(k, 1, with_coarse=True)
        total_ms += r["total_ms"]; coarse_ms += r["coarse_ms"]; launches += r["launches"]
    clocks = sampler.stop()
    cand, resc, flags = eng.batch_stats(BATCH)
    # the timed path must compute the answer: compare a few queries with the single-query kernels
    agree = True
    for qi in (0, BATCH // 2, BATCH - 1):
        got, flag = eng.bench_batch_result(qi, k)
        agree = agree and flag == 0 and got == eng.retrieve(queries[qi], k)
    # e2e: the public C-ABI call with host buffers (H2D of the batch from pinned host memory, D2H of the results
    # into pinned host memory, inside every call)
    from svs_b200 import pinned_empty
    hq = pinned_empty((BATCH, d), np.float32)
    hq[:] = queries
    out = (pinned_empty((BATCH, k), np.float32), pinned_empty((BATCH, k), np.int64), np.zeros(BATCH, dtype=np.int32))
    eng.query_batch(hq, k, out=out)
    t0 = time.perf_counter()
    for s in range(args.steps):
        eng.query_batch(hq, k, out=out)
    e2e_qps = args.steps * BATCH / (time.perf_counter() - t0)
    ref_s, ref_i, _ = eng.query_batch(queries, k)               # pageable buffers take the staged path: same answer
    agree = agree and np.array_equal(out[0].view(np.uint32), ref_s.view(np.uint32)) and np.array_equal(out[1], ref_i)
    nq = args.steps * BATCH
    flop = 2.0 * n * d * BATCH
    achieved = flop * args.steps / (coarse_ms / 1e3) / 1e12
    line = {
        "metric": "retrieve_queries_per_sec", "value": nq / (total_ms / 1e3), "unit": "queries/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 results (f16 tensor-core coarse pass + exact f32 re-score)",
        "data": "synthetic",
        "config": {"workload": desc, "rows": n, "dims": d, "k": k, "queries_per_step": BATCH,
                   "l2": "inputs larger than L2 (f16 shadow matrix %.2f GB)" % (n * d * 2 / 1e9), "parallelism": "1 GPU"},
        "ms_per_query": total_ms / nq,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "coarse_gemm_kernel<0> (filter pass)", "peak_source": peak_src,
                     "algorithmic_flop_per_launch": flop, "coarse_ms_per_batch": coarse_ms / args.steps,
                     "hbm_floor_ms": n * d * 2 / 1e9 / measured_peak_gbs()[0] * 1e3},
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": BATCH * d * 4, "d2h_bytes_per_step": BATCH * (k * 12 + 8)},
        "gpu_launches": int(launches), "clocks": clocks, "load_synthetic_s": load_s,
        "batch_stats": {"candidates_mean": float(cand.mean()), "candidates_max": int(cand.max()), "rescored_mean": float(resc.mean()),
                        "rescored_max": int(resc.max()), "fallback_queries": int((flags != 0).sum()),
                        "agrees_with_single_query_bits": bool(agree)},
    }
    traffic_file = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(traffic_file):
        try:
            line["roofline"]["traffic"] = json.load(open(traffic_file)).get(args.workload)
        except Exception:
            pass
    if not args.no_cpu_baseline:
        rows, one_query, kind = cpu_arm(n, d, k)
        for i in range(2):
            one_query(queries[i])
        t0 = time.perf_counter(); cnt = 0
        while cnt < 10 or (time.perf_counter() - t0 < 5.0 and cnt < 64):
            one_query(queries[cnt % len(queries)]); cnt += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": cnt / dt * (rows / n), "unit": "queries/s", "cores": blas_threads(), "kind": kind,
            "sample": f"{cnt} of the {BATCH} queries, each np.dot + get_top_k on {rows} x {d} host rows (the reference has no batched API)"}
    eng.close()
    print(json.dumps(line), flush=True)


def run_batch_arm_sharded(args, sr, dist, torch, rank, world, local_rank, n, d, k, desc, l0):
    """c3 on N GPUs: rows sharded, every rank runs the batched pipeline on its shard, ONE all-gather of 1024 records
    per rank and ONE merge launch per batch."""
    import svs_b200
    rng = np.random.default_rng(2)
    queries = rng.random((BATCH, d), dtype=np.float32)
    queries /= np.sqrt((queries * queries).sum(axis=1))[:, None]
    sr.set_queries(queries)
    for _ in range(args.warmup * 10):
        sr.run_batch(k)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dist.barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        sr.run_batch(k)
    ev1.record()
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms[0])
    clocks = sampler.stop() if rank == 0 else None
    fallbacks = sr.last_fallbacks
    for _ in range(2):
        sr.retrieve_many_arrays(queries, k)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sr.retrieve_many_arrays(queries, k)                       # host (b, d) array in, host (b, k) arrays out
    torch.cuda.synchronize(); dist.barrier()
    e2e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    launches = svs_b200.launch_count() - l0
    if rank == 0:
        nq = args.steps * BATCH
        peak, peak_src = measured_peak_tflops()
        flop_per_gpu = 2.0 * sr.local_rows * d * BATCH
        achieved = flop_per_gpu * args.steps / (total_ms / 1e3) / 1e12
        line = {
            "metric": "retrieve_queries_per_sec", "value": nq / (total_ms / 1e3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 results (f16 tensor-core coarse pass + exact f32 re-score)", "data": "synthetic",
            "config": {"workload": desc, "rows": n, "dims": d, "k": k, "queries_per_step": BATCH,
                       "l2": "per-GPU f16 shadow shard %.2f GB" % (sr.local_rows * d * 2 / 1e9),
                       "parallelism": f"row-sharded over {world} GPUs, one NCCL all-gather of {BATCH} k-candidate records per rank per batch + merge kernel"},
            "ms_per_query": total_ms / nq,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "whole batch step per GPU (coarse passes + refine + exchange + merge), rank-max time",
                         "peak_source": peak_src, "algorithmic_flop_per_launch": flop_per_gpu},
            "e2e": {"value": nq / float(e2e[0]), "unit": "queries/s", "h2d_bytes_per_step": BATCH * d * 4,
                    "d2h_bytes_per_step": BATCH * (k * 12 + 4)},
            "gpu_launches": int(launches), "clocks": clocks, "batch_stats": {"fallback_queries_rank0": int(fallbacks)},
        }
        print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "collective"],
                    help="N>1, single queries: fused push over NVLink peer memory (default) or NCCL all-gather + merge")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    n, d, k, desc = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, n, d, k, desc)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    import svs_b200                                             # fails loudly if the CUDA library is missing
    rng = np.random.default_rng(1)
    queries = rng.random((128, d), dtype=np.float32)            # distinct queries: 128 x 6 KB, never L2-useful
    queries /= np.sqrt((queries * queries).sum(axis=1))[:, None]
    peak, peak_src = measured_peak_gbs()
    l0 = svs_b200.launch_count()

    if world == 1 and args.workload == "c3":
        run_batch_arm(args, n, d, k, desc)
        return

    if world == 1:
        from svs_b200.engine import Engine
        eng = Engine([0])
        t_load = time.perf_counter()
        eng.load_synthetic(n, d, seed=0, id0=1, id_step=1)
        load_s = time.perf_counter() - t_load
        eng.bench_set_queries(queries)
        for _ in range(args.warmup):
            eng.bench_run(k, QUERIES_PER_STEP)
        sampler = ClockSampler(0)
        sampler.start()
        total_ms = gemv_ms = 0.0
        launches = 0
        for _ in range(args.steps):                             # each step: events on the launching stream, sync on both sides
            r = eng.bench_run(k, QUERIES_PER_STEP, with_gemv=True)
            total_ms += r["total_ms"]; gemv_ms += r["gemv_ms"]; launches += r["launches"]
        clocks = sampler.stop()
        nq = args.steps * QUERIES_PER_STEP
        value = nq / (total_ms / 1e3)
        # latency: one query in flight, host-visible (host buffers in and out)
        lat = []
        for i in range(40):
            t0 = time.perf_counter(); eng.query(queries[i % len(queries)], k); lat.append(time.perf_counter() - t0)
        # e2e: the public C-ABI call with host buffers, H2D + D2H inside every call
        for i in range(8):
            eng.query(queries[i], k)
        t0 = time.perf_counter()
        for s in range(args.steps):
            for j in range(QUERIES_PER_STEP):
                eng.query(queries[(s * QUERIES_PER_STEP + j) % len(queries)], k)
        e2e_qps = nq / (time.perf_counter() - t0)
        algo_bytes = n * d * 4
        achieved = algo_bytes * nq / (gemv_ms / 1e3) / 1e9
        line = {
            "metric": "retrieve_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "rows": n, "dims": d, "k": k, "queries_per_step": QUERIES_PER_STEP,
                       "l2": "inputs larger than L2 (matrix %.2f GB, distinct queries)" % (algo_bytes / 1e9),
                       "parallelism": "1 GPU"},
            "ms_per_query": total_ms / nq, "latency_ms": {"median": float(np.median(lat)) * 1e3, "min": float(min(lat)) * 1e3},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "gemv_tma_kernel", "peak_source": peak_src,
                         "timed_launches": "1 in 8 bracketed with CUDA events inside the timed loop",
                         "algorithmic_bytes_per_launch": algo_bytes, "frac_of_nominal_8TBs": achieved / 8000.0,
                         "whole_query_frac": (algo_bytes * nq / (total_ms / 1e3) / 1e9) / peak},
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": QUERIES_PER_STEP * d * 4,
                    "d2h_bytes_per_step": QUERIES_PER_STEP * (k * 12 + 4)},
            "gpu_launches": int(launches), "clocks": clocks, "load_synthetic_s": load_s,
        }
        traffic_file = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(traffic_file):
            try:
                line["roofline"]["traffic"] = json.load(open(traffic_file)).get(args.workload)
            except Exception:
                pass
        if not args.no_cpu_baseline:
            rows, one_query, kind = cpu_arm(n, d, k)
            for i in range(2):
                one_query(queries[i])
            t0 = time.perf_counter(); cnt = 0
            while cnt < 10 or (time.perf_counter() - t0 < 5.0 and cnt < 40):
                one_query(queries[cnt % len(queries)]); cnt += 1
            dt = time.perf_counter() - t0
            scale = rows / n
            line["cpu_baseline"] = {
                "value": cnt / dt * scale, "unit": "queries/s", "cores": blas_threads(), "kind": kind,
                "sample": f"{cnt} queries of np.dot + get_top_k on {rows} x {d} host rows"
                          + ("" if rows == n else f", scaled x{scale:.3f}")}
        eng.close()
        print(json.dumps(line), flush=True)
        return

    # ---- N > 1: one rank per GPU, row shards, NCCL all-gather of the candidate lists ----------
    import torch
    import torch.distributed as dist
    from svs_b200.sharded import ShardedRetriever
    torch.cuda.set_device(local_rank)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"                       # keep NCCL's version banner off stdout: one JSON line only
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sr = ShardedRetriever(rank, world, local_rank, exchange=args.exchange)
    sr.load_synthetic(n, d, seed=0, id0=1, id_step=1)
    if args.workload == "c3":
        run_batch_arm_sharded(args, sr, dist, torch, rank, world, local_rank, n, d, k, desc, l0)
        sr.close()
        dist.destroy_process_group()
        return
    sr.set_queries(queries)
    for _ in range(args.warmup):
        sr.run_queries(k, QUERIES_PER_STEP)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dist.barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    gemv_ms = 0.0
    for _ in range(args.steps):
        gemv_ms += sr.run_queries(k, QUERIES_PER_STEP, time_gemv=True)
    ev1.record()
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1), gemv_ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms, gemv_ms = float(ms[0]), float(ms[1])
    clocks = sampler.stop() if rank == 0 else None
    nq = args.steps * QUERIES_PER_STEP
    # e2e: host query in, host result out, every call
    for i in range(4):
        sr.retrieve_arrays(queries[i], k)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(args.steps):
        for j in range(QUERIES_PER_STEP):                       # host query in, host (scores, ids) arrays out, as Engine.query at N=1
            sr.retrieve_arrays(queries[(s * QUERIES_PER_STEP + j) % len(queries)], k)
    torch.cuda.synchronize(); dist.barrier()
    e2e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    launches = torch.tensor([svs_b200.launch_count() - l0], device="cuda", dtype=torch.int64)
    if rank == 0:
        shard_bytes = sr.local_rows * d * 4
        achieved = shard_bytes * nq / (gemv_ms / 1e3) / 1e9
        line = {
            "metric": "retrieve_queries_per_sec", "value": nq / (total_ms / 1e3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "rows": n, "dims": d, "k": k, "queries_per_step": QUERIES_PER_STEP,
                       "l2": "per-GPU shard %.2f GB, distinct queries" % (shard_bytes / 1e9),
                       "parallelism": (f"row-sharded over {world} GPUs, k-candidate records pushed into every rank's window over "
                                       "NVLink peer memory by the selection kernel, merge kernel waits on flags (no collective call)"
                                       if sr.exchange == "peer" else
                                       f"row-sharded over {world} GPUs, NCCL all-gather of k candidates + merge kernel")},
            "ms_per_query": total_ms / nq,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "gemv_tma_kernel (per GPU, rank-max time)", "peak_source": peak_src,
                         "timed_launches": "1 in 8 bracketed with CUDA events inside the timed loop",
                         "algorithmic_bytes_per_launch": shard_bytes},
            "e2e": {"value": nq / float(e2e[0]), "unit": "queries/s", "h2d_bytes_per_step": QUERIES_PER_STEP * d * 4,
                    "d2h_bytes_per_step": QUERIES_PER_STEP * (k * 12 + 4)},
            "gpu_launches": int(launches[0]), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    sr.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
