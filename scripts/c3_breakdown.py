"""Batched pipeline on ONE GPU at a given shard size: ms per 1024-query batch (device-resident), candidates / re-scored rows
per query.  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list of a batch.

    python scripts/c3_breakdown.py [rows] [dims] [k] [batch] [iters]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
b = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 20

rng = np.random.default_rng(2)
q = rng.standard_normal((b, d)).astype(np.float32)
q /= np.sqrt((q * q).sum(axis=1))[:, None]
eng = Engine([0])
eng.load_synthetic(n, d, seed=0, id0=1, id_step=1)
eng.bench_set_queries(q)
for _ in range(3):
    eng.bench_run_batch(k, 10)
r = eng.bench_run_batch(k, iters)
rc = eng.bench_run_batch(k, iters, with_coarse=True)
cand, resc, flags = eng.batch_stats(b)
print(f"rows={n} d={d} k={k} b={b}: {r['total_ms'] / iters * 1e3:.1f} us per batch ({b * iters / r['total_ms'] * 1e3:.0f} queries/s); "
      f"filter pass {rc['coarse_ms'] / iters * 1e3:.1f} us; candidates mean {cand.mean():.0f} max {cand.max()}, "
      f"re-scored mean {resc.mean():.0f} max {resc.max()}, flagged {int((flags != 0).sum())}, launches/batch {r['launches'] / iters:.1f}")
eng.close()
