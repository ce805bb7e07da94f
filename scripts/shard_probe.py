"""Fixed per-query costs of the single-query path as a function of shard size (one GPU): per-query time and
similarity-kernel span (CUDA events inside svsb_bench_run) for shards of 1/8 .. 1x of the 1M x 1536 matrix,
pipelined (selection overlapped on a free SM) and serial."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_b200.engine import Engine

d, k = 1536, 100
rng = np.random.default_rng(1)
q = rng.random((128, d), dtype=np.float32)
q /= np.sqrt((q * q).sum(axis=1))[:, None]
for rows in (125_000, 250_000, 500_000, 1_000_000):
    eng = Engine([0])
    eng.load_synthetic(rows, d, seed=0, id0=1, id_step=1)
    eng.bench_set_queries(q)
    for pipe in ("1", "0"):
        os.environ["SVSB_PIPELINE"] = pipe
        for timed in (True, False):
            for _ in range(3):
                eng.bench_run(k, 64)
            tot = gem = 0.0
            for _ in range(10):
                r = eng.bench_run(k, 64, with_gemv=timed)
                tot += r["total_ms"]; gem += r["gemv_ms"] or 0.0
            ideal = rows * d * 4 / 7.2e12 * 1e6
            print(f"rows={rows} pipelined={pipe} kernel_events={timed}: {tot / 640 * 1e3:.1f} us/query, kernel span {gem / 640 * 1e3:.1f} us, "
                  f"at 7.2 TB/s {ideal:.1f} us", flush=True)
    eng.close()
