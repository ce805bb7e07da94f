"""Stress the synchronous single-query path (svsb_query) with programmatic dependent launch on / off: every result is
checked against NumPy on the host.  Shapes alternate so that buffers and template instantiations change under it."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_b200.engine import Engine

rng = np.random.default_rng(7)
shapes = [(4096, 768, 100), (10_548, 1536, 10), (3001, 1537, 17), (20_000, 3072, 1000)]
mats = []
for n, d, k in shapes:
    m = rng.standard_normal((n, d)).astype(np.float32)
    m /= np.sqrt((m * m).sum(axis=1))[:, None]
    mats.append(m)
eng = Engine([0])
for pdl in ("1", "0"):
    for variant in ("2", "1"):
        os.environ["SVSB_PDL"] = pdl
        os.environ["SVSB_GEMV_VARIANT"] = variant
        bad = total = 0
        prev_q = None
        t0 = time.time()
        for rep in range(int(os.environ.get('REPS', '6'))):
            for (n, d, k), m in zip(shapes, mats):
                eng.load(m, np.arange(1, n + 1, dtype=np.int64))
                for it in range(60):
                    q = rng.standard_normal(d).astype(np.float32)
                    q /= np.sqrt((q * q).sum())
                    s, ids = eng.query(q, k)
                    x = m @ q
                    want = np.sort(x)[::-1][:k]
                    total += 1
                    if len(s) != k or not np.allclose(s, want, rtol=2e-5, atol=1e-6) or not np.allclose(x[ids - 1], s, rtol=2e-5, atol=1e-6):
                        bad += 1
                        if bad <= 6 and len(s) == k:
                            own = np.isclose(x[ids - 1], s, rtol=2e-5, atol=1e-6)           # score belongs to the returned id under THIS q
                            prev_x = m @ prev_q if prev_q is not None and prev_q.shape == q.shape else None
                            stale = None if prev_x is None else int(np.isclose(prev_x[ids - 1], s, rtol=2e-5, atol=1e-6).sum())
                            missing = int((~np.isin(np.argsort(-x)[:k] + 1, ids)).sum())
                            print("MISMATCH", pdl, variant, (n, d, k), "iter", it, "| entries with a correct own score:", int(own.sum()), "of", k,
                                  "| entries matching the PREVIOUS query's scores:", stale, "| true top-k rows missing:", missing,
                                  "| wrong entry rows:", (ids[~own] - 1)[:8].tolist(), flush=True)
                    prev_q = q
        print(f"pdl={pdl} variant={variant}: {bad} bad of {total} in {time.time() - t0:.1f}s", flush=True)
eng.close()
