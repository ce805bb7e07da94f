"""Which rows go wrong under PDL?  k = 2048 of 3001 rows returns most of the score vector."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_b200.engine import Engine
rng = np.random.default_rng(7)
os.environ["SVSB_PDL"] = "1"; os.environ["SVSB_GEMV_VARIANT"] = "2"
eng = Engine([0])
for (n, d, k) in [(3001, 1537, 2048), (10_548, 1536, 2048)]:
    m = rng.standard_normal((n, d)).astype(np.float32); m /= np.sqrt((m * m).sum(axis=1))[:, None]
    eng.load(m, np.arange(1, n + 1, dtype=np.int64))
    bad = 0; prev_x = None
    for it in range(20000):
        q = rng.standard_normal(d).astype(np.float32); q /= np.sqrt((q * q).sum())
        s, ids = eng.query(q, k)
        x = m @ q
        own = np.isclose(x[ids - 1], s, rtol=2e-5, atol=1e-6)
        if not own.all():
            bad += 1
            rows = np.sort(ids[~own] - 1)
            st = "" if prev_x is None else f" stale(prev query's score): {int(np.isclose(prev_x[ids[~own]-1], s[~own], rtol=2e-5, atol=1e-6).sum())}"
            print((n, d, k), "iter", it, "wrong-score rows:", rows.tolist()[:40], "count", len(rows), st, flush=True)
            if bad >= 5: break
        prev_x = x
    print((n, d, k), "bad", bad, "of", it + 1, flush=True)
eng.close()
