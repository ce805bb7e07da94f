"""GPU smoke of the batched path: batch results must equal the single-query results bit for bit.

    python scripts/batch_check.py [n] [d] [k] [b] [dist]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svs_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
b = int(sys.argv[4]) if len(sys.argv) > 4 else 300
dist = sys.argv[5] if len(sys.argv) > 5 else "uniform"

rng = np.random.default_rng(7)
gen = (lambda *s: rng.random(s, dtype=np.float32)) if dist == "uniform" else (lambda *s: rng.standard_normal(s).astype(np.float32))
m = gen(n, d); m /= np.sqrt((m * m).sum(axis=1))[:, None]
q = gen(b, d); q /= np.sqrt((q * q).sum(axis=1))[:, None]
ids = np.arange(1, n + 1, dtype=np.int64)
e = svs_b200.Engine()
e.load(m, ids)
t0 = time.perf_counter()
s, i, c = e.query_batch(q, k)
t1 = time.perf_counter()
cand, resc, flags = e.batch_stats(min(b, 2048))
print(f"batch: {t1 - t0:.3f}s  counts ok={bool((c == min(k, n)).all())}  cand mean={cand.mean():.0f} max={cand.max()}  "
      f"rescored mean={resc.mean():.0f} max={resc.max()}  flagged={int((flags != 0).sum())}")
bad = 0
for j in range(min(b, 64)):
    ss, ii = e.query(q[j], k)
    if not (np.array_equal(ss.view(np.uint32), s[j, :len(ss)].view(np.uint32)) and np.array_equal(ii, i[j, :len(ii)])):
        bad += 1
        if bad <= 3:
            diff = np.nonzero((ss.view(np.uint32) != s[j, :len(ss)].view(np.uint32)) | (ii != i[j, :len(ii)]))[0]
            print("MISMATCH query", j, "first ranks", diff[:5], ss[diff[:3]], s[j, diff[:3]], ii[diff[:3]], i[j, diff[:3]])
print("bit-exact vs single-query path:", "OK" if bad == 0 else f"{bad} queries differ")
e.close()
sys.exit(1 if bad else 0)
