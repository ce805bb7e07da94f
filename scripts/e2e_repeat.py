"""Repeatability of the host-buffer throughput paths at C2: svsb_query (one in flight) and svsb_query_submit / _wait (3 in
flight), several passes each, optionally with programmatic dependent launch off (SVSB_PDL=0).

    python scripts/e2e_repeat.py [rows] [dims] [k] [passes] [queries per pass]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 8
nq = int(sys.argv[5]) if len(sys.argv) > 5 else 640
rng = np.random.default_rng(1)
q = rng.standard_normal((128, d)).astype(np.float32)
q /= np.sqrt((q * q).sum(axis=1))[:, None]
eng = Engine([0])
eng.load_synthetic(n, d, seed=0, id0=1, id_step=1)
for i in range(16):
    eng.query(q[i], k)
sync, piped = [], []
for p in range(passes):
    t0 = time.perf_counter()
    for j in range(nq):
        eng.query(q[j % 128], k)
    sync.append(nq / (time.perf_counter() - t0))
    pend = []
    gaps = []
    t0 = time.perf_counter()
    for j in range(nq):
        pend.append(eng.submit(q[j % 128], k))
        if len(pend) == 3:
            t1 = time.perf_counter(); pend.pop(0).result(); gaps.append(time.perf_counter() - t1)
    for h in pend:
        h.result()
    piped.append(nq / (time.perf_counter() - t0))
    g = np.array(gaps) * 1e3
    print(f"pass {p}: sync {sync[-1]:.1f} q/s, 3 in flight {piped[-1]:.1f} q/s; wait per result ms: median {np.median(g):.3f} p90 {np.percentile(g, 90):.3f} max {g.max():.3f}", flush=True)
print(f"PDL={os.environ.get('SVSB_PDL', '1')}: sync median {np.median(sync):.1f} (min {min(sync):.1f}), 3 in flight median {np.median(piped):.1f} (min {min(piped):.1f}, max {max(piped):.1f})")
eng.close()
