"""Synchronous single-query latency (svsb_query, host buffers) with programmatic dependent launch on / off."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_b200.engine import Engine

for rows, d, k in ((10_548, 1536, 10), (125_000, 1536, 100), (1_000_000, 1536, 100)):
    rng = np.random.default_rng(1)
    q = rng.random((64, d), dtype=np.float32)
    q /= np.sqrt((q * q).sum(axis=1))[:, None]
    eng = Engine([0])
    eng.load_synthetic(rows, d, seed=0, id0=1, id_step=1)
    ref = None
    for pdl in ("1", "0", "1", "0"):
        os.environ["SVSB_PDL"] = pdl
        for i in range(20):
            eng.query(q[i], k)
        lat = []
        for i in range(400):
            t0 = time.perf_counter(); s, ids = eng.query(q[i % 64], k); lat.append(time.perf_counter() - t0)
        if ref is None:
            ref = (s.copy(), ids.copy())
        assert np.array_equal(ref[0].view(np.uint32), s.view(np.uint32)) and np.array_equal(ref[1], ids)
        lat = np.array(lat) * 1e6
        print(f"rows={rows} k={k} pdl={pdl}: median {np.median(lat):.1f} us, p10 {np.percentile(lat, 10):.1f}, min {lat.min():.1f}", flush=True)
    eng.close()
