"""Print the selection kernel's phase durations (ns) on the GPU box."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svs_b200  # noqa: E402

n, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 1536, int(sys.argv[3]) if len(sys.argv) > 3 else 100
e = svs_b200.Engine()
e.load_synthetic(n, d, 0, 1, 1)
rng = np.random.default_rng(1)
q = rng.random((8, d), dtype=np.float32); q /= np.sqrt((q * q).sum(axis=1))[:, None]
e.bench_set_queries(q)
names = ["stage keys", "threshold", "hit list", "candidates", "sort", "epilogue"]
for it in range(4):
    st = (C.c_uint64 * 16)()
    svs_b200._lib.check(e._lib.svsb_debug_select_phases(e._h, it, k, st))
    t = list(st)
    print({nm: t[i + 1] - t[i] for i, nm in enumerate(names)}, "total_ns", t[6] - t[0], "cands", t[8], "groups", t[9])
e.close()
