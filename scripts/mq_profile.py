"""A few small exact batches at the C2 shape for ncu: gemv_tma_mq_kernel + the batched selection (one CTA per query)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SVSB_MQ_MAX"] = "64"
import svs_b200  # noqa: E402

n, d, k = 1_000_000, 1536, 100
rng = np.random.default_rng(1)
qs = rng.random((8, d), dtype=np.float32)
qs /= np.sqrt((qs * qs).sum(axis=1))[:, None]
eng = svs_b200.Engine([0])
eng.load_synthetic(n, d, seed=0, id0=1, id_step=1)
for b in (4, 8, 4, 8, 4, 8):
    s, i, c = eng.query_batch(qs[:b], k)
ss, ii = eng.query(qs[3], k)
assert np.array_equal(ss.view(np.uint32), s[3].view(np.uint32)) and np.array_equal(ii, i[3])
print("mq_profile ok")
eng.close()
