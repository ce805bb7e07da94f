"""Pairwise top pairs: svsb_top_pairs vs the reference's np.dot(M, M.T) + get_top_pairs on the host.

    python scripts/pairs_bench.py [rows] [dims] [n_pairs] [--no-cpu]
Default = the dad-jokes notebook's shape (examples/dad_jokes/Build Dad Jokes KB.ipynb:338-340: 4,875 x 1536, 10,000 pairs,
0.47 s + 0.03 s on the author's i3-8100).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import svs_b200  # noqa: E402
import svs_oracle as oracle  # noqa: E402  (checker / CPU baseline only)

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if len(args) > 0 else 4875
d = int(args[1]) if len(args) > 1 else 1536
n = int(args[2]) if len(args) > 2 else 10000
rng = np.random.default_rng(0)
m = rng.random((N, d), dtype=np.float32)
m /= np.sqrt((m * m).sum(axis=1))[:, None]
ids = np.arange(1, N + 1, dtype=np.int64)
e = svs_b200.Engine()
e.load(m, ids)
l0 = svs_b200.launch_count()
got = e.top_pairs(n)                                   # first call builds the fp16 shadow
ts = []
for _ in range(5):
    t0 = time.perf_counter(); got = e.top_pairs(n); ts.append(time.perf_counter() - t0)
print(f"svs_b200 top_pairs: {N} x {d}, n={n}: best {min(ts) * 1e3:.2f} ms, median {sorted(ts)[2] * 1e3:.2f} ms "
      f"({0.5 * N * (N - 1) / min(ts) / 1e9:.2f} G pairs/s, {N * (N - 1) * d / min(ts) / 1e12:.1f} useful TFLOP/s), "
      f"{svs_b200.launch_count() - l0} launches in 6 calls")
if "--no-cpu" not in sys.argv:
    t0 = time.perf_counter()
    want = oracle.top_pairwise(m, ids, n)
    dt = time.perf_counter() - t0
    print(f"reference NumPy path on {os.cpu_count()} host cores: {dt * 1e3:.1f} ms  -> {dt / min(ts):.0f}x")
    rep = oracle.compare_pairs(got, want, np.dot(m, m.T), ids)
    print("parity:", rep)
e.close()
