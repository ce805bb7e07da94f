"""Small exact batches at BASELINE shapes: wall-clock of svsb_query_batch (host buffers in and out) per batch size, the
multi-query passes against the alternatives (looped single queries; the tensor-core coarse path where it applies).

    python scripts/mq_sweep.py [c2|c3|c5]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svs_b200  # noqa: E402

SHAPES = {"c2": (1_000_000, 1536, 100), "c3": (1_000_000, 768, 100), "c5": (1_000_000, 3072, 1000)}


def timed(eng, q, k, reps=12):
    eng.query_batch(q, k)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); eng.query_batch(q, k); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n, d, k = SHAPES[name]
    rng = np.random.default_rng(1)
    qs = rng.random((64, d), dtype=np.float32)
    qs /= np.sqrt((qs * qs).sum(axis=1))[:, None]
    eng = svs_b200.Engine([0])
    eng.load_synthetic(n, d, seed=0, id0=1, id_step=1)
    t0 = time.perf_counter(); eng.query(qs[0], k); eng.query(qs[1], k)
    one = []
    for i in range(20):
        t0 = time.perf_counter(); eng.query(qs[i], k); one.append(time.perf_counter() - t0)
    print(f"{name}: {n} x {d}, k={k}; single query {np.median(one) * 1e3:.3f} ms")
    print("b   | mq groups<=1 | groups<=2 | groups<=4 | groups<=8 | coarse path | loop of single queries")
    for b in (2, 4, 8, 16, 32, 64):
        row = [f"{b:<3d}"]
        for groups in (1, 2, 4, 8):
            os.environ["SVSB_MQ_GROUPS"] = str(groups); os.environ["SVSB_MQ_MAX"] = "1000000"
            row.append(f"{timed(eng, qs[:b], k):9.3f}")
        os.environ["SVSB_MQ_MAX"] = "1"
        row.append(f"{timed(eng, qs[:b], k):9.3f}" if b >= 4 else "      n/a")
        os.environ["SVSB_MQ"] = "0"; os.environ["SVSB_BATCH_MIN"] = "1000000"
        row.append(f"{timed(eng, qs[:b], k, reps=3):9.3f}")
        del os.environ["SVSB_MQ"]; del os.environ["SVSB_BATCH_MIN"]
        print(" | ".join(row), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
