"""Sweep the GEMV variants/knobs on the GPU box and print achieved GB/s (algorithmic bytes / time).

    python scripts/gpu_tune.py [--n 1000000] [--d 1536] [--k 100] [--iters 50] [--quick]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svs_b200  # noqa: E402


def run(e, k, iters, env):
    for key in ("SVSB_GEMV_VARIANT", "SVSB_GEMV_TUNE_A", "SVSB_GEMV_TUNE_B"):
        os.environ.pop(key, None)
    os.environ.update({k_: str(v) for k_, v in env.items()})
    e.bench_run(k, 5, with_gemv=True)                       # warm-up
    return e.bench_run(k, iters, with_gemv=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=1536)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    e = svs_b200.Engine()
    e.load_synthetic(a.n, a.d, seed=0, id0=1, id_step=1)
    rng = np.random.default_rng(1)
    q = rng.random((64, a.d), dtype=np.float32)
    q /= np.sqrt((q * q).sum(axis=1))[:, None]
    e.bench_set_queries(q)
    gb = a.n * a.d * 4 / 1e9
    configs = [{"SVSB_GEMV_VARIANT": 2}]
    if not a.quick:
        row_bytes = ((a.d + 3) // 4) * 16
        for tr in (2, 4, 8, 16, 32):
            for st in (2, 3, 4, 6, 8):
                if tr * st * row_bytes > 216 * 1024:
                    continue
                for cw in (8, 16):
                    configs.append({"SVSB_GEMV_VARIANT": 2, "SVSB_GEMV_TUNE_A": tr, "SVSB_GEMV_TUNE_B": cw * 100 + st})
        for ta in (1, 5, 6, 8):
            configs.append({"SVSB_GEMV_VARIANT": 1, "SVSB_GEMV_TUNE_A": ta})
        configs = configs * 2
        import random
        random.Random(0).shuffle(configs)
    else:
        configs += [{"SVSB_GEMV_VARIANT": 1, "SVSB_GEMV_TUNE_A": 5}]
    for env in configs:
        try:
            r = run(e, a.k, a.iters, env)
        except Exception as ex:
            print(json.dumps({"env": env, "error": str(ex)}))
            continue
        per = r["total_ms"] / a.iters
        gper = r["gemv_ms"] / a.iters
        print(json.dumps({"n": a.n, "d": a.d, "k": a.k, "env": env, "ms_per_query": round(per, 4),
                          "gemv_ms": round(gper, 4), "gemv_GBs": round(gb / gper * 1e3, 1),
                          "query_GBs": round(gb / per * 1e3, 1), "select_us": round((per - gper) * 1e3, 1)}), flush=True)
    e.close()


if __name__ == "__main__":
    main()
