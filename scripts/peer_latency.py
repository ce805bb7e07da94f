"""NVLink-side timing of the fused exchange under torchrun (SVSB_XCHG_STAMPS=1): for the synchronous svsb_query_peer,
%globaltimer stamps written by the selection kernel (with its push) and the waiting merge kernel on every rank.

    SVSB_XCHG_STAMPS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29631 scripts/peer_latency.py [rows] [dims] [k]

Per query and rank (all on that rank's own clock): selection kernel duration (threshold, candidates, sort, push of the
record into all `world` windows + flags), gap until the merge kernel starts, how long the merge then waited for the LAST
peer's flag (rank skew + NVLink delivery), merge duration; plus the host-visible latency of the call.
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svs_b200.sharded import ShardedRetriever


def pct(a, p):
    return float(np.percentile(a, p)) if len(a) else float("nan")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sr = ShardedRetriever(rank, world, local, exchange="peer")
    sr.load_synthetic(n, d, seed=0, id0=1, id_step=1)
    rng = np.random.default_rng(1)
    qs = rng.random((128, d), dtype=np.float32)
    qs /= np.sqrt((qs * qs).sum(axis=1))[:, None]
    for i in range(20):
        sr.retrieve_arrays(qs[i], k)
    dist.barrier(); torch.cuda.synchronize()
    lat = []
    nq = 1000
    for i in range(nq):
        t0 = time.perf_counter(); sr.retrieve_arrays(qs[i % len(qs)], k); lat.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    words = 1024 * 40
    buf = np.zeros(words, dtype=np.uint64)
    lib = sr.backend._lib
    sr.backend._check(lib.svsb_xchg_read_stamps(sr.backend.engine._h, buf.ctypes.data, words))
    st = buf.reshape(1024, 40).astype(np.int64)
    st = st[st[:, 16] > 20]                                      # entries of the timed queries (seq > warm-up)
    sel = (st[:, 6] - st[:, 0]) / 1e3                            # selection kernel incl. push, us
    push = (st[:, 6] - st[:, 5]) / 1e3                           # its epilogue: result + record stores into `world` windows + flags
    gap = (st[:, 17] - st[:, 6]) / 1e3                           # selection done -> merge kernel running
    seen = st[:, 20:20 + world]
    wait_last = (seen.max(axis=1) - st[:, 17]) / 1e3             # merge start -> last peer's flag seen
    own = (seen[:, rank] - st[:, 17]) / 1e3                      # ... -> own flag seen (already published: the polling cost)
    merge = (st[:, 18] - st[:, 17]) / 1e3
    mine = {"rank": rank, "sel": [pct(sel, 50), pct(sel, 90)], "push": [pct(push, 50), pct(push, 90)], "gap": [pct(gap, 50), pct(gap, 90)],
            "wait_last": [pct(wait_last, 50), pct(wait_last, 90), pct(wait_last, 99)], "own": [pct(own, 50), pct(own, 90)],
            "merge": [pct(merge, 50), pct(merge, 90)], "lat": [pct(np.array(lat) * 1e6, 50), pct(np.array(lat) * 1e6, 90)], "entries": int(len(st))}
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    sr.close()
    if rank == 0:
        print(f"peer_latency world={world} rows={n} dims={d} k={k}: {nq} synchronous svsb_query_peer calls, stamps of the last {mine['entries']} (us; p50 / p90 [/ p99])")
        print("rank | selection kernel (incl. push) | push epilogue | selection end -> merge running | merge start -> LAST peer flag seen | -> own flag seen | merge kernel | host latency of the call")
        for r in allr:
            print(f"{r['rank']:>4} | {r['sel'][0]:6.1f} / {r['sel'][1]:6.1f} | {r['push'][0]:5.2f} / {r['push'][1]:5.2f} | {r['gap'][0]:5.1f} / {r['gap'][1]:5.1f} | "
                  f"{r['wait_last'][0]:6.1f} / {r['wait_last'][1]:6.1f} / {r['wait_last'][2]:6.1f} | {r['own'][0]:5.2f} / {r['own'][1]:5.2f} | "
                  f"{r['merge'][0]:5.1f} / {r['merge'][1]:5.1f} | {r['lat'][0]:6.1f} / {r['lat'][1]:6.1f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
