"""The global-threshold batch protocol with WORLD virtual ranks on ONE GPU (shard engines side by side, all-gathers by
torch.stack): rank 0's kernels see exactly the thresholds / candidate counts of a real WORLD-GPU run, so `ncu` on this
script gives the per-kernel launch list of one rank's batch.

    python scripts/c3_virtual_ranks.py [rows] [dims] [k] [batch] [world] [iters]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_b200.sharded import CudaShardBackend, partition  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
b = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
world = int(sys.argv[5]) if len(sys.argv) > 5 else 8
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5

backs = []
for r in range(world):
    be = CudaShardBackend(0)
    row0, cnt = partition(n, world, r)
    be.set_shard(row0)
    be.load_synthetic(cnt, d, 0, 1, 1)
    backs.append(be)
rng = np.random.default_rng(2)
q = rng.standard_normal((b, d)).astype(np.float32)
q /= np.sqrt((q * q).sum(axis=1))[:, None]
dq = backs[0].device_queries(q)
probes = [be.batch_global_probe(k) for be in backs]
assert all(p[0] for p in probes), probes
f = max(min(1.0, p[1] / p[2]) for p in probes)
lam = min(k, n) * f
rank = int(np.ceil(lam + 6.0 * np.sqrt(lam) + 4.0))
share = min(k, n) / world
cap = min(k, int(np.ceil(share + 6.0 * np.sqrt(share) + 4.0)))
norm = max(p[3] for p in probes)
tops = [be.new_tops(b) for be in backs]
recs = [be.new_records(b, cap) for be in backs]
o_s, o_i, o_c = backs[0].new_outputs(b, k)
names = ["sample maxima", "threshold+filter+refine", "merge"]
acc = np.zeros(3)
for it in range(iters + 2):
    for r in range(1, world):
        backs[r].batch_sample_tops(dq, k, norm, tops[r])
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    backs[0].batch_sample_tops(dq, k, norm, tops[0]); ev[1].record()
    torch.cuda.synchronize()
    tops_all = torch.stack(tops, dim=0).contiguous()
    for r in range(1, world):
        backs[r].batch_global_records(dq, k, tops_all, world, rank, cap, recs[r])
    torch.cuda.synchronize()
    ev[1].record()
    backs[0].batch_global_records(dq, k, tops_all, world, rank, cap, recs[0]); ev[2].record()
    torch.cuda.synchronize()
    gathered = torch.stack(recs, dim=0).contiguous()
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    backs[0].enqueue_merge_verified(gathered, world, b, cap, k, min(k, n), o_s, o_i, o_c); e3.record()
    torch.cuda.synchronize()
    if it >= 2:
        acc += np.array([ev[0].elapsed_time(ev[1]) if False else 0.0, ev[1].elapsed_time(ev[2]), e2.elapsed_time(e3)])
cand, resc, flags = backs[0].engine.batch_stats(b)
cnt = o_c.cpu().numpy()
print(f"rows={n} d={d} k={k} b={b} virtual world={world} order statistic {rank} record cap {cap}: rank 0 threshold+filter+refine "
      f"{acc[1] / iters * 1e3:.1f} us, merge {acc[2] / iters * 1e3:.1f} us; candidates mean {cand.mean():.0f} max {cand.max()}, "
      f"re-scored mean {resc.mean():.0f} max {resc.max()}; unanswered {(cnt < 0).sum()}")
for be in backs:
    be.close()
