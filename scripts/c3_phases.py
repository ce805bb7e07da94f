"""torchrun: where a sharded 1024-query batch spends its time, per phase, with CUDA events on the batch's stream
(rank 0's view and the max over ranks).  Phases of the global-threshold path (svs_b200/sharded.py _batch):
sample maxima | all-gather (b x 32 floats) | union threshold + filter pass + exact refine | all-gather (records) | verifying merge.

    python -m torch.distributed.run --nproc-per-node N scripts/c3_phases.py [rows] [dims] [k] [batch] [iters] [peer|collective]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_b200.sharded import ShardedRetriever  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
b = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 30

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
exchange = sys.argv[6] if len(sys.argv) > 6 else "peer"
sr = ShardedRetriever(rank, world, local, exchange=exchange)
sr.load_synthetic(n, d, seed=0, id0=1, id_step=1)
rng = np.random.default_rng(2)
q = rng.standard_normal((b, d)).astype(np.float32)
q /= np.sqrt((q * q).sum(axis=1))[:, None]
sr.set_queries(q)
plan = sr._global_plan(k)
for _ in range(30):
    sr.run_batch(k)
torch.cuda.synchronize(); dist.barrier()
be = sr.backend
dq = sr._queries
rec, gath, (o_s, o_i, o_c) = sr._bufs[("batch", k, b, k if plan is None else plan[2])]
names = ["sample maxima", "all-gather tops", "threshold+filter+refine", "all-gather records", "merge"]
acc = np.zeros(len(names))
if plan is not None and sr._batch_peer_ready:
    names, acc = [], np.zeros(0)      # fused exchange: no phase boundaries on the host side
elif plan is not None:
    tops, tops_all = sr._bufs[("tops", b)]
    for it in range(iters):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()
        be.batch_sample_tops(dq, k, plan[1], tops); ev[1].record()
        dist.all_gather_into_tensor(tops_all.view(-1), tops.view(-1)); ev[2].record()
        be.batch_global_records(dq, k, tops_all, world, plan[0], plan[2], rec); ev[3].record()
        dist.all_gather_into_tensor(gath.view(-1), rec.view(-1)); ev[4].record()
        be.enqueue_merge_verified(gath, world, b, plan[2], k, min(k, n), o_s, o_i, o_c); ev[5].record()
        torch.cuda.synchronize()
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(5)])
else:
    names = ["local records", "all-gather records", "merge"]
    acc = np.zeros(3)
    for it in range(iters):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        be.batch_local(dq, k, rec); ev[1].record()
        dist.all_gather_into_tensor(gath.view(-1), rec.view(-1)); ev[2].record()
        be.enqueue_merge(gath, world, b, k, o_s, o_i, o_c); ev[3].record()
        torch.cuda.synchronize()
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
acc = acc / iters * 1e3
t = torch.tensor(acc, device="cuda")
mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
# back-to-back batches (what bench.py times)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
e0.record()
t_host0 = time.perf_counter()
for _ in range(iters):
    sr.run_batch(k, defer=True)
sr.flush_batches()
t_host = (time.perf_counter() - t_host0) / iters * 1e6          # host time to ENQUEUE a batch (nothing synchronises)
e1.record(); torch.cuda.synchronize()
loop = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device="cuda"); dist.all_reduce(loop, op=dist.ReduceOp.MAX)
unanswered = sr.last_batch_unanswered()
if rank == 0:
    print(f"rows={n} d={d} k={k} b={b} world={world} plan={plan} exchange={'fused peer' if sr._batch_peer_ready else exchange}")
    for nm, a, m in zip(names, acc, mx.tolist()):
        print(f"  {nm:28s} rank0 {a:8.1f} us   max over ranks {m:8.1f} us")
    print(f"  host enqueue time per batch (rank 0): {t_host:.1f} us")
    print(f"  sum of phases (rank 0) {acc.sum():.1f} us; back-to-back loop {loop.item():.1f} us per batch = {b / loop.item() * 1e6:.0f} queries/s; unanswered {unanswered}")
sr.close()
dist.destroy_process_group()
