"""Run under torchrun on >= 2 GPUs: the row-sharded retrieve with the fused peer exchange (default) must give the
same bits as the NCCL all-gather exchange, the same answer on every rank, and match the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        scripts/sharded_check.py [rows] [dims]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import svs_oracle as oracle                                      # checker only
from svs_b200.sharded import ShardedRetriever


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_003
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = oracle.synth_matrix_uniform(n, d, 31)
    m[5] = m[n - 2]                                              # exact tie across the first and the last shard
    ids = np.cumsum(np.random.default_rng(32).integers(1, 4, size=n)).astype(np.int64)
    qs = oracle.synth_queries(48, d, 33)
    qs[3] = m[5]
    peer = ShardedRetriever(rank, world, local, exchange="peer")
    coll = ShardedRetriever(rank, world, local, exchange="collective")
    peer.load_global(m, ids)
    coll.load_global(m, ids)
    assert peer.exchange == "peer"
    checked = 0
    for k in (1, 100, 1000, 2048):
        # synchronous single queries (svsb_query_peer)
        for j, q in enumerate(qs[:12]):
            a = peer.retrieve(q, k)
            b = coll.retrieve(q, k)
            assert a == b, f"rank {rank}: peer and collective exchange differ (k={k}, query {j})"
            if j % 4 == 0:
                oracle.compare_retrieval(a, oracle.superheavy(m, ids, q, k), oracle.scores_of(m, q), ids)
            checked += 1
        assert [x[1] for x in peer.retrieve(qs[3], 2)] == [int(ids[5]), int(ids[n - 2])]
        # device-resident pipelined loop (svsb_enqueue_query_peer on the side stream)
        peer.set_queries(qs)
        coll.set_queries(qs)
        for sr in (peer, coll):
            sr.run_queries(k, len(qs), time_gemv=True)
        torch.cuda.synchronize()
        nb = 16                                                  # the last micro-batch's outputs are still in the buffers
        ps, pi, pc = [t[:nb].cpu().numpy() for t in peer._buffers(k)[2]]
        cs, ci, cc = [t[:nb].cpu().numpy() for t in coll._buffers(k)[2]]
        assert np.array_equal(ps.view(np.uint32), cs.view(np.uint32)) and np.array_equal(pi, ci) and np.array_equal(pc, cc)
        # identical on every rank
        t = torch.from_numpy(pi.copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref)
    peer.close()
    coll.close()
    dist.barrier()
    if rank == 0:
        print(f"sharded_check ok: world={world} rows={n} dims={d} queries checked={checked}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
