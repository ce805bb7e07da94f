"""One process, one GPU, world_size 1: the pipelined peer-exchange query loop (svsb_enqueue_query_peer) on a 1M x 1536
shard, so that ncu can capture select_topk_kernel (with the fused push) and merge_window_kernel."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_b200.sharded import ShardedRetriever

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29631")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sr = ShardedRetriever(0, 1, 0, exchange="peer")
sr.load_synthetic(n, 1536, seed=0, id0=1, id_step=1)
rng = np.random.default_rng(1)
q = rng.random((64, 1536), dtype=np.float32)
q /= np.sqrt((q * q).sum(axis=1))[:, None]
sr.set_queries(q)
for _ in range(3):
    sr.run_queries(100, 32)
torch.cuda.synchronize()
print("peer_profile done", sr.retrieve(q[0], 3))
sr.close()
dist.destroy_process_group()
