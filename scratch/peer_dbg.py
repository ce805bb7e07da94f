import os, sys, threading, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle"); sys.path.insert(0, "/root/repo/tests")
os.environ["SVSB_XCHG_TIMEOUT_MS"] = "3000"
import numpy as np
import svs_oracle as oracle
from svs_b200.sharded import CudaShardBackend, partition

def backs_for(m, ids, world):
    backs = []
    for r in range(world):
        b = CudaShardBackend(0)
        row0, cnt = partition(len(m), world, r)
        b.set_shard(row0)
        b.load_rows(np.ascontiguousarray(m[row0:row0 + cnt]), np.ascontiguousarray(ids[row0:row0 + cnt]))
        backs.append(b)
    for r, b in enumerate(backs):
        b.exchange_handle(world, r)
    for b in backs:
        b.exchange_connect_local(backs)
    return backs

n, d, k = 20_000, 128, 50
m = oracle.synth_matrix_uniform(n, d, 25)
ids = np.arange(10, 10 + n, dtype=np.int64)
qs = oracle.synth_queries(4, d, 26)
for rows, world in ((n, 2), (2, 3), (n, 1)):
    backs = backs_for(m[:rows], ids[:rows], world)
    log = []
    def work(r):
        for j, q in enumerate(qs):
            t0 = time.time()
            try:
                s, i = backs[r].query_peer(q, k)
                log.append((rows, world, r, j, "ok", len(s), round(time.time() - t0, 4)))
            except Exception as ex:
                log.append((rows, world, r, j, "ERR", str(ex)[:90], round(time.time() - t0, 4)))
                break
    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]; [t.join() for t in ts]
    for l in log: print(l, flush=True)
    for b in backs: b.close()
