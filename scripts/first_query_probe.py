"""Where does the first retrieve of a fresh process go?  Times library load, engine creation (CUDA context), the first
small load (pinned slabs, device buffers) and the first query (lazy kernel loading), then the same steps again."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t = [time.perf_counter()]
def lap(label):
    t.append(time.perf_counter()); print(f"{label}: {1e3 * (t[-1] - t[-2]):.1f} ms", flush=True)
import svs_b200
from svs_b200 import _lib
lap("import svs_b200")
_lib.load(); lap("dlopen libsvsb200.so + bind symbols")
e = svs_b200.Engine([0]); lap("Engine([0]) = svsb_create (CUDA context)")
rng = np.random.default_rng(0)
m = rng.standard_normal((10_548, 1536)).astype(np.float32); m /= np.sqrt((m * m).sum(axis=1))[:, None]
lap("(host matrix)")
e.load(m); lap("first load 10,548 x 1536 (pinned slabs, device buffers, norm kernel)")
e.query(m[3], 10); lap("first query (workspaces, lazy kernel load)")
e.query(m[4], 10); lap("second query")
e.load(m); lap("second load")
e.close(); lap("close")
e = svs_b200.Engine([0]); lap("second Engine([0])")
e.load(m); lap("load on the second engine")
e.query(m[3], 10); lap("its first query")
e.close()
