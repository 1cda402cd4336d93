"""Per-level profile and wall time of the RegularizedCholesky apply on a general graph (CORA_B200_GEN_PROFILE=1)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cora_b200 import capi
name = sys.argv[1] if len(sys.argv) > 1 else "tiers"
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
arrays = {k: g[k] for k in g.files if k not in ("d", "n", "l")}
d, n, l = int(g["d"]), int(g["n"]), int(g["l"])
Q = capi.assemble(d, n, l, arrays)
m = len(arrays["rg_w"])
with capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
    V = np.random.default_rng(0).standard_normal((Q.shape[0], d + 1))
    for _ in range(3):
        t0 = time.perf_counter(); Z = h.precondition(V); print("precondition wall %.1f us" % (1e6 * (time.perf_counter() - t0)))
