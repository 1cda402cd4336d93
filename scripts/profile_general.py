"""ncu driver: a few TNT outer iterations with the general sparse factor as preconditioner (TIERS)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cora_b200 import capi
name = sys.argv[1] if len(sys.argv) > 1 else "tiers"
outer = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
arrays = {k: g[k] for k in g.files if k not in ("d", "n", "l")}
d, n, l = int(g["d"]), int(g["n"]), int(g["l"])
Q = capi.assemble(d, n, l, arrays)
m = len(arrays["rg_w"])
r = d + 1
h = capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_REG_CHOLESKY)
x0 = h.project_to_manifold(np.random.default_rng(0).uniform(-1, 1, size=(Q.shape[0], r)))
h.set_iterate(x0)
res = h.tnt_resident(capi.default_tnt_params(max_iterations=20, max_computation_time=0.0))
rt = torch.cuda.cudart()
torch.cuda.synchronize()
rt.cudaProfilerStart()
t0 = time.perf_counter()
res = h.tnt_resident(capi.default_tnt_params(max_iterations=outer, max_computation_time=0.0, Delta0=res.trust_region_radius[-1]))
torch.cuda.synchronize()
t1 = time.perf_counter()
rt.cudaProfilerStop()
print("outer %d, CG %s, launches %d, device %.3f ms, wall %.3f ms" % (len(res.inner_iterations), res.inner_iterations, res.kernel_launches, 1e3 * res.device_time, 1e3 * (t1 - t0)))
