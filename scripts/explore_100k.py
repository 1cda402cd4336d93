"""Exploration on the GPU box: TNT traces and staircase timing on the synthetic 100k problem.
usage: explore_100k.py [n_poses] [precon 1|3] [outer] [mode tnt|solve] [init odom|warm|random]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cora_b200 import capi, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 1
outer = int(sys.argv[3]) if len(sys.argv) > 3 else 60
mode = sys.argv[4] if len(sys.argv) > 4 else "tnt"
l, m, d, r = max(10, n // 10000), n // 5, 3, 5
t0 = time.time()
arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=42)
Q = capi.assemble(d, n, l, arrays)
m = len(arrays["rg_w"])
N = d * n + m + n + l
print("assembled N=%d nnz=%d in %.2fs" % (N, Q.nnz, time.time() - t0), flush=True)
t0 = time.time()
h = capi.Handle(d, n, m, n + l, Q, preconditioner=pre)
print("handle (precon %d) in %.2fs" % (pre, time.time() - t0), flush=True)
init = sys.argv[5] if len(sys.argv) > 5 else "odom"
x0 = (synthetic.odometry_initialization(d, n, l, arrays, r, seed=0) if init == "odom" else
      synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=0) if init == "warm" else
      np.asfortranarray(np.random.default_rng(0).uniform(-1, 1, size=(N, r))))
if mode == "tnt":
    x0 = h.project_to_manifold(x0)
    h.set_iterate(x0)
    prm = capi.default_tnt_params(max_iterations=outer, max_computation_time=0.0)
    res = h.tnt_resident(prm)
    print("status", res.status, "f", res.f, "gnorm", res.gradfx_norm, "device_time", res.device_time,
          "launches", res.kernel_launches)
    tt = np.diff(np.array(res.time))
    for i, (it, dt, f, g, D, rho) in enumerate(zip(res.inner_iterations, tt, res.objective_values, res.gradient_norms,
                                                  res.trust_region_radius, res.gain_ratios)):
        print("%3d inner %3d dt %.3f ms (%.1f us/CG) f %.6e g %.3e Delta %.3e rho %.3f" % (
            i, it, dt * 1e3, dt * 1e6 / max(it, 1), f, g, D, rho))
    tot = sum(res.inner_iterations)
    print("total CG", tot, "CG it/s", tot / res.device_time)
else:
    t0 = time.time()
    out = h.solve(x0, max_rank=10, params=capi.default_tnt_params(max_computation_time=0.0), verbose=False)
    print("solve wall %.3fs" % (time.time() - t0))
    for s in out["stages"]:
        print(s)
    print({k: v for k, v in out.items() if k not in ("x", "stages")})
