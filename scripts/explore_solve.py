"""Solve-to-certificate of the bench workload under several stopping rules (exploration / evidence)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cora_b200 import capi, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
l, m, d, r = max(10, n // 10000), n // 5, 3, 5
arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=42)
Q = capi.assemble(d, n, l, arrays)
m = len(arrays["rg_w"])
h = capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_REG_CHOLESKY)
x0 = h.project_to_manifold(synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=0))
for name, kw in (("reference defaults", {}),
                 ("tight", dict(relative_decrease_tolerance=0.0, stepsize_tolerance=0.0, gradient_tolerance=1e-3,
                                preconditioned_gradient_tolerance=0.0, max_iterations=250))):
    t0 = time.perf_counter()
    out = h.solve(x0, max_rank=7, params=capi.default_tnt_params(max_computation_time=0.0, **kw))
    t = time.perf_counter() - t0
    eta = min(max(out["f"] * 5e-6, 1e-7), 1e-1)
    t1 = time.perf_counter()
    cert = h.certify_solution(out["x"], eta, 10)
    tc = time.perf_counter() - t1
    print(name, "seconds %.3f f %.9f lifted %.9f certified %s refined_certified %s | psd test of refined: %s (%s) %.3fs theta %.3e eta %.3e"
          % (t, out["f"], out["lifted_f"], out["certified"], out["refined_certified"], cert.is_certified, h.last_cert_branch, tc, cert.theta, eta))
    for s in out["stages"]:
        print("   ", {k: (round(v, 6) if isinstance(v, float) else v) for k, v in s.items()})
