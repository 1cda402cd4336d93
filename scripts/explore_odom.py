"""Cold-start staircase at size: odometry initialisation (product C++ routine), RegularizedCholesky."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cora_b200 import capi, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
r0 = int(sys.argv[2]) if len(sys.argv) > 2 else 5
l, m, d = max(10, n // 10000), n // 5, 3
arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=42)
Q = capi.assemble(d, n, l, arrays)
m = len(arrays["rg_w"])
h = capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_REG_CHOLESKY)
x0 = capi.odometry_initialization(d, n, l, arrays, r0, seed=0)
t0 = time.perf_counter()
out = h.solve(x0, max_rank=r0 + 3, params=capi.default_tnt_params(max_computation_time=0.0))
t = time.perf_counter() - t0
print("odom init r0=%d: seconds %.3f f %.6f lifted %.6f (rank %d) certified %s refined_certified %s cg %d"
      % (r0, t, out["f"], out["lifted_f"], out["lifted_rank"], out["certified"], out["refined_certified"], out["total_cg_iterations"]))
for s in out["stages"]:
    print("   ", {k: (round(v, 6) if isinstance(v, float) else v) for k, v in s.items()})
