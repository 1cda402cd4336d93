"""The bench's solve-to-certificate leg run several times in one process (stage times, verbose certification):
how much of `solve_to_cert.seconds` is the refinement stage's eigen-search and how much it varies.
usage: python scripts/solve_leg_repeat.py [reps=3]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cora_b200 import capi

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
w = bench.WORKLOAD
d, n, l = w["d"], w["n"], w["l"]
arrays, gt, Q, m = bench.build_problem()
h = capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_REG_CHOLESKY)
x0 = h.project_to_manifold(bench.initial_guess(arrays, gt, 0, "warm"))
for i in range(reps):
    torch.cuda.synchronize()
    ts = time.perf_counter()
    out = h.solve(x0, max_rank=7, params=capi.default_tnt_params(max_computation_time=0.0), verbose=(i == reps - 1))
    torch.cuda.synchronize()
    print("solve %d: %.3f s  f %.9f  stages %s" % (i, time.perf_counter() - ts, out["f"],
          [(s["rank"], s["cg"], round(s["tnt_seconds"], 3), round(s["cert_seconds"], 4), s.get("cert_branch")) for s in out["stages"]]), flush=True)
