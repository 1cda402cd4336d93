"""Cold-start staircase on the synthetic chain + landmark workload (BASELINE cfg3 shape): random U[-1,1] projection or
the odometry initialisation, reference default preconditioner, the reference's own stopping rules; prints the stage
table (rank lifts, which certificate decided) and appends a record.
usage: cold_staircase.py [n_poses] [rank0] [init=random|odom] [out.jsonl]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cora_b200 import capi, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
r0 = int(sys.argv[2]) if len(sys.argv) > 2 else 5
init = sys.argv[3] if len(sys.argv) > 3 else "random"
out_path = sys.argv[4] if len(sys.argv) > 4 else None
d, l, m = 3, max(10, n // 10000), n // 5
arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=42)
Q = capi.assemble(d, n, l, arrays)
m = len(arrays["rg_w"])
N = Q.shape[0]
with capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
    if init == "odom":
        x0 = h.project_to_manifold(synthetic.odometry_initialization(d, n, l, arrays, r0, seed=0))
    else:
        x0 = h.project_to_manifold(np.random.default_rng(0).uniform(-1, 1, size=(N, r0)))
    t = time.perf_counter()
    out = h.solve(x0, max_rank=r0 + 4, params=capi.default_tnt_params(max_computation_time=0.0))
    wall = time.perf_counter() - t
rec = {"n_poses": n, "N": int(N), "init": init, "rank0": r0, "seconds": wall, "f": out["f"], "lifted_f": out["lifted_f"],
       "lifted_rank": out["lifted_rank"], "certified": out["certified"], "refined_certified": out["refined_certified"],
       "cg_iterations": out["total_cg_iterations"],
       "stages": [{k: s[k] for k in ("rank", "status", "outer", "cg", "f", "grad", "certified", "cert_branch", "theta", "eta",
                                     "tnt_seconds", "cert_seconds")} for s in out["stages"]]}
print(json.dumps(rec))
if out_path:
    with open(out_path, "a") as fh:
        fh.write(json.dumps(rec) + "\n")
