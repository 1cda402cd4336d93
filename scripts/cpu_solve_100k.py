"""CPU solve-to-certificate of the bench workload (BASELINE configs[2]) with the reference's default preconditioner:
oracle staircase driver + C++ restatement of the reference's TNT (oracle/cpu_solve.py).  Prints one JSON line.
usage: cpu_solve_100k.py [n_poses=100000] [threads=all]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cora_b200 import synthetic
from oracle import cora_oracle as co, cpu_solve

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
l, m, d, r = max(10, n // 10000), n // 5, 3, 5
arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=42)
t0 = time.time()
p = co.Problem.from_arrays(d, n, l, arrays, rank=r, preconditioner=co.REG_CHOLESKY)
p.update_problem_data()
t_setup = time.time() - t0
x0 = p.project_to_manifold(synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=0))
out, t = cpu_solve.solve_cora_cpu(p, x0, max_rank=7, threads=threads, verbose=True)
print(json.dumps({"n_poses": n, "threads": t["threads"], "setup_s": t_setup, "seconds": t["total_seconds"],
                  "tnt_seconds": t["tnt_seconds"], "f": out.result.f, "lifted_f": out.lifted_f,
                  "lifted_rank": out.lifted_rank, "certified_lifted": bool(out.stages[0]["certified"]),
                  "cg_iterations": out.total_cg_iterations, "stages": out.stages}, default=float))
