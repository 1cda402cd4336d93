"""BASELINE configs[4]: synthetic 1M-pose SE(3) graph, rank 5 -- HBM-roofline sweep of the data-matrix product
(Problem::dataMatrixProduct) and of full CG iterations, on one GPU (run N replicas for N GPUs: the path does not
shard a single solve).  Prints one JSON line.
usage: sweep_1m.py [n_poses=1000000] [reps=50] [outer=2]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cora_b200 import capi, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
outer = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d, r = 3, 5
l, m = max(10, n // 10000), n // 5
t0 = time.time()
arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=42)
Q = capi.assemble(d, n, l, arrays)
m = len(arrays["rg_w"])
N = d * n + m + n + l
t_asm = time.time() - t0
t0 = time.time()
h = capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_JACOBI)
t_create = time.time() - t0
x0 = h.project_to_manifold(synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=0))
h.set_iterate(x0)
peak = 6550.4
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
V = 8 * N * r
q = 12 * Q.nnz + 4 * (N + 1)
h.spmm_resident(3)
ms = h.spmm_resident(reps)
t_spmm = ms * 1e-3 / reps
b_spmm = q + 2 * V
# CG iterations: warm start, 12 untimed outer iterations, then `outer` timed ones
pre = h.tnt_resident(capi.default_tnt_params(max_iterations=12, max_computation_time=0.0))
res = h.tnt_resident(capi.default_tnt_params(max_iterations=outer, max_computation_time=0.0,
                                             Delta0=pre.trust_region_radius[-1]))
cg = int(sum(res.inner_iterations))
prof, grid, _ = h.phase_profile()
b_cg = q + 15 * V + 8 * N
b_outer = 2 * q + 19 * V + 8 * N
line = {"workload": "synthetic %d-pose SE(3) + %d ranges, %d landmarks, rank %d (BASELINE configs[4])" % (n, m, l, r),
        "N": N, "nnz": int(Q.nnz), "assemble_s": t_asm, "create_s": t_create, "grid": grid,
        "spmm": {"us": 1e6 * t_spmm, "algorithmic_bytes": b_spmm, "achieved_gbs": b_spmm / t_spmm / 1e9,
                 "frac_of_measured_hbm": b_spmm / t_spmm / 1e9 / peak, "reps": reps},
        "cg": {"iterations": cg, "outer": len(res.inner_iterations), "device_s": res.device_time,
               "cg_it_per_s": cg / res.device_time,
               "achieved_gbs": (cg * b_cg + len(res.inner_iterations) * b_outer) / res.device_time / 1e9,
               "frac_of_measured_hbm": (cg * b_cg + len(res.inner_iterations) * b_outer) / res.device_time / 1e9 / peak,
               "phases_us": {k: v[0] / v[1] for k, v in prof.items() if v[1]}},
        "peak_gbs": peak}
print(json.dumps(line))
