"""Debug driver for the streaming kernels: SPMM through spmm_resident vs scipy, TNT traces vs the oracle."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_handle
from synth import make_synthetic
from oracle import cora_oracle as co
from cora_b200 import capi, synthetic
d = int(sys.argv[1]) if len(sys.argv) > 1 else 3
r = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n, l, m = 300, 3, 120
p = make_synthetic(n=n, l=l, m=m, d=d, seed=12, rank=r)
p.update_problem_data()
arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=12)
x0 = p.project_to_manifold(synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=2))
with make_handle(p) as h:
    h.set_iterate(x0)
    h.spmm_resident(1)
    got = h.get_work_vector(1, r)
    ref = p.Q @ x0
    err = np.abs(got - ref)
    print("SPMM max err %.3e (scale %.3e)" % (err.max(), np.abs(ref).max()))
    bad = np.argwhere(err > 1e-9 * np.abs(ref).max())
    print("bad rows (ref order):", sorted(set(bad[:, 0].tolist()))[:40], "of N", p.Q.shape[0], "d*n", d * n, "m", p.m)
    res = h.tnt(x0, capi.default_tnt_params(max_iterations=4, max_computation_time=0.0))
oref = co.problem_tnt(p, x0, co.cora_tnt_params(max_iterations=4))
for name in ["objective_values", "gradient_norms", "preconditioned_gradient_norms", "inner_iterations", "gain_ratios"]:
    print(name, "\n  ref", np.array(getattr(oref, name)), "\n  got", np.array(getattr(res, name)))
with make_handle(p) as h:
    res = h.tnt(x0, capi.default_tnt_params(max_iterations=1, max_computation_time=0.0))
    print("status", res.status, res.objective_values, res.gradient_norms, res.preconditioned_gradient_norms)
    # after one accepted outer iteration the roles were swapped: X holds x1; compare what can be compared at x1
    x1 = h.get_work_vector(100 + 0, r)
    G = h.get_work_vector(100 + 1, r); GR = h.get_work_vector(100 + 2, r); PG = h.get_work_vector(100 + 3, r)
    eg = p.Q @ x1
    rg = p.tangent_space_projection(x1, eg)
    pg = p.tangent_space_projection(x1, p.precondition(rg))
    for nm, a, b in (("QX", G, eg), ("grad", GR, rg), ("pgrad", PG, pg)):
        e = np.abs(a - b)
        rows = sorted(set(np.argwhere(e > 1e-8 * np.abs(b).max())[:, 0].tolist()))
        print(nm, "max err %.3e scale %.3e  bad rows %d: %s" % (e.max(), np.abs(b).max(), len(rows), rows[:30]))
