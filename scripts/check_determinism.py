import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_handle
from synth import make_synthetic
from oracle import cora_oracle as co
from cora_b200 import capi, synthetic
r = int(sys.argv[1]) if len(sys.argv) > 1 else 11
d, n, l, m = 3, 600, 3, 200
p = make_synthetic(n=n, l=l, m=m, d=d, seed=9, rank=r); p.update_problem_data()
arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=9)
x0 = p.project_to_manifold(synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=1))
outs = []
for k in range(3):
    with make_handle(p) as h:
        got = h.tnt(x0, capi.default_tnt_params(max_iterations=7, max_computation_time=0.0))
    outs.append(np.array(got.objective_values))
print("run0 == run1:", np.array_equal(outs[0], outs[1]), " run0 == run2:", np.array_equal(outs[0], outs[2]))
print(outs[0][-3:], outs[1][-3:])
# operator-level check at this rank: hessvec through tier-1 (old kernels) vs oracle, and one TNT step with 1 CG iteration
ref = co.problem_tnt(p, x0, co.cora_tnt_params(max_iterations=6, max_TPCG_iterations=3))
with make_handle(p) as h:
    got = h.tnt(x0, capi.default_tnt_params(max_iterations=6, max_TPCG_iterations=3, max_computation_time=0.0))
print("3-CG-cap rel:", np.array(got.objective_values) / np.array(ref.objective_values) - 1)
