import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_fixture, make_handle
from oracle import cora_oracle as co
from cora_b200 import capi
g, p = load_fixture("small_ra_slam_problem")
p.preconditioner = co.JACOBI
p.update_problem_data()
p.rank = 3
x0 = p.random_initial_guess(np.random.default_rng(0))
ref = co.problem_tnt(p, x0, co.cora_tnt_params(max_iterations=8))
with make_handle(p) as h:
    got = h.tnt(x0, capi.default_tnt_params(max_iterations=8, max_computation_time=0.0))
for name in ["objective_values", "gradient_norms", "preconditioned_gradient_norms", "trust_region_radius",
             "inner_iterations", "update_step_norms", "update_step_M_norms", "gain_ratios"]:
    print(name)
    print("  ref", np.array(getattr(ref, name)))
    print("  got", np.array(getattr(got, name)))
