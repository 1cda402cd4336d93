"""Run the staircase on a dataset through the C-ABI and print the stage table."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_dataset, make_handle
from cora_b200 import capi
name = sys.argv[1] if len(sys.argv) > 1 else "plaza2"
r0 = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pre = int(sys.argv[3]) if len(sys.argv) > 3 else capi.PRECON_REG_CHOLESKY
p = load_dataset(name, preconditioner=pre)
p.update_problem_data()
x0 = np.random.default_rng(0).uniform(-1, 1, size=(p.N, r0))
with make_handle(p, preconditioner=pre) as h:
    t = time.time()
    out = h.solve(x0, max_rank=10, params=capi.default_tnt_params(max_computation_time=0.0), verbose=True)
    print("wall", time.time() - t)
for s in out["stages"]:
    print(s)
print({k: v for k, v in out.items() if k not in ("x", "stages")})
