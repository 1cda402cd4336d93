"""Profile driver (run under ncu with --profile-from-start off): builds the bench workload,
runs the untimed start-up phase, then brackets `outer` TNT outer iterations (or `spmm` data-matrix
products) with cudaProfilerStart/Stop.
usage: profile_cg.py [outer=2] [n_poses=100000] [precon=1] [pre_outer=12] [what=tnt|spmm]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cora_b200 import capi, synthetic

outer = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pre_outer = int(sys.argv[4]) if len(sys.argv) > 4 else 12
what = sys.argv[5] if len(sys.argv) > 5 else "tnt"
l, m, d, r = max(10, n // 10000), n // 5, 3, 5
arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=42)
Q = capi.assemble(d, n, l, arrays)
m = len(arrays["rg_w"])
h = capi.Handle(d, n, m, n + l, Q, preconditioner=pre)
x0 = h.project_to_manifold(synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=0))
h.set_iterate(x0)
res = h.tnt_resident(capi.default_tnt_params(max_iterations=pre_outer, max_computation_time=0.0))
prm = capi.default_tnt_params(max_iterations=outer, max_computation_time=0.0, Delta0=res.trust_region_radius[-1])
rt = torch.cuda.cudart()
torch.cuda.synchronize()
rt.cudaProfilerStart()
if what == "spmm":
    ms = h.spmm_resident(outer)
    print("spmm: %d launches, %.2f us each" % (outer, 1e3 * ms / outer))
else:
    res = h.tnt_resident(prm)
    print("profiled: outer %d, CG %s, launches %d, device_time %.3f ms" % (
        len(res.inner_iterations), res.inner_iterations, res.kernel_launches, 1e3 * res.device_time))
torch.cuda.synchronize()
rt.cudaProfilerStop()
