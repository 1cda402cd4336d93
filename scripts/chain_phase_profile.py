"""Per-phase times of the persistent TNT kernel with RegularizedCholesky (chain factor applied as phases).
usage: CORA_B200_PHASE_PROFILE=1 python scripts/chain_phase_profile.py [n_poses=100000] [outer=6] [pre_outer=12]"""
import os, sys
import numpy as np
os.environ.setdefault("CORA_B200_PHASE_PROFILE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cora_b200 import capi, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
outer = int(sys.argv[2]) if len(sys.argv) > 2 else 6
pre_outer = int(sys.argv[3]) if len(sys.argv) > 3 else 12
l, m, d, r = max(10, n // 10000), n // 5, 3, 5
arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=42)
Q = capi.assemble(d, n, l, arrays)
m = len(arrays["rg_w"])
h = capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_REG_CHOLESKY)
x0 = h.project_to_manifold(synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=0))
h.set_iterate(x0)
res = h.tnt_resident(capi.default_tnt_params(max_iterations=pre_outer, max_computation_time=0.0))
prm = capi.default_tnt_params(max_iterations=outer, max_computation_time=0.0, Delta0=res.trust_region_radius[-1])
res = h.tnt_resident(prm)
cg = int(np.sum(res.inner_iterations))
print("outer %d, CG %d, device_time %.3f ms, %.1f us per CG iteration (outer work included), f %.9e" % (
    len(res.inner_iterations), cg, 1e3 * res.device_time, 1e6 * res.device_time / max(cg, 1), res.f))
prof, grid, bars = h.phase_profile()
mx = h.phase_profile_ctas()
print("grid %d, barriers %d" % (grid, bars))
for k, (tot, cnt) in prof.items():
    if cnt:
        print("  %-10s total %10.1f us  count %7d  avg %8.2f us  per CG it %8.2f us   slowest/median CTA avg %s" % (
            k, tot, cnt, tot / cnt, tot / max(cg, 1), mx.get(k)))
