"""Timing record of the general sparse block Cholesky (loop closures / several robots) and of the
implicit formulation (8f-3) on TIERS, MR.CLAM 2 and the synthetic 100k-pose variant B (chain + n/10 loop closures):
factor structure, preconditioner apply, PSD test, TNT with RegularizedCholesky vs Jacobi, solve-to-certificate.
usage: general_graphs.py [out.json]   (reads tests/golden/{tiers,mrclam2}.npz; nothing under /root/reference)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402  (CUDA events)
from cora_b200 import capi, synthetic  # noqa: E402


def dataset(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    arrays = {k: g[k] for k in g.files if k not in ("d", "n", "l")}
    return int(g["d"]), int(g["n"]), int(g["l"]), arrays


def loops_near(n, k, seed, near):
    rng = np.random.default_rng(seed)
    out = set()
    while len(out) < k:
        i = int(rng.integers(0, n))
        j = int(np.clip(i + rng.integers(-near, near + 1), 0, n - 1))
        if abs(i - j) > 1:
            out.add((min(i, j), max(i, j)))
    return sorted(out)


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def run(name, d, n, l, arrays, r, solve=True):
    rec = {"workload": name, "d": d, "n_poses": n, "n_landmarks": l}
    Q = capi.assemble(d, n, l, arrays)
    m = len(arrays["rg_w"])
    rec.update(n_ranges=m, N=Q.shape[0], nnz=int(Q.nnz))
    rec["factor_structure"] = capi.debug_factor_stats(d, n, m, n + l, Q)
    prm = lambda **kw: capi.default_tnt_params(max_computation_time=0.0, **kw)
    rng = np.random.default_rng(0)
    t0 = time.perf_counter()
    h = capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_REG_CHOLESKY)
    rec["create_with_reg_cholesky_s"] = time.perf_counter() - t0
    with h:
        assert h.effective_preconditioner == capi.PRECON_REG_CHOLESKY
        N = Q.shape[0]
        x0 = h.project_to_manifold(rng.uniform(-1, 1, size=(N, r)))
        # preconditioner apply and PSD test on resident data
        h.set_iterate(x0)
        lam = h.reg_lambda
        t = timed(lambda: setattr(h, "reg_lambda", lam), 5)
        rec["refactor_ms"] = 1e3 * t
        t = timed(lambda: h.psd_test(1e-3, r=r), 5)
        rec["psd_test_ms"] = 1e3 * t
        for pre, key in ((capi.PRECON_REG_CHOLESKY, "reg_cholesky"), (capi.PRECON_JACOBI, "jacobi")):
            h.set_preconditioner(pre)
            h.set_iterate(x0)
            res = h.tnt_resident(prm(max_iterations=30))
            cg = int(np.sum(res.inner_iterations))
            rec["tnt_30_outer_" + key] = {"cg_iterations": cg, "device_s": res.device_time, "f": res.f,
                                          "us_per_cg_iteration": 1e6 * res.device_time / max(cg, 1),
                                          "kernel_launches": int(res.kernel_launches)}
        h.set_preconditioner(capi.PRECON_REG_CHOLESKY)
        if solve:
            t0 = time.perf_counter()
            out = h.solve(x0, max_rank=r + 5, params=prm())
            rec["solve_to_cert"] = {"seconds": time.perf_counter() - t0, "f": out["f"], "certified": out["certified"],
                                    "lifted_rank": out["lifted_rank"], "cg_iterations": out["total_cg_iterations"],
                                    "stages": [{k: s[k] for k in ("rank", "status", "outer", "cg", "certified", "cert_branch",
                                                                  "tnt_seconds", "cert_seconds")} for s in out["stages"]]}
            # the same solve in the implicit formulation (translations marginalised)
            h.set_formulation(capi.FORMULATION_IMPLICIT)
            k = d * n + m
            t0 = time.perf_counter()
            oi = h.solve(x0[:k], max_rank=r + 5, params=prm())
            rec["solve_to_cert_implicit"] = {"seconds": time.perf_counter() - t0, "f": oi["f"], "certified": oi["certified"],
                                             "lifted_rank": oi["lifted_rank"], "cg_iterations": oi["total_cg_iterations"]}
            h.set_formulation(capi.FORMULATION_EXPLICIT)
    return rec


def main():
    out = {"device": torch.cuda.get_device_name(0), "records": []}
    for name in ("tiers", "mrclam2"):
        d, n, l, a = dataset(name)
        out["records"].append(run(name, d, n, l, a, d + 1))
        print(json.dumps(out["records"][-1]), flush=True)
    n = 100_000
    arrays, gt = synthetic.make_arrays(n, 10, 20_000, d=3, seed=42, loop_closures=loops_near(n, n // 10, 7, 50))
    out["records"].append(run("synthetic 100k poses + 10k loop closures (BASELINE cfg3 variant B)", 3, n, 10, arrays, 5, solve=False))
    print(json.dumps(out["records"][-1]), flush=True)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "general_graphs.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
