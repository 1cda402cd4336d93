"""solveCORA on the host cores at full problem sizes -- TEST / BASELINE INFRASTRUCTURE ONLY.

The staircase driver, certification, saddle escape and rounding are the NumPy/SciPy oracle's
(oracle/cora_oracle.py: solve_cora, src/CORA.cpp:26-441); the truncated-Newton solves -- where the time goes -- run
in the C++ restatement of the reference's CPU path (oracle/cpu_ref.cpp) with the reference's default preconditioner
(RegularizedCholesky).  Used by bench.py (`solve_to_cert.cpu_seconds`) and by the 1e-6 parity tests; nothing under
cora_b200/ imports it.
"""
import time

import numpy as np

from . import cora_oracle as co
from . import cpu_ref


def _params_c(params):
    """oracle TNTParams -> the C struct cpu_ref.tnt takes (cora_b200.capi.TntParams: struct definition only)."""
    from cora_b200 import capi
    return capi.default_tnt_params(
        Delta0=params.Delta0, eta1=params.eta1, eta2=params.eta2, alpha1=params.alpha1, alpha2=params.alpha2,
        max_TPCG_iterations=params.max_TPCG_iterations, max_iterations=params.max_iterations,
        kappa_fgr=params.kappa_fgr, theta=params.theta,
        preconditioned_gradient_tolerance=params.preconditioned_gradient_tolerance,
        gradient_tolerance=params.gradient_tolerance,
        relative_decrease_tolerance=params.relative_decrease_tolerance,
        stepsize_tolerance=params.stepsize_tolerance, Delta_tolerance=params.Delta_tolerance,
        max_computation_time=0.0)


def solve_cora_cpu(problem, x0, max_rank=20, params=None, threads=None, verbose=False):
    """Returns (oracle CoraResult, dict(tnt_seconds, total_seconds, threads))."""
    params = params or co.cora_tnt_params()
    if problem.preconditioner == co.REG_CHOLESKY:
        R = cpu_ref.CpuRef(problem.d, problem.n, problem.m, problem.n + problem.l, problem.Q, preconditioner=3,
                           reg_lambda=problem.lambda_reg, threads=threads)
    else:
        R = cpu_ref.CpuRef(problem.d, problem.n, problem.m, problem.n + problem.l, problem.Q, preconditioner=1,
                           threads=threads)
    t_tnt = [0.0]
    pc = _params_c(params)

    def tnt_fn(prob, X, prm):
        t0 = time.perf_counter()
        r = R.tnt(np.asfortranarray(X), pc)
        t_tnt[0] += time.perf_counter() - t0
        return r   # capi.TntResult carries the fields solve_cora reads: x, f, gradfx_norm, status, inner_iterations

    t0 = time.perf_counter()
    out = co.solve_cora(problem, x0, max_rank=max_rank, params=params, verbose=verbose, tnt_fn=tnt_fn)
    total = time.perf_counter() - t0
    used = R.threads
    R.close()
    return out, dict(tnt_seconds=t_tnt[0], total_seconds=total, threads=used)
