"""North-star parity: the cost reached by the CUDA path matches the CPU restatement of the reference to 1e-6
relative (BASELINE.json north_star; SURVEY F13 / BASELINE.md section 3: with the reference's own stopping rules two
correct implementations agree only to ~1e-4, so BOTH sides run with relative_decrease_tolerance = stepsize_tolerance
= 0 and a tight gradient tolerance -- src/CORA.cpp:95-109, TNT.h:561-570).

Both sides: Preconditioner::RegularizedCholesky with the same lambda, the same initial point.  Stage 1: TNT at the
lifted rank (where the relaxation certifies); stage 2: both refine at rank d from the SAME rounded
point (the oracle's projectSolution of the CPU result, src/CORA.cpp:352-441).

What "the same optimum" means: at the lifted rank the landscape of Plaza2 / Single-Drone has several near-optimal
critical points whose costs differ by 1e-5 .. 1e-4 relative and which all pass the reference's certificate
(S + eta I >= 0 with eta = 5e-6 f, src/CORA.cpp:154) -- the CPU restatement itself lands on 723.9580, 723.9665 or
724.0626 on Plaza2 at rank 4 depending on nothing but its thread count (summation order).  So the tight solves of
both sides start from a COMMON point already inside one basin: the CPU restatement's result with the reference's own
stopping rules.  From there two correct implementations must agree to 1e-6."""
import numpy as np
import pytest

from conftest import load_dataset, make_handle
from oracle import cora_oracle as co
from oracle import cpu_ref
from synth import make_synthetic

pytestmark = pytest.mark.gpu

REL = 1e-6   # north_star: "certified optimum matching reference to 1e-6 relative cost"


def _tight(max_iterations):
    from cora_b200 import capi
    return capi.default_tnt_params(max_iterations=max_iterations, max_computation_time=0.0,
                                   relative_decrease_tolerance=0.0, stepsize_tolerance=0.0,
                                   gradient_tolerance=1e-4, preconditioned_gradient_tolerance=0.0)


def _both(p, x0, max_iterations):
    from cora_b200 import capi
    prm = _tight(max_iterations)
    R = cpu_ref.CpuRef(p.d, p.n, p.m, p.n + p.l, p.Q, preconditioner=3, reg_lambda=p.lambda_reg)
    cpu = R.tnt(x0, prm)
    R.close()
    with make_handle(p, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
        h.reg_lambda = p.lambda_reg
        gpu = h.tnt(x0, prm)
    return cpu, gpu


def _problem(name, r):
    """The problem and the initial point: the datasets start from project(U[-1,1]) (seed 0);
    the 5k-pose synthetic chain from the perturbed ground truth of the bench workload (a random point on a
    5000-pose chain does not converge within the iteration limit on either side)."""
    if name == "synthetic_5k":
        from cora_b200 import synthetic
        n, l, m = 5000, 5, 1500
        p = make_synthetic(n, l, m, d=3, seed=31, rank=r, preconditioner=co.REG_CHOLESKY)
        p.update_problem_data()
        arrays, gt = synthetic.make_arrays(n, l, m, d=3, seed=31)
        x0 = p.project_to_manifold(synthetic.perturbed_ground_truth(3, n, l, arrays, gt, r, seed=0))
    else:
        p = load_dataset(name, preconditioner=co.REG_CHOLESKY)
        p.update_problem_data()
        p.rank = r
        x0 = p.project_to_manifold(np.random.default_rng(0).uniform(-1, 1, size=(p.N, r)))
    return p, x0


@pytest.mark.parametrize("name,r_lift", [("plaza2", 4), ("single_drone", 5), ("synthetic_5k", 5)])
def test_final_cost_matches_cpu_restatement_to_1e6(lib, name, r_lift):
    from cora_b200 import capi
    p, x0 = _problem(name, r_lift)
    d = p.d
    p.rank = r_lift
    R = cpu_ref.CpuRef(p.d, p.n, p.m, p.n + p.l, p.Q, preconditioner=3, reg_lambda=p.lambda_reg)
    loose = R.tnt(x0, capi.default_tnt_params(max_iterations=250, max_computation_time=0.0))  # src/CORA.cpp:95-109
    R.close()
    cpu, gpu = _both(p, loose.x, 400)
    assert abs(gpu.f - cpu.f) <= REL * abs(cpu.f), ("lifted", gpu.f, cpu.f, gpu.status, cpu.status)
    # rounding of the CPU solution (oracle), then both refine at rank d from that same point
    Yd = co.project_solution(p, cpu.x)
    p.rank = d
    cpu2, gpu2 = _both(p, Yd, 200)
    assert abs(gpu2.f - cpu2.f) <= REL * abs(cpu2.f), ("refined", gpu2.f, cpu2.f, gpu2.status, cpu2.status)
    # the rank-d cost can only be above the relaxation's (up to the convergence tolerance of the two solves: on the
    # synthetic chain the relaxation is tight and the two costs coincide to ~1e-9)
    assert gpu2.f >= gpu.f * (1 - REL)
