"""RegularizedCholesky and the Cholesky PSD certificate on graphs that are NOT one odometry chain -- loop closures,
several robots (TIERS, MR.CLAM) -- through the C-ABI: the general sparse block Cholesky of
cora_b200/csrc/gen_chol.hpp with its level-scheduled device solves (gen_chol_dev.cuh) against the oracle's sparse LU
of the same matrix (src/CORA_problem.cpp:544-614, src/CORA_preconditioners.cpp:16-83, src/CORA_utils.cpp:33-57)."""
import numpy as np
import pytest

from conftest import load_dataset, make_handle
from oracle import cora_oracle as co
from synth import make_synthetic
from test_general_cholesky_cpu import _loops, _two_robots

pytestmark = pytest.mark.gpu


def _params(**kw):
    from cora_b200 import capi
    base = dict(max_computation_time=0.0)
    base.update(kw)
    return capi.default_tnt_params(**base)


def _cases():
    yield "one loop closure", make_synthetic(n=60, l=3, m=40, d=3, seed=5, loop_closures=[(0, 30)])
    yield "d2 random loops", make_synthetic(n=400, l=2, m=150, d=2, seed=3, loop_closures=_loops(400, 60, 1))
    yield "d3 random loops", make_synthetic(n=700, l=4, m=300, d=3, seed=11, loop_closures=_loops(700, 70, 2))
    yield "d3 near loops, no landmarks", make_synthetic(n=900, l=0, m=0, d=3, seed=4, loop_closures=_loops(900, 200, 3, near=12))
    yield "d3 hub pose (> 8 couplings)", make_synthetic(n=300, l=2, m=100, d=3, seed=6,
                                                        loop_closures=[(7, j) for j in range(20, 300, 9)])
    yield "two robots (two components)", _two_robots()
    yield "three poses, one closure", make_synthetic(n=3, l=1, m=2, d=2, seed=1, loop_closures=[(0, 2)])
    yield "dense little graph", make_synthetic(n=9, l=0, m=0, d=3, seed=2,
                                               loop_closures=[(i, j) for i in range(9) for j in range(i + 2, 9)])
    yield "tiers", load_dataset("tiers")
    yield "mrclam2", load_dataset("mrclam2")


@pytest.mark.parametrize("name,p", list(_cases()), ids=[c[0] for c in _cases()])
def test_general_regularized_cholesky_preconditioner(lib, name, p):
    """Problem::precondition with Preconditioner::RegularizedCholesky (the reference default) == splu of the same matrix."""
    from cora_b200 import capi
    p.preconditioner = co.REG_CHOLESKY
    p.update_problem_data()
    rng = np.random.default_rng(0)
    with make_handle(p, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
        assert h.effective_preconditioner == capi.PRECON_REG_CHOLESKY
        assert abs(h.reg_lambda - p.lambda_reg) <= 1e-2 * p.lambda_reg
        h.reg_lambda = p.lambda_reg
        for r in (1, p.d, 5, 12):  # 12 columns: two passes of the warp's column window
            V = rng.standard_normal((p.N, r))
            Z = h.precondition(V)
            ref = p.precondition(V)
            assert np.abs(Z - ref).max() <= 1e-8 * np.abs(ref).max(), (name, r)
            assert np.all(Z[-1] == 0.0)  # CORA_preconditioners.cpp:77-80
            assert np.array_equal(Z, h.precondition(V))  # fixed summation order: bit-reproducible


def test_general_psd_test_matches_dense_eigenvalues(lib):
    p = make_synthetic(n=120, l=3, m=60, d=3, seed=8, rank=4, loop_closures=_loops(120, 25, 7))
    p.update_problem_data()
    Y = p.project_to_manifold(np.random.default_rng(0).standard_normal((p.N, 4)))
    lam_min = float(np.linalg.eigvalsh(p.certificate_matrix(Y).toarray())[0])
    assert lam_min < 0
    with make_handle(p) as h:
        assert h.psd_test(-lam_min * 1.05, Y) is True
        assert h.psd_test(-lam_min * 0.95, Y) is False
        assert h.psd_test(1e-6, Y) is False


@pytest.mark.parametrize("name,r", [("synthetic", 5), ("tiers", 3)])
def test_general_regularized_cholesky_tnt(lib, name, r):
    """TNT with the general factor as preconditioner (multi-launch path): leading iterations agree with the oracle."""
    from cora_b200 import capi
    if name == "synthetic":
        p = make_synthetic(n=700, l=4, m=300, d=3, seed=11, preconditioner=co.REG_CHOLESKY, loop_closures=_loops(700, 70, 2))
    else:
        p = load_dataset(name, preconditioner=co.REG_CHOLESKY)
    p.update_problem_data()
    p.rank = r
    x0 = p.random_initial_guess(np.random.default_rng(0))
    ref = co.problem_tnt(p, x0, co.cora_tnt_params(max_iterations=5))
    with make_handle(p, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
        h.reg_lambda = p.lambda_reg
        got = h.tnt(x0, _params(max_iterations=5))
    assert got.inner_iterations[:3] == ref.inner_iterations[:3]
    np.testing.assert_allclose(got.objective_values[:4], ref.objective_values[:4], rtol=1e-6)
    np.testing.assert_allclose(got.preconditioned_gradient_norms[:3], ref.preconditioned_gradient_norms[:3], rtol=1e-6)


def test_staircase_with_loop_closures_certifies_by_cholesky(lib):
    """The staircase on a loop-closure graph: every rank lift follows a verified direction of negative curvature, no
    verdict is 'inconclusive' (S + eta I is factored by the general Cholesky), and the certified solution has the
    oracle's cost.  (Which of the reference's two certificates ends the staircase -- Cholesky of S + eta I or the
    singular-value ratio of a rank-deficient Y, src/CORA_problem.cpp:1039-1049 -- depends on the escape direction.)"""
    from cora_b200 import capi
    p = make_synthetic(n=80, l=3, m=50, d=3, seed=3, preconditioner=co.REG_CHOLESKY, loop_closures=[(0, 40), (10, 70)])
    p.update_problem_data()
    x0 = np.random.default_rng(1).uniform(-1, 1, size=(p.N, 4))
    with make_handle(p, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
        assert h.effective_preconditioner == capi.PRECON_REG_CHOLESKY
        out = h.solve(x0, max_rank=8, params=_params())
    st = out["stages"]
    assert out["certified"], st
    lifted = [s for s in st if s["certified"]][0]
    assert lifted["cert_branch"] in ("psd", "sv_ratio"), st
    assert all(s["cert_branch"] == "eigenpair" and s["theta"] < -s["eta"] / 2 for s in st[: st.index(lifted)]), st
    assert all(s["cert_branch"] != "inconclusive" for s in st), st
    p.rank = 4
    ref = co.solve_cora(p, p.project_to_manifold(x0), max_rank=8)
    assert abs(out["f"] - ref.result.f) <= 1e-4 * max(abs(ref.result.f), 1.0), (out["f"], ref.result.f)


def test_tiers_staircase_certifies(lib):
    """TIERS (4 robots, inter-robot ranges): the drop-in call with the reference's default preconditioner builds
    the general factor, converges and certifies."""
    from cora_b200 import capi
    p = load_dataset("tiers", preconditioner=co.REG_CHOLESKY)
    p.update_problem_data()
    x0 = np.random.default_rng(0).uniform(-1, 1, size=(p.N, 4))
    with make_handle(p, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
        out = h.solve(x0, max_rank=10, params=_params())
    st = out["stages"]
    assert np.isfinite(out["f"]) and out["x"].shape == (p.N, p.d)
    assert all(s["cert_branch"] in ("psd", "eigenpair", "sv_ratio") for s in st), st
    assert out["certified"], st
