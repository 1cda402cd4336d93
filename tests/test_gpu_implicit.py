"""Formulation::Implicit (translations marginalised; src/CORA_problem.cpp:714-757, :878-885, :1085-1100,
:1168-1197; examples/config.json:8) through the C-ABI against the oracle's restatement, on the three reference
fixtures, Plaza2 (chain factor of the translation Laplacian) and a loop-closure graph (general factor)."""
import numpy as np
import pytest

from conftest import FIXTURES, load_dataset, load_fixture, make_handle
from oracle import cora_oracle as co
from synth import make_synthetic

pytestmark = pytest.mark.gpu


def _params(**kw):
    from cora_b200 import capi
    base = dict(max_computation_time=0.0)
    base.update(kw)
    return capi.default_tnt_params(**base)


def _problems():
    for name in FIXTURES:
        g, p = load_fixture(name)
        if p.n + p.l >= 2:
            yield name, p
    yield "plaza2", load_dataset("plaza2")
    yield "single_drone", load_dataset("single_drone")
    yield "synthetic loops", make_synthetic(n=150, l=3, m=80, d=3, seed=2, loop_closures=[(3, 90), (20, 140)])
    yield "synthetic d2 no landmarks", make_synthetic(n=200, l=0, m=0, d=2, seed=3)


def _close(a, b, tol):
    assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300), np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name,p", list(_problems()), ids=[c[0] for c in _problems()])
def test_implicit_operators(lib, name, p):
    from cora_b200 import capi
    p.preconditioner = co.REG_CHOLESKY
    p.update_problem_data()
    p.set_formulation(co.IMPLICIT)
    k = p.rot_and_range_size
    r = p.d + 2
    rng = np.random.default_rng(0)
    Y = p.project_to_manifold(rng.standard_normal((k, r)))
    V = rng.standard_normal((k, r))
    with make_handle(p, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
        h.reg_lambda = p.lambda_reg
        h.set_formulation(capi.FORMULATION_IMPLICIT)
        assert h.rows == k
        with pytest.raises(capi.InvalidArgument):
            h.data_matrix_product(np.zeros((p.N, r)))  # the explicit shape is rejected (src/CORA.cpp:38)
        _close(h.translation_explicit_solution(Y), p.translation_explicit_solution(Y), 1e-9)
        ref = p.data_matrix_product(Y)
        _close(h.data_matrix_product(Y), ref, 1e-9)
        f = p.evaluate_objective(Y)
        assert abs(h.evaluate_objective(Y) - f) <= 1e-9 * abs(f)
        eg = p.euclidean_gradient(Y)
        _close(h.riemannian_gradient(Y), p.riemannian_gradient(Y, eg), 1e-9)
        _close(h.hessvec(Y, eg, V), p.hessvec(Y, eg, V), 1e-9)
        _close(h.hessvec(Y, None, V), p.hessvec(Y, eg, V), 1e-9)
        _close(h.precondition(V), p.precondition(V), 1e-8)
        _close(h.retract(Y, 0.1 * V), p.retract(Y, 0.1 * V), 1e-9)
        _close(h.tangent_space_projection(Y, V), p.tangent_space_projection(Y, V), 1e-12)
        # back to the explicit formulation: the handle is the same problem
        h.set_formulation(capi.FORMULATION_EXPLICIT)
        X = p.translation_explicit_solution(Y)
        p.set_formulation(co.EXPLICIT)
        _close(h.data_matrix_product(X), p.data_matrix_product(X), 1e-9)


@pytest.mark.parametrize("name", ["plaza2", "synthetic loops"])
def test_implicit_tnt_and_certificate(lib, name):
    """TNT in the implicit formulation: leading iterations agree with the oracle; a failed certificate returns the
    normalised rotation/range part of the direction with its implicit Rayleigh quotient (:1085-1100)."""
    from cora_b200 import capi
    if name == "plaza2":
        p = load_dataset("plaza2", preconditioner=co.REG_CHOLESKY)
    else:
        p = make_synthetic(n=150, l=3, m=80, d=3, seed=2, preconditioner=co.REG_CHOLESKY, loop_closures=[(3, 90), (20, 140)])
    p.update_problem_data()
    p.set_formulation(co.IMPLICIT)
    p.rank = p.d + 1
    k = p.rot_and_range_size
    x0 = p.random_initial_guess(np.random.default_rng(0))
    ref = co.problem_tnt(p, x0, co.cora_tnt_params(max_iterations=5))
    with make_handle(p, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
        h.reg_lambda = p.lambda_reg
        h.set_formulation(capi.FORMULATION_IMPLICIT)
        got = h.tnt(x0, _params(max_iterations=5))
        assert got.x.shape == (k, p.rank)
        assert got.inner_iterations[:3] == ref.inner_iterations[:3]
        np.testing.assert_allclose(got.objective_values[:4], ref.objective_values[:4], rtol=1e-6)
        np.testing.assert_allclose(got.preconditioned_gradient_norms[:3], ref.preconditioned_gradient_norms[:3], rtol=1e-6)
        # certificate at the (non-optimal) iterate
        Y = got.x
        eta = 1e-5
        c = h.certify_solution(Y, eta, 10)
        assert not c.is_certified
        assert c.x.shape == (k,) and abs(np.linalg.norm(c.x) - 1) <= 1e-12
        Lam = p.lambda_from_blocks(p.compute_lambda_blocks(Y), k)
        th = c.x @ (p.data_matrix_product(c.x[:, None])[:, 0] - Lam @ c.x)
        assert abs(th - c.theta) <= 1e-8 * max(abs(th), 1.0)
        assert c.theta < -eta / 2
        Yp = h.saddle_escape(Y, c.theta, c.x)
        assert Yp.shape == (k, p.rank + 1)
        p.rank += 1
        assert p.evaluate_objective(Yp) < p.evaluate_objective(np.hstack([Y, np.zeros((k, 1))]))


def test_implicit_staircase_reaches_the_explicit_optimum(lib):
    """solveCORA in the implicit formulation (src/CORA.cpp:30-39,161-164) ends at the cost of the explicit solve:
    the two formulations share their optimum."""
    from cora_b200 import capi
    p = load_dataset("plaza2", preconditioner=co.REG_CHOLESKY)
    p.update_problem_data()
    k = p.d * p.n + p.m
    x0 = np.random.default_rng(0).uniform(-1, 1, size=(p.N, 3))
    with make_handle(p, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
        ex = h.solve(x0, max_rank=8, params=_params())
        h.set_formulation(capi.FORMULATION_IMPLICIT)
        im = h.solve(x0[:k], max_rank=8, params=_params())  # examples/paper_experiments.cpp:623-625
        assert im["x"].shape == (k, p.d)
        assert im["certified"], im["stages"]
        assert abs(im["f"] - ex["f"]) <= 1e-4 * ex["f"], (im["f"], ex["f"])
        assert abs(im["f"] - 734.328) <= 1e-4 * 734.328  # run_utils/parse_data.py:40
        Xf = h.translation_explicit_solution(im["x"])
        h.set_formulation(capi.FORMULATION_EXPLICIT)
        assert abs(h.evaluate_objective(Xf) - im["f"]) <= 1e-9 * im["f"]
