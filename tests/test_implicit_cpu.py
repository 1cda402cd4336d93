"""Oracle restatement of Formulation::Implicit (src/CORA_problem.cpp:714-757, :878-885, :1085-1100, :1168-1197).

The reference has no test or golden vector for the implicit formulation, so the restatement is pinned through the
identities that tie it to the explicit operators (which ARE pinned by the reference's goldens,
tests/test_oracle_golden.py): with X = getTranslationExplicitSolution(Y) = [Y; t*(Y)],
  * Q_implicit Y = rows of Q X, and the translation rows of Q X vanish (t* minimises over the translations),
  * f_implicit(Y) = f_explicit(X) <= f_explicit([Y; t]) for any other translations t,
  * the ground truth (rotations, ranges) is in the kernel of Q_implicit (tests/test_construct_problem.cpp:45-76)."""
import numpy as np
import pytest

from conftest import FIXTURES, load_dataset, load_fixture
from oracle import cora_oracle as co
from synth import make_synthetic


def _problems():
    for name in FIXTURES:
        g, p = load_fixture(name)
        if p.n + p.l >= 2:
            yield name, p, np.asarray(g["X_gt"])
    yield "plaza2", load_dataset("plaza2"), None
    yield "synthetic loops", make_synthetic(n=150, l=3, m=80, d=3, seed=2, loop_closures=[(3, 90), (20, 140)]), None


@pytest.mark.parametrize("name,p,Xgt", list(_problems()), ids=[c[0] for c in _problems()])
def test_implicit_operators_against_explicit(name, p, Xgt):
    p.preconditioner = co.REG_CHOLESKY
    p.update_problem_data()
    q = co.Problem.__new__(co.Problem)
    q.__dict__.update(p.__dict__)
    q.set_formulation(co.IMPLICIT)
    k = q.rot_and_range_size
    assert q.expected_variable_size == k == p.d * p.n + p.m
    r = p.d + 2
    rng = np.random.default_rng(0)
    Y = co.project_to_manifold(p.d, p.n, p.m, rng.standard_normal((k, r)))
    X = q.translation_explicit_solution(Y)
    assert X.shape == (p.N, r) and np.array_equal(X[:k], Y) and not X[-1].any()
    QX = p.data_matrix_product(X)
    scale = np.abs(QX[:k]).max() + 1e-300
    np.testing.assert_allclose(q.data_matrix_product(Y), QX[:k], atol=1e-9 * scale)
    assert np.abs(QX[k:]).max() <= 1e-8 * scale  # stationarity in the translations (the pinned last row too)
    f = q.evaluate_objective(Y)
    assert abs(f - p.evaluate_objective(X)) <= 1e-9 * abs(f)
    Xp = X.copy()
    Xp[k:] += 1e-3 * rng.standard_normal(Xp[k:].shape)
    assert p.evaluate_objective(Xp) >= f
    # gradient / Hessian-vector product: the explicit ones at X restricted to the top rows, for directions whose
    # translation part is the completion of the direction
    V = rng.standard_normal((k, r))
    Vf = q.translation_explicit_solution(V)
    np.testing.assert_allclose(q.riemannian_gradient(Y), p.riemannian_gradient(X)[:k], atol=1e-8 * scale)
    H = q.hessvec(Y, q.euclidean_gradient(Y), V)
    Hf = p.hessvec(X, p.euclidean_gradient(X), Vf)
    np.testing.assert_allclose(H, Hf[:k], atol=1e-8 * max(np.abs(Hf).max(), 1e-300))
    # preconditioner: lift with zero translations, solve, keep the top rows (:878-885)
    lift = np.zeros((p.N, r)); lift[:k] = V
    np.testing.assert_allclose(q.precondition(V), p.precondition(lift)[:k], rtol=0, atol=1e-12 * np.abs(V).max() / p.lambda_reg)
    if Xgt is not None:
        G = q.data_matrix_product(np.asarray(Xgt)[:k])
        assert np.abs(G).max() <= 1e-8 * max(1.0, abs(p.Q).max())


def test_implicit_certificate_truncation():
    p = make_synthetic(n=60, l=2, m=30, d=3, seed=4, rank=4)
    p.update_problem_data()
    p.set_formulation(co.IMPLICIT)
    Y = p.random_initial_guess(np.random.default_rng(1))
    assert Y.shape[0] == p.rot_and_range_size
    eta = 1e-3
    res = p.certify_solution(Y, eta, 10, p.translation_explicit_solution(Y))
    assert not res.is_certified
    assert res.x.shape == (p.rot_and_range_size,) and abs(np.linalg.norm(res.x) - 1) < 1e-12
    Lam = p.lambda_from_blocks(p.compute_lambda_blocks(Y), p.rot_and_range_size)
    th = res.x @ (p.data_matrix_product(res.x[:, None])[:, 0] - Lam @ res.x)
    assert abs(th - res.theta) <= 1e-12 * abs(th)
    assert res.theta < 0  # a direction of negative curvature of the implicit problem
