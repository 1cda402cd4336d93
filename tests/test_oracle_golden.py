"""The oracle against every golden vector the reference's own tests hold for the path
(SURVEY.md 8c).  CPU only."""
import math

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import FIXTURES, load_fixture
from oracle import cora_oracle as co

SUB = ["Apose", "Arange", "OmegaPose", "OmegaRange", "RangeDistances", "T", "RotConLaplacian"]


@pytest.mark.parametrize("name", FIXTURES)
def test_submatrices_and_data_matrix(name):  # tests/test_utils.cpp:110-178
    g, p = load_fixture(name)
    sub = p.submatrices()
    for k in SUB:
        ref = g[k]
        got = sub[k].toarray() if sub[k].shape[0] else np.zeros(ref.shape)
        if ref.size == 0 and got.size == 0:
            continue
        assert got.shape == ref.shape, k
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12, err_msg=k)
    p.preconditioner = co.JACOBI
    p.update_problem_data()
    np.testing.assert_allclose(p.Q.toarray(), g["DataMatrix"], rtol=1e-12, atol=1e-10)
    # integer indexing bit-exact: the sorted (row, col) pattern with exact zeros dropped
    ref = sp.csr_matrix(g["DataMatrix"]); ref.eliminate_zeros(); ref.sort_indices()
    assert np.array_equal(ref.indptr, p.Q.indptr) and np.array_equal(ref.indices, p.Q.indices)


@pytest.mark.parametrize("name", FIXTURES)
def test_operators(name):  # tests/test_optimizer_helpers.cpp:13-38
    g, p = load_fixture(name)
    p.preconditioner = co.JACOBI
    p.rank = 2
    p.update_problem_data()
    X, dX = g["X_rand_dim2"], g["rand_dX"]
    assert abs(p.evaluate_objective(X) - float(g["expected_cost"])) < 1e-9
    eg = p.euclidean_gradient(X)
    np.testing.assert_allclose(eg, g["expected_egrad"], atol=1e-9)
    np.testing.assert_allclose(p.riemannian_gradient(X, eg), g["expected_rgrad"], atol=1e-9)
    np.testing.assert_allclose(p.hessvec(X, eg, dX), g["hessProd"], atol=1e-9)


@pytest.mark.parametrize("name", FIXTURES)
def test_ground_truth_in_nullspace(name):  # tests/test_construct_problem.cpp:45-76,110-125
    g, p = load_fixture(name)
    p.preconditioner = co.JACOBI
    p.update_problem_data()
    Xgt = g["X_gt"]
    assert np.linalg.norm(p.Q @ Xgt) < 1e-6  # fixture text carries ~1e-9 digits
    assert abs(p.evaluate_objective(Xgt)) < 1e-9 * max(1.0, np.abs(p.Q).sum())
    rng = np.random.default_rng(0)
    R, _ = np.linalg.qr(rng.standard_normal((Xgt.shape[1], Xgt.shape[1])))
    assert np.linalg.norm(p.Q @ (Xgt @ R)) < 1e-6


@pytest.mark.parametrize("name", FIXTURES)
def test_certificate_matrix(name):  # tests/test_certification.cpp:81-125
    g, p = load_fixture(name)
    p.preconditioner = co.JACOBI
    p.update_problem_data()
    np.testing.assert_allclose(p.certificate_matrix(g["X_rand_dim2"]).toarray(), g["S_rand"], atol=1e-9)
    st, ob = p.compute_lambda_blocks(g["X_gt"])
    assert np.abs(st).max(initial=0) < 1e-6 and np.abs(ob).max(initial=0) < 1e-6
    S = p.certificate_matrix(g["X_gt"]).toarray()
    assert np.linalg.eigvalsh(S)[0] > -1e-6


@pytest.mark.parametrize("n", [10, 1000])
def test_fast_verification_known_answers(n):  # tests/test_certification.cpp:45-79
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n); x /= np.linalg.norm(x)
    I = sp.identity(n, format="csr")
    res = co.fast_verification(I, 1e-8, min(4, n))
    assert res.is_certified
    res = co.fast_verification(sp.csr_matrix(I - np.outer(x, x)), 1e-8, min(4, n))
    assert res.is_certified
    res = co.fast_verification(sp.csr_matrix(I - 2 * np.outer(x, x)), 1e-8, min(4, n))
    assert not res.is_certified
    assert abs(res.theta + 1.0) < 1e-6
    v = res.x / np.linalg.norm(res.x)
    assert min(np.linalg.norm(v - x), np.linalg.norm(v + x)) < 1e-5


def test_stpcg_known_answers():  # libs/Optimization/tests/IterativeSolvers_unit_test.cpp:138-310
    A = np.diag([1000.0, 100.0, 1.0])
    g = np.array([21.0, -0.4, 19.0])
    inner = lambda a, b: float(a @ b)
    H = lambda v: A @ v
    s, nrm, it = co.stpcg(g, H, inner, 1e6, 1000, 1e-12, 1.0)
    np.testing.assert_allclose(s, -np.linalg.solve(A, g), rtol=1e-8)
    assert it <= 3 and abs(nrm - np.linalg.norm(s)) < 1e-8
    # negative curvature -> step to the boundary
    An = np.diag([1000.0, 100.0, -1.0])
    s, nrm, it = co.stpcg(g, lambda v: An @ v, inner, 5.0, 1000, 1e-12, 1.0)
    assert abs(np.linalg.norm(s) - 5.0) < 1e-9 and nrm == 5.0
    # preconditioned: M = diag(A) -> one iteration, M-norm of the step
    P = lambda v: v / np.diag(A)
    s, nrm, it = co.stpcg(g, H, inner, 1e6, 1000, 1e-12, 1.0, P)
    np.testing.assert_allclose(s, -np.linalg.solve(A, g), rtol=1e-8)
    assert it == 1 and abs(nrm - math.sqrt(float(s @ (A @ s)))) < 1e-8
    # truncated by the trust region
    s, nrm, it = co.stpcg(g, H, inner, 1e-3, 1000, 1e-12, 1.0)
    assert abs(np.linalg.norm(s) - 1e-3) < 1e-12


def test_tnt_sphere_known_answer():  # libs/Optimization/tests/TNT_unit_test.cpp:126-187
    Pn = np.array([0.0, 0.0, 1.0])
    f = lambda x: float(np.sum((x - Pn) ** 2))
    proj = lambda x, v: v - x * float(x @ v)

    def QM(x):
        eg = 2 * (x - Pn)
        grad = proj(x, eg)
        return grad, (lambda v: proj(x, 2 * v) - float(x @ eg) * v)

    retract = lambda x, v: (x + v) / np.linalg.norm(x + v)
    metric = lambda a, b: float(a @ b)
    x0 = np.array([-0.5, -0.5, -0.707107])
    prm = co.TNTParams(relative_decrease_tolerance=0, stepsize_tolerance=0,
                       preconditioned_gradient_tolerance=0, gradient_tolerance=1e-6)
    for precon in (None, lambda x, v: np.array([1.0, 2.0, 3.0]) * v):
        res = co.tnt(f, QM, metric, retract, x0, precon, prm)
        assert res.status == "Gradient"
        assert np.linalg.norm(QM(res.x)[0]) < 1e-6 and res.f < f(x0)
        assert np.linalg.norm(res.x - Pn) < 1e-5


def test_small_problem_staircase_known_answer():
    g, p = load_fixture("small_ra_slam_problem")
    p.preconditioner = co.JACOBI
    p.rank = 2
    p.update_problem_data()
    x0 = p.random_initial_guess(np.random.default_rng(0))
    res = co.solve_cora(p, x0, max_rank=6)
    assert res.certified or res.stages[-2]["certified"]
    assert abs(res.lifted_f) < 1e-8
