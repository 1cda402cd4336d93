"""The C++ CPU restatement (oracle/cpu_ref.cpp, the timed baseline of bench.py) against the NumPy/SciPy
oracle that the reference's golden fixtures pin (tests/test_oracle_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import FIXTURES, load_dataset, load_fixture
from oracle import cora_oracle as co
from oracle import cpu_ref
from synth import make_synthetic


def _ref(p, **kw):
    return cpu_ref.CpuRef(p.d, p.n, p.m, p.n + p.l, p.Q, **kw)


@pytest.mark.parametrize("name", FIXTURES)
def test_operators_on_reference_fixtures(name):
    g, p = load_fixture(name)
    p.preconditioner = co.JACOBI
    p.update_problem_data()
    X = np.asarray(g["X_rand_dim2"], dtype=float)
    dX = np.asarray(g["rand_dX"], dtype=float)
    R = _ref(p)
    # the same goldens the reference asserts in tests/test_optimizer_helpers.cpp:21-37
    eg = R.data_matrix_product(X)
    assert np.allclose(eg, g["expected_egrad"], atol=1e-9)
    assert np.allclose(R.tangent_space_projection(X, eg), g["expected_rgrad"], atol=1e-9)
    assert np.allclose(R.hessvec(X, eg, dX), g["hessProd"], atol=1e-9)


@pytest.mark.parametrize("d,r", [(2, 3), (3, 5), (3, 7)])
def test_operators_against_numpy_oracle(d, r):
    p = make_synthetic(300, 3, 120, d=d, seed=5, rank=r)
    p.update_problem_data()
    rng = np.random.default_rng(1)
    A = rng.standard_normal((p.N, r))
    V = rng.standard_normal((p.N, r))
    R = _ref(p)
    Y = R.project_to_manifold(A)
    assert np.abs(Y - p.project_to_manifold(A)).max() < 1e-12
    eg = p.euclidean_gradient(Y)
    assert np.abs(R.data_matrix_product(Y) - eg).max() <= 1e-12 * np.abs(eg).max()
    hv = p.hessvec(Y, eg, V)
    assert np.abs(R.hessvec(Y, eg, V) - hv).max() <= 1e-12 * np.abs(hv).max()


@pytest.mark.parametrize("threads", [1, 3])
def test_tnt_trajectory_matches_numpy_oracle(threads):
    from cora_b200 import capi, synthetic
    d, n, l, m, r = 3, 400, 3, 150, 5
    p = make_synthetic(n, l, m, d=d, seed=11, rank=r)
    p.update_problem_data()
    arrays, _ = synthetic.make_arrays(n, l, m, d=d, seed=11)
    x0 = p.project_to_manifold(synthetic.odometry_initialization(d, n, l, arrays, r, seed=0))
    R = _ref(p, threads=threads)
    assert R.threads == threads
    got = R.tnt(x0, capi.default_tnt_params(max_iterations=16, max_computation_time=0.0))
    ref = co.problem_tnt(p, x0, co.cora_tnt_params(max_iterations=16))
    assert got.status == ref.status
    assert got.inner_iterations == ref.inner_iterations
    assert np.allclose(got.objective_values, ref.objective_values, rtol=1e-9)
    assert np.allclose(got.trust_region_radius, ref.trust_region_radius, rtol=1e-9)
    # the reference's operation count: per outer iteration f, QM and the model decrease each cost a
    # data-matrix product on top of one per CG iteration (TNT.h:508,511-512,573)
    assert R.spmm_count() >= sum(got.inner_iterations) + 2 * len(got.inner_iterations)


@pytest.mark.parametrize("case", ["synth2", "synth3", "plaza2", "single_drone", "no_landmarks"])
def test_regularized_cholesky_against_oracle_sparse_lu(case):
    """RegularizedCholesky of the C++ port (ranges, block-tridiagonal chain, landmark border) against the oracle's
    sparse LU of the same matrix (Q + lambda I)[0:N-1, 0:N-1] (src/CORA_problem.cpp:544-614), last row zero
    (src/CORA_preconditioners.cpp:77-80) -- the equality tests/test.cpp:25-214 asserts for CHOLMOD."""
    if case == "synth2":
        p = make_synthetic(300, 3, 120, d=2, seed=5, rank=4, preconditioner=co.REG_CHOLESKY)
    elif case == "synth3":
        p = make_synthetic(300, 4, 160, d=3, seed=6, rank=5, preconditioner=co.REG_CHOLESKY)
    elif case == "no_landmarks":
        p = make_synthetic(120, 0, 0, d=3, seed=7, rank=4, preconditioner=co.REG_CHOLESKY)
    else:
        p = load_dataset(case, rank=4, preconditioner=co.REG_CHOLESKY)
    p.update_problem_data()
    V = np.random.default_rng(2).standard_normal((p.N, 3))
    R = _ref(p, preconditioner=3, reg_lambda=p.lambda_reg)
    Z = R.precondition(V)
    ref = p.precondition(V)
    assert np.abs(Z - ref).max() <= 1e-8 * np.abs(ref).max()
    assert np.all(Z[-1] == 0.0)


def test_chain_posdef_matches_dense_eigenvalues():
    """The Cholesky PSD test of the port (PSD half of fast_verification, src/CORA_utils.cpp:33-57) against dense
    eigenvalues of the certificate matrix at a random point and at the ground truth."""
    p = make_synthetic(60, 3, 40, d=3, seed=8, rank=4)
    p.update_problem_data()
    Y = p.project_to_manifold(np.random.default_rng(0).standard_normal((p.N, 4)))
    S = p.certificate_matrix(Y)
    lam_min = float(np.linalg.eigvalsh(S.toarray())[0])
    assert lam_min < 0
    for shift, want in ((-lam_min * 0.9, False), (-lam_min * 1.1, True), (0.0, False)):
        assert cpu_ref.chain_posdef(p.d, p.n, p.m, p.n + p.l, S, shift) == want


def test_tnt_with_regularized_cholesky_matches_numpy_oracle():
    from cora_b200 import capi, synthetic
    d, n, l, m, r = 3, 300, 3, 120, 5
    p = make_synthetic(n, l, m, d=d, seed=13, rank=r, preconditioner=co.REG_CHOLESKY)
    p.update_problem_data()
    arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=13)
    x0 = p.project_to_manifold(synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=1))
    R = _ref(p, preconditioner=3, reg_lambda=p.lambda_reg)
    got = R.tnt(x0, capi.default_tnt_params(max_iterations=12, max_computation_time=0.0))
    ref = co.problem_tnt(p, x0, co.cora_tnt_params(max_iterations=12))
    k = min(len(ref.inner_iterations), len(got.inner_iterations), 6)
    assert got.inner_iterations[:k] == ref.inner_iterations[:k]
    assert np.allclose(got.objective_values[:k], ref.objective_values[:k], rtol=1e-8)
