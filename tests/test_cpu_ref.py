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
