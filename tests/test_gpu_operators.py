"""Tier-1 parity: every operator of the hot path through the C-ABI against the golden
fixtures of the reference's own tests and against the oracle (fp64, tolerance stated per test).

Tolerances: the reference asserts 1e-6 absolute on these operators
(tests/test_optimizer_helpers.cpp:13-38); we hold the CUDA path to 1e-9 relative to the
magnitude of the data (summation order differs from Eigen's, nothing else does)."""
import numpy as np
import pytest

from conftest import FIXTURES, load_dataset, load_fixture, make_handle
from oracle import cora_oracle as co
from synth import make_synthetic

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _close(got, ref, rtol=RTOL):
    scale = max(1.0, float(np.abs(ref).max(initial=0.0)))
    err = float(np.abs(np.asarray(got) - np.asarray(ref)).max(initial=0.0))
    assert err <= rtol * scale, "max abs err %.3e (scale %.3e)" % (err, scale)


@pytest.mark.parametrize("name", FIXTURES)
def test_golden_operators(lib, name):
    g, p = load_fixture(name)
    p.preconditioner = co.JACOBI
    p.rank = 2
    p.update_problem_data()
    X, dX = g["X_rand_dim2"], g["rand_dX"]
    with make_handle(p) as h:
        assert abs(h.evaluate_objective(X) - float(g["expected_cost"])) < 1e-9 * max(1, abs(float(g["expected_cost"])))
        eg = h.euclidean_gradient(X)
        _close(eg, g["expected_egrad"])
        _close(h.riemannian_gradient(X, eg), g["expected_rgrad"])
        _close(h.riemannian_gradient(X), g["expected_rgrad"])
        _close(h.hessvec(X, eg, dX), g["hessProd"])
        _close(h.hessvec(X, None, dX), g["hessProd"])
        # certificate matrix S(X_rand) applied to the identity == S_rand.mm
        S = h.certificate_product(X, np.eye(p.N))
        _close(S, g["S_rand"])
        st, ob = h.compute_lambda_blocks(g["X_gt"])
        assert np.abs(st).max(initial=0) < 1e-6 and np.abs(ob).max(initial=0) < 1e-6
        # ground truth in the null space (tests/test_construct_problem.cpp:45-76)
        assert np.abs(h.data_matrix_product(g["X_gt"])).max() < 1e-6


def _operator_suite(p, r, seed=0, rtol=RTOL):
    rng = np.random.default_rng(seed)
    p.rank = r
    N = p.N
    Y = p.project_to_manifold(rng.uniform(-1, 1, size=(N, r)))
    V = rng.standard_normal((N, r))
    with make_handle(p) as h:
        _close(h.data_matrix_product(V), p.data_matrix_product(V), rtol)
        f_ref = p.evaluate_objective(Y)
        assert abs(h.evaluate_objective(Y) - f_ref) <= rtol * max(1.0, abs(f_ref))
        eg = p.euclidean_gradient(Y)
        _close(h.euclidean_gradient(Y), eg, rtol)
        _close(h.riemannian_gradient(Y, eg), p.riemannian_gradient(Y, eg), rtol)
        _close(h.tangent_space_projection(Y, V), p.tangent_space_projection(Y, V), rtol)
        _close(h.hessvec(Y, eg, V), p.hessvec(Y, eg, V), rtol)
        _close(h.precondition(V), p.precondition(V), rtol)
        # retraction: polar factor / normalisation (PARITY UNPINNED in the reference; the
        # oracle's SVD-based polar factor defines it).  Small steps and a generic block.
        _close(h.retract(Y, 0.1 * V), p.retract(Y, 0.1 * V), 1e-9)
        A = rng.uniform(-1, 1, size=(N, r))
        Pg, Pr = h.project_to_manifold(A), p.project_to_manifold(A)
        _close(Pg, Pr, 1e-7)
        d, n, m = p.d, p.n, p.m
        B = Pg[: d * n].reshape(n, d, r)
        assert np.abs(np.einsum("nir,njr->nij", B, B) - np.eye(d)).max(initial=0) < 1e-13
        if m:
            assert np.abs(np.linalg.norm(Pg[d * n: d * n + m], axis=1) - 1).max() < 1e-13
        st, ob = h.compute_lambda_blocks(Y)
        st_ref, ob_ref = p.compute_lambda_blocks(Y)
        _close(st.T.reshape(n, d, d), st_ref, rtol)
        _close(ob, ob_ref, rtol)
        X = rng.standard_normal((N, 3))
        _close(h.certificate_product(Y, X), p.certificate_matrix(Y) @ X, rtol)


@pytest.mark.parametrize("name,r", [("plaza2", 3), ("plaza2", 4), ("single_drone", 5), ("single_drone", 3)])
def test_datasets(lib, name, r):
    p = load_dataset(name, preconditioner=co.JACOBI)
    p.update_problem_data()
    _operator_suite(p, r)


@pytest.mark.parametrize("d,r", [(2, 2), (2, 5), (3, 3), (3, 5), (3, 7), (3, 12)])
def test_synthetic_ranks(lib, d, r):
    p = make_synthetic(n=700, l=4, m=300, d=d, seed=11)
    p.update_problem_data()
    _operator_suite(p, r, seed=r)


def test_edge_cases(lib):
    # no ranges / no landmarks; single relative pose; loop closures beyond the 8 block slots
    p = make_synthetic(n=30, l=0, m=0, d=3, seed=1)
    p.update_problem_data()
    _operator_suite(p, 4)
    p = make_synthetic(n=2, l=0, m=0, d=2, seed=2)
    p.update_problem_data()
    _operator_suite(p, 2)
    p = make_synthetic(n=60, l=3, m=40, d=3, seed=5, loop_closures=[(0, j) for j in range(2, 60, 2)])
    p.update_problem_data()
    _operator_suite(p, 5)
    p = make_synthetic(n=300, l=2, m=260, d=2, seed=9)  # many ranges per landmark: hub rows
    p.update_problem_data()
    _operator_suite(p, 3)


def test_shape_errors(lib):
    from cora_b200 import capi
    g, p = load_fixture("small_ra_slam_problem")
    p.preconditioner = co.JACOBI
    p.update_problem_data()
    with make_handle(p) as h:
        with pytest.raises(capi.InvalidArgument):  # MatrixShapeException in the reference
            h.evaluate_objective(np.zeros((p.N + 1, 2)))
        with pytest.raises(capi.InvalidArgument):
            h.hessvec(np.zeros((p.N, 2)), None, np.zeros((p.N, 3)))
    with pytest.raises(capi.NotImplementedInReference):
        make_handle(p, preconditioner=capi.PRECON_BLOCK_CHOLESKY)


def test_large_synthetic_properties(lib):
    """BASELINE-size check (100k poses) through size-independent properties: the noise-free
    ground truth is in the null space scaled by noise, Q is symmetric (<u,Qv> == <v,Qu>),
    and the Hessian operator is self-adjoint on the tangent space."""
    p = make_synthetic(n=100_000, l=10, m=20_000, d=3, seed=42)
    p.update_problem_data()
    rng = np.random.default_rng(0)
    r = 5
    N = p.N
    U, V = rng.standard_normal((N, r)), rng.standard_normal((N, r))
    with make_handle(p) as h:
        QU, QV = h.data_matrix_product(U), h.data_matrix_product(V)
        a, b = float(np.sum(V * QU)), float(np.sum(U * QV))
        assert abs(a - b) <= 1e-10 * max(abs(a), abs(b))
        _close(QU, p.Q @ U, 1e-10)
        Y = h.project_to_manifold(rng.uniform(-1, 1, size=(N, r)))
        eg = h.euclidean_gradient(Y)
        T1 = h.tangent_space_projection(Y, U)
        T2 = h.tangent_space_projection(Y, V)
        H1, H2 = h.hessvec(Y, eg, T1), h.hessvec(Y, eg, T2)
        a, b = float(np.sum(T2 * H1)), float(np.sum(T1 * H2))
        assert abs(a - b) <= 1e-9 * max(abs(a), abs(b))
        # idempotence of the projections
        _close(h.tangent_space_projection(Y, T1), T1, 1e-12)
        _close(h.project_to_manifold(Y), Y, 1e-12)


def test_regularized_cholesky_preconditioner(lib):
    """Preconditioner::RegularizedCholesky (src/CORA_problem.cpp:544-614): (Q + lambda I) with the
    last row pinned, solved exactly by the chain Cholesky; the oracle solves the same matrix by
    sparse LU.  lambda is passed in (the reference's own estimate starts from a random vector)."""
    from cora_b200 import capi
    cases = [load_dataset("plaza2", preconditioner=co.REG_CHOLESKY),
             load_dataset("single_drone", preconditioner=co.REG_CHOLESKY),
             make_synthetic(n=3000, l=6, m=900, d=3, seed=2, preconditioner=co.REG_CHOLESKY),
             make_synthetic(n=700, l=0, m=0, d=2, seed=3, preconditioner=co.REG_CHOLESKY),
             make_synthetic(n=20, l=2, m=12, d=3, seed=4, preconditioner=co.REG_CHOLESKY)]
    rng = np.random.default_rng(0)
    for p in cases:
        p.update_problem_data()
        for r in (p.d, 5):
            V = rng.standard_normal((p.N, r))
            with make_handle(p, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
                # the device Lanczos estimate of ||Q||_2 agrees with the converged value to 1e-2
                assert abs(h.reg_lambda - p.lambda_reg) <= 1e-2 * p.lambda_reg
                h.reg_lambda = p.lambda_reg
                Z = h.precondition(V)
            ref = p.precondition(V)
            _close(Z, ref, 1e-8)
            assert np.all(Z[-1] == 0.0)  # CORA_preconditioners.cpp:77-80


def test_regularized_cholesky_falls_back_to_jacobi_without_factorisation(lib, monkeypatch):
    """RegularizedCholesky is the reference's default (src/pyfg_text_parser.cpp:116-120): on a graph without a device
    factorisation (here: a loop closure with the general sparse Cholesky switched off) the handle must still be
    created; it applies Jacobi and says so."""
    from cora_b200 import capi
    monkeypatch.setenv("CORA_B200_GENERAL_CHOLESKY", "0")
    p = make_synthetic(n=60, l=3, m=40, d=3, seed=5, loop_closures=[(0, 30)])
    p.update_problem_data()
    V = np.asfortranarray(np.random.default_rng(0).standard_normal((p.N, 4)))
    with make_handle(p, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
        assert h.effective_preconditioner == capi.PRECON_JACOBI
        Z = h.precondition(V)
    assert np.abs(Z - V / p.Q.diagonal()[:, None]).max() <= 1e-12 * np.abs(Z).max()
    pc = make_synthetic(n=60, l=3, m=40, d=3, seed=5)
    pc.update_problem_data()
    with make_handle(pc, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
        assert h.effective_preconditioner == capi.PRECON_REG_CHOLESKY


@pytest.mark.gpu
@pytest.mark.parametrize("d,r", [(2, 3), (3, 5), (3, 7), (3, 12)])
def test_persistent_spmm_kernel(lib, d, r):
    """The tile-pipelined data-matrix product that scripts/sweep_1m.py times (k_spmm_persistent) against Q @ X,
    including landmark hub rows (more than 64 couplings) and a range-row tail."""
    from synth import make_synthetic
    p = make_synthetic(n=1500, l=3, m=900, d=d, seed=21, rank=r)
    p.update_problem_data()
    X = np.asfortranarray(np.random.default_rng(5).standard_normal((p.N, r)))
    with make_handle(p) as h:
        h.set_iterate(X)
        assert h.spmm_resident(2) > 0
        got = h.get_work_vector(1, r)
    ref = p.Q @ X
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.gpu
def test_full_size_tnt_against_cpu_port(lib):
    """BASELINE configs[2] at full size (100k poses): the leading trust-region iterations of the persistent kernel
    against the C++ CPU restatement (oracle/cpu_ref.cpp), and size-independent properties of the result."""
    import os
    from cora_b200 import capi, synthetic
    from oracle import cpu_ref
    d, n, l, m, r = 3, 100_000, 10, 20_000, 5
    arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=42)
    Q = capi.assemble(d, n, l, arrays)
    m = len(arrays["rg_w"])
    x0 = synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=0)
    prm = capi.default_tnt_params(max_iterations=8, max_TPCG_iterations=10, max_computation_time=0.0)
    with capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_JACOBI) as h:
        x0 = h.project_to_manifold(x0)
        got = h.tnt(x0, prm)
        f_check = h.evaluate_objective(got.x)
    R = cpu_ref.CpuRef(d, n, m, n + l, Q, preconditioner=1, threads=os.cpu_count() or 1)
    ref = R.tnt(x0, prm)
    assert got.inner_iterations == ref.inner_iterations
    np.testing.assert_allclose(got.objective_values, ref.objective_values, rtol=1e-8)
    np.testing.assert_allclose(got.trust_region_radius, ref.trust_region_radius, rtol=1e-8)
    ov = got.objective_values
    assert all(ov[i + 1] <= ov[i] for i in range(len(ov) - 1))
    assert abs(f_check - got.f) <= 1e-8 * abs(got.f)   # f = 1/2 <x, Qx> cancels terms of order 1e9 here
    B = got.x[: d * n].reshape(n, d, r)
    assert np.abs(np.einsum("nir,njr->nij", B, B) - np.eye(d)).max() < 1e-10     # rotations on the Stiefel manifold
    rows = got.x[d * n: d * n + m]
    assert np.abs(np.linalg.norm(rows, axis=1) - 1).max() < 1e-12               # range rows on the sphere


@pytest.mark.gpu
def test_row_partition_device_buffers(lib):
    """The device side of the row-partitioned product (cora_b200/rowpart.py) on one GPU: the raw device
    buffers + row order the C-ABI exposes, a two-slab partition whose 'exchange' is done by hand between two handles of
    this process, against the full product."""
    import torch
    from cora_b200 import capi, rowpart, synthetic
    d, n, l, m, r = 3, 900, 4, 500, 5
    arrays, _ = synthetic.make_arrays(n, l, m, d=d, seed=3, loop_closures=[(10, 700), (449, 451)])
    Q = capi.assemble(d, n, l, arrays)
    X = np.random.default_rng(0).standard_normal((Q.shape[0], r))
    with capi.Handle(d, n, len(arrays["rg_w"]), n + l, Q, preconditioner=capi.PRECON_JACOBI) as hf:
        Y = hf.data_matrix_product(X)
    parts = [rowpart.LocalProblem(d, n, l, arrays, 2, g) for g in range(2)]
    lm = 0
    for P in parts:
        Ql = capi.assemble(d, P.n_loc, l, P.arrays)
        with capi.Handle(d, P.n_loc, P.m_loc, P.n_loc + l, Ql, preconditioner=capi.PRECON_JACOBI) as h:
            h.set_iterate(np.asfortranarray(X[P.local_to_global]))   # (ghost rows filled from the global vector)
            no_exchange = [[np.zeros(0, np.int64)] * 2] * 2
            op, x, y, row_of = rowpart.device_product(h, P, no_exchange, r, None)
            assert x.shape == (P.N_loc, r) and x.is_cuda
            np.testing.assert_array_equal(x.cpu().numpy()[row_of], X[P.local_to_global])   # the buffer IS the iterate
            op.product_fn()
            torch.cuda.synchronize()
            yl = y.cpu().numpy()[row_of]
            own = P.owned
            assert np.abs(yl[own] - Y[P.local_to_global[own]]).max() <= 1e-12 * np.abs(Y).max()
            lm = lm + yl[P.landmark_rows]
    assert np.abs(lm - Y[parts[0].local_to_global[parts[0].landmark_rows]]).max() <= 1e-12 * np.abs(Y).max()


@pytest.mark.gpu
def test_peer_memory_product_world_size_one(lib):
    """cora_b200_peer_* (the row-partitioned product with the library's own exchange kernels, peer_product.cuh) at
    world size 1: flag barrier with itself, no ghosts, landmark 'sum' over one rank -- equals the plain product.
    (N > 1 runs on 2/4/8 GPUs by scripts/rowpart_bench.py, which asserts 1e-12 against the full product.)"""
    from cora_b200 import capi, rowpart, synthetic
    d, n, l, m, r = 3, 700, 3, 300, 5
    arrays, _ = synthetic.make_arrays(n, l, m, d=d, seed=4)
    Q = capi.assemble(d, n, l, arrays)
    X = np.random.default_rng(1).standard_normal((Q.shape[0], r))
    parts = [rowpart.LocalProblem(d, n, l, arrays, 1, 0)]
    with capi.Handle(d, n, len(arrays["rg_w"]), n + l, Q, preconditioner=capi.PRECON_JACOBI) as h:
        Y = h.data_matrix_product(X)
        h.set_iterate(np.asfortranarray(X))
        pp, row_of = rowpart.peer_product(h, parts, 0, r, None)
        assert pp.product(3) > 0
        Y2 = h.get_work_vector(1, r)
        pp.close()
    assert np.abs(Y2 - Y).max() <= 1e-13 * np.abs(Y).max()


@pytest.mark.gpu
def test_million_pose_spmm_against_scipy(lib):
    """BASELINE configs[4] at full size: the persistent SpMM kernel (k_spmm_persistent, the kernel behind
    `roofline_cfg5`) on the 1M-pose problem (N = 4.2M, nnz = 43.4M, rank 5) against SciPy's CSR product of the same
    matrix -- Problem::dataMatrixProduct, src/CORA_problem.cpp:742-757.  fp64, only the summation order differs."""
    from cora_b200 import capi, synthetic
    d, n, l, m, r = 3, 1_000_000, 100, 200_000, 5
    arrays, _ = synthetic.make_arrays(n, l, m, d=d, seed=42)
    Q = capi.assemble(d, n, l, arrays)
    m = len(arrays["rg_w"])
    X = np.random.default_rng(0).standard_normal((Q.shape[0], r))
    with capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_JACOBI) as h:
        h.set_iterate(np.asfortranarray(X))
        assert h.spmm_resident(2) > 0
        Y = h.get_work_vector(1, r)
    ref = Q @ X
    assert np.abs(Y - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.gpu
def test_full_size_regularized_cholesky_against_sparse_lu(lib):
    """BASELINE configs[2] at full size (100k poses) with the reference's default preconditioner: the device chain
    factorisation + apply against the NumPy oracle's sparse LU of (Q + lambda I)[:-1, :-1]
    (src/CORA_problem.cpp:544-614, src/CORA_preconditioners.cpp:46-83), and the leading trust-region iterations with
    that preconditioner against the oracle's TNT (same lambda on both sides)."""
    from cora_b200 import capi, synthetic
    d, n, l, m, r = 3, 100_000, 10, 20_000, 5
    arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=42)
    p = co.Problem.from_arrays(d, n, l, arrays, rank=r, preconditioner=co.REG_CHOLESKY)
    p.chol_ordering = "pose_major"   # same matrix and solution; SuperLU's default ordering takes minutes at this size
    # (a wide initial trust region: with Delta0 = 5 the first outer iterations are boundary steps with no CG at all)
    prm = dict(max_iterations=4, max_TPCG_iterations=10, Delta0=1e5)
    V = np.random.default_rng(1).standard_normal((p.N, r))
    Q = capi.assemble(d, n, l, arrays)
    with capi.Handle(d, n, len(arrays["rg_w"]), n + l, Q, preconditioner=capi.PRECON_REG_CHOLESKY) as h:
        assert h.effective_preconditioner == capi.PRECON_REG_CHOLESKY
        p.lambda_reg = h.reg_lambda
        p.update_problem_data()
        assert abs(Q - p.Q).max() <= 1e-9 * abs(p.Q).max()
        x0 = p.project_to_manifold(synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=0))
        Z = h.precondition(V)
        got = h.tnt(x0, capi.default_tnt_params(max_computation_time=0.0, **prm))
    ref_pre = p.precondition(V)
    assert np.abs(Z - ref_pre).max() <= 1e-8 * np.abs(ref_pre).max()
    assert np.all(Z[-1] == 0.0)
    ref = co.problem_tnt(p, x0, co.cora_tnt_params(**prm))
    assert got.inner_iterations == ref.inner_iterations and sum(ref.inner_iterations) > 0
    np.testing.assert_allclose(got.objective_values, ref.objective_values, rtol=1e-7)
    np.testing.assert_allclose(got.preconditioned_gradient_norms[:3], ref.preconditioned_gradient_norms[:3], rtol=1e-6)
