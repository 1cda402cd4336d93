"""The chain Cholesky (RegularizedCholesky preconditioner + PSD test of the certificate) executed
on the HOST through the same per-chunk routines the device kernels call, against the oracle's sparse
LU of the same matrix.  CPU only (test hook cora_b200_debug_chain_host)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_dataset, load_fixture
from oracle import cora_oracle as co
from synth import make_synthetic


def _cases():
    yield "small_ra_slam", load_fixture("small_ra_slam_problem")[1]
    yield "single_rpm", load_fixture("single_rpm")[1]
    yield "plaza2", load_dataset("plaza2")
    yield "single_drone", load_dataset("single_drone")
    yield "synthetic d3", make_synthetic(n=700, l=4, m=300, d=3, seed=11)
    yield "synthetic d2 no landmarks", make_synthetic(n=700, l=0, m=0, d=2, seed=11)
    yield "synthetic d2", make_synthetic(n=5000, l=7, m=3000, d=2, seed=1)
    yield "33 poses (2 levels)", make_synthetic(n=33, l=1, m=20, d=3, seed=1)
    yield "17 poses (1 level)", make_synthetic(n=17, l=2, m=10, d=3, seed=1)
    yield "513 poses (3 levels)", make_synthetic(n=513, l=3, m=100, d=3, seed=2)


@pytest.mark.parametrize("name,p", list(_cases()), ids=[c[0] for c in _cases()])
def test_regularized_cholesky_solve_matches_sparse_lu(lib, name, p):
    from cora_b200 import capi
    p.preconditioner = co.REG_CHOLESKY
    p.update_problem_data()
    rng = np.random.default_rng(0)
    for r in (1, p.d, 5):
        V = rng.standard_normal((p.N, r))
        pd, Z = capi.debug_chain_host(p.d, p.n, p.m, p.n + p.l, p.Q, p.lambda_reg, True, V)
        assert pd
        ref = p.precondition(V)   # splu((Q + lambda I)[:-1, :-1]); last row pinned to 0
        assert np.abs(Z - ref).max() <= 1e-9 * np.abs(ref).max()
        assert not Z[-1].any()


def test_psd_verdict_matches_dense_eigenvalues(lib):
    """S + eta I positive definite <=> lambda_min(S) + eta > 0 (src/CORA_utils.cpp:33-57)."""
    from cora_b200 import capi
    p = make_synthetic(n=120, l=3, m=60, d=3, seed=8)
    p.update_problem_data()
    p.rank = 4
    rng = np.random.default_rng(0)
    Y = p.random_initial_guess(rng)
    S = p.certificate_matrix(Y)
    w = np.linalg.eigvalsh(S.toarray())
    for shift in (0.0, -w[0] * 0.5, -w[0] * 0.99, -w[0] * 1.01, -w[0] * 2.0):
        pd, _ = capi.debug_chain_host(p.d, p.n, p.m, p.n + p.l, S, shift, False)
        assert pd == bool(w[0] + shift > 0), (shift, w[0])
    # the data matrix itself is PSD with a nontrivial kernel: Q + eps I is PD, Q - eps I is not
    pd, _ = capi.debug_chain_host(p.d, p.n, p.m, p.n + p.l, p.Q, 1e-6, False)
    assert pd
    pd, _ = capi.debug_chain_host(p.d, p.n, p.m, p.n + p.l, p.Q, -1e-3, False)
    assert not pd
