"""General sparse block Cholesky of the pose system (cora_b200/csrc/gen_chol.hpp: nested-dissection order,
left-looking numeric factorisation, level/cluster schedule) for graphs that are NOT one odometry chain -- loop
closures, several robots (TIERS, MR.CLAM) -- executed on the HOST through the test hook and pinned
against the oracle's sparse LU of the same matrix (src/CORA_problem.cpp:544-614,
src/CORA_preconditioners.cpp:16-83; the reference tests Cholesky solves the same way, tests/test.cpp:25-214)."""
import numpy as np
import pytest

from conftest import load_dataset
from oracle import cora_oracle as co
from synth import make_synthetic


def _loops(n, k, seed, near=None):
    rng = np.random.default_rng(seed)
    out = set()
    while len(out) < k:
        i = int(rng.integers(0, n))
        j = int(rng.integers(0, n)) if near is None else int(np.clip(i + rng.integers(-near, near + 1), 0, n - 1))
        if abs(i - j) > 1:
            out.add((min(i, j), max(i, j)))
    return sorted(out)


def _cases():
    yield "one loop closure", make_synthetic(n=60, l=3, m=40, d=3, seed=5, loop_closures=[(0, 30)])
    yield "d2 random loops", make_synthetic(n=400, l=2, m=150, d=2, seed=3, loop_closures=_loops(400, 60, 1))
    yield "d3 random loops", make_synthetic(n=700, l=4, m=300, d=3, seed=11, loop_closures=_loops(700, 70, 2))
    yield "d3 near loops, no landmarks", make_synthetic(n=900, l=0, m=0, d=3, seed=4, loop_closures=_loops(900, 200, 3, near=12))
    yield "d3 hub pose (> 8 couplings)", make_synthetic(n=300, l=2, m=100, d=3, seed=6,
                                                        loop_closures=[(7, j) for j in range(20, 300, 9)])
    yield "tiers", load_dataset("tiers")
    yield "mrclam2", load_dataset("mrclam2")


@pytest.mark.parametrize("name,p", list(_cases()), ids=[c[0] for c in _cases()])
def test_general_cholesky_solve_matches_sparse_lu(lib, name, p):
    from cora_b200 import capi
    p.preconditioner = co.REG_CHOLESKY
    p.update_problem_data()
    rng = np.random.default_rng(0)
    for r in (1, 5):
        V = rng.standard_normal((p.N, r))
        pd, Z = capi.debug_chain_host(p.d, p.n, p.m, p.n + p.l, p.Q, p.lambda_reg, True, V)
        assert pd
        ref = p.precondition(V)   # splu((Q + lambda I)[:-1, :-1]); last row pinned to 0
        assert np.abs(Z - ref).max() <= 1e-8 * np.abs(ref).max()
        assert not Z[-1].any()


def test_general_psd_verdict_matches_dense_eigenvalues(lib):
    """S + eta I positive definite <=> lambda_min(S) + eta > 0 (src/CORA_utils.cpp:33-57), loop-closure graph."""
    from cora_b200 import capi
    p = make_synthetic(n=120, l=3, m=60, d=3, seed=8, loop_closures=_loops(120, 25, 7))
    p.update_problem_data()
    p.rank = 4
    Y = p.random_initial_guess(np.random.default_rng(0))
    S = p.certificate_matrix(Y)
    w = np.linalg.eigvalsh(S.toarray())
    for shift in (0.0, -w[0] * 0.5, -w[0] * 0.99, -w[0] * 1.01, -w[0] * 2.0):
        pd, _ = capi.debug_chain_host(p.d, p.n, p.m, p.n + p.l, S, shift, False)
        assert pd == bool(w[0] + shift > 0), (shift, w[0])
    pd, _ = capi.debug_chain_host(p.d, p.n, p.m, p.n + p.l, p.Q, 1e-6, False)
    assert pd
    pd, _ = capi.debug_chain_host(p.d, p.n, p.m, p.n + p.l, p.Q, -1e-3, False)
    assert not pd


def _two_robots(n=300, cut=140, loops=((5, 100), (20, 60), (150, 290), (200, 260))):
    """Two robots: the odometry chain is cut between poses cut-1 and cut, loop closures inside each robot only, so the
    pose graph has two connected components (coupled through the landmarks alone)."""
    from cora_b200 import synthetic
    arrays, gt = synthetic.make_arrays(n, 3, 150, d=3, seed=9, loop_closures=list(loops))
    keep_rp = ~((arrays["rp_i"] == cut - 1) & (arrays["rp_j"] == cut))
    keep_rot = ~((arrays["rot_i"] == cut - 1) & (arrays["rot_j"] == cut))
    a = dict(arrays)
    for k in ("rp_i", "rp_j", "rp_t", "rp_tau"):
        a[k] = arrays[k][keep_rp]
    for k in ("rot_i", "rot_j", "rot_R", "rot_kappa"):
        a[k] = arrays[k][keep_rot]
    return co.Problem.from_arrays(3, n, 3, a, preconditioner=co.REG_CHOLESKY)


@pytest.mark.parametrize("name,p", [("two components", _two_robots()),
                                    ("three poses, one closure", make_synthetic(n=3, l=1, m=2, d=2, seed=1, loop_closures=[(0, 2)])),
                                    ("dense little graph", make_synthetic(n=9, l=0, m=0, d=3, seed=2,
                                                                          loop_closures=[(i, j) for i in range(9) for j in range(i + 2, 9)]))],
                         ids=["two components", "three poses", "dense"])
def test_general_cholesky_edge_cases(lib, name, p):
    from cora_b200 import capi
    p.preconditioner = co.REG_CHOLESKY
    p.update_problem_data()
    st = capi.debug_factor_stats(p.d, p.n, p.m, p.n + p.l, p.Q)
    assert st["chain"] == 0 and st["poses"] == p.n
    V = np.random.default_rng(0).standard_normal((p.N, 3))
    pd, Z = capi.debug_chain_host(p.d, p.n, p.m, p.n + p.l, p.Q, p.lambda_reg, True, V)
    assert pd
    ref = p.precondition(V)
    assert np.abs(Z - ref).max() <= 1e-8 * np.abs(ref).max()
