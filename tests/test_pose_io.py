"""Host utilities either side of the solver: odometry initialisation
(examples/paper_experiments.cpp:426-534) and TUM / g2o export (src/CORA_utils.cpp:204-350).  CPU only."""
import numpy as np
import pytest

from conftest import load_dataset
from oracle import cora_oracle as co


@pytest.mark.parametrize("d", [2, 3])
def test_odometry_initialization_composes_the_chain(lib, d):
    from cora_b200 import capi, synthetic
    n, l, m, r = 200, 3, 80, d + 2
    arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=3)
    X = capi.odometry_initialization(d, n, l, arrays, r, seed=5)
    ref = synthetic.odometry_initialization(d, n, l, arrays, r, seed=5)   # NumPy restatement of the same function
    N = d * n + m + n + l
    assert X.shape == (N, r)
    # invariant to the random SO(r) factor: Gram matrix of the pose rows (rotations + translations)
    rows = np.r_[np.arange(d * n), d * n + m + np.arange(n)]
    G, Gref = X[rows] @ X[rows].T, ref[rows] @ ref[rows].T
    assert np.abs(G - Gref).max() <= 1e-9 * np.abs(Gref).max()
    # rotation blocks are orthonormal, range rows unit, the chain reproduces the odometry measurements
    B = X[: d * n].reshape(n, d, r)
    assert np.abs(np.einsum("nik,njk->nij", B, B) - np.eye(d)).max() < 1e-12
    assert np.abs(np.linalg.norm(X[d * n: d * n + m], axis=1) - 1).max() < 1e-12
    p = co.Problem.from_arrays(d, n, l, arrays, rank=r, preconditioner=co.JACOBI)
    p.update_problem_data()
    assert np.isfinite(p.evaluate_objective(X))
    # sign of the range rows: the default is the one the data matrix implies (lower cost than the reference's)
    Xr = capi.odometry_initialization(d, n, l, arrays, r, seed=5, reference_sign=True)
    assert p.evaluate_objective(X) < p.evaluate_objective(Xr)


def test_odometry_initialization_rejects_bad_input(lib):
    from cora_b200 import capi, synthetic
    arrays, _ = synthetic.make_arrays(20, 1, 5, d=3, seed=1)
    with pytest.raises(capi.InvalidArgument):
        capi.odometry_initialization(3, 20, 1, arrays, 2)   # rank < dim


@pytest.mark.parametrize("d", [2, 3])
def test_save_solution_tum_and_g2o(lib, tmp_path, d):
    from cora_b200 import capi, synthetic
    n, l, m = 30, 2, 10
    arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=2)
    R, t, L = gt
    N = d * n + m + n + l
    X = np.zeros((N, d))
    X[: d * n] = np.transpose(R, (0, 2, 1)).reshape(d * n, d)
    X[d * n + m: d * n + m + n] = t
    tum, g2o = tmp_path / "a.tum", tmp_path / "a.g2o"
    capi.save_solution(tum, X, d, n, m, n + l, "tum")
    capi.save_solution(g2o, X, d, n, m, n + l, "g2o", first=5, count=10)
    rows = np.loadtxt(tum)
    assert rows.shape == (n, 8) and np.array_equal(rows[:, 0], np.arange(n))
    assert np.allclose(rows[:, 1:1 + d], t, rtol=1e-5, atol=1e-5)   # default ostream precision: 6 significant digits
    q = rows[:, 4:8]
    assert np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-5)
    # quaternion -> rotation reproduces R (x y z w order)
    x, y, z, w = q.T
    Rq = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                   np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                   np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)
    assert np.allclose(Rq[:, :d, :d], R, atol=2e-5)
    lines = g2o.read_text().strip().splitlines()
    assert len(lines) == 10 and lines[0].split()[0] == ("VERTEX_SE3:QUAT" if d == 3 else "VERTEX_SE2")
    assert lines[0].split()[1] == "0"
    if d == 2:
        th = float(lines[0].split()[4])
        assert abs(th - np.arctan2(R[5][1, 0], R[5][0, 0])) < 1e-5
    # getRotation's checks (src/CORA_utils.cpp:219-229): a reflected block is rejected
    Xb = X.copy(); Xb[0, :] *= -1
    with pytest.raises(capi.CoraB200Error):
        capi.save_solution(tum, Xb, d, n, m, n + l, "tum")
