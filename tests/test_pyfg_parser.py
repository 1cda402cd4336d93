"""The C++ PyFG parser (cora_b200/csrc/pyfg.hpp) against the oracle's parser, which the
reference's own assembly goldens pin (tests/test_oracle_golden.py): same variable order, same stacks,
same data matrix.  CPU only."""
import numpy as np
import pytest

from conftest import FIXTURES, load_fixture
from oracle import cora_oracle as co


def _same_arrays(A, B):
    for k in ("rp_i", "rp_j", "rot_i", "rot_j", "rg_a", "rg_b"):
        assert np.array_equal(np.asarray(A[k]), np.asarray(B[k])), k          # index maps: bit exact
    for k in ("rp_t", "rp_tau", "rot_R", "rot_kappa", "rg_r", "rg_w"):
        a, b = np.asarray(A[k], dtype=float), np.asarray(B[k], dtype=float)
        assert a.shape == b.shape, k
        assert np.allclose(a, b, rtol=1e-15, atol=0), k


@pytest.mark.parametrize("name", FIXTURES)
def test_reference_fixture_files(lib, name):
    from cora_b200 import capi
    g, p = load_fixture(name)
    text = str(g["pyfg"])
    d, n, l, A = capi.parse_pyfg(text, from_text=True)
    assert (d, n, l) == (p.d, p.n, p.l)
    _same_arrays(A, p.measurement_arrays())
    # and the assembled data matrix equals the reference's DataMatrix.mm golden
    Q = capi.assemble(d, n, l, A)
    assert abs(Q - g_matrix(g)).max() < 1e-12


def g_matrix(g):
    import scipy.sparse as sp
    return sp.csr_matrix(np.asarray(g["DataMatrix"])) if np.asarray(g["DataMatrix"]).ndim == 2 else g["DataMatrix"]


def _quat(R):
    # rotation matrix -> (x, y, z, w), Shepperd's method
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        return np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
    q = np.zeros(4)
    q[i] = 0.25 * s
    q[j] = (R[j, i] + R[i, j]) / s
    q[k] = (R[k, i] + R[i, k]) / s
    q[3] = (R[k, j] - R[j, k]) / s
    return q


def _f(*vals):
    return " ".join(repr(float(v)) for v in vals)


def _se3_text(n=12, l=2, with_priors=True, seed=3):
    """A synthetic SE(3) PyFG file exercising every 3-D line type of src/pyfg_text_parser.cpp:122-135."""
    from cora_b200 import synthetic
    rng = np.random.default_rng(seed)
    arrays, (R, t, L) = synthetic.make_arrays(n, l, 6, d=3, seed=seed)
    lines = []
    for i in range(n):
        q = _quat(R[i])
        lines.append("VERTEX_SE3:QUAT %.3f A%d %s" % (i * 0.1, i, _f(*t[i], *q)))
    for j in range(l):
        lines.append("VERTEX_XYZ L%d %s" % (j, _f(*L[j])))
    cov6 = np.diag([0.01, 0.02, 0.03, 0.001, 0.002, 0.003])
    up6 = " ".join(repr(float(cov6[i, j])) for i in range(6) for j in range(i, 6))
    for k in range(len(arrays["rot_kappa"])):
        i, j = int(arrays["rot_i"][k]), int(arrays["rot_j"][k])
        q = _quat(arrays["rot_R"][k])
        lines.append("EDGE_SE3:QUAT %.3f A%d A%d %s %s" % (k * 0.1, i, j, _f(*arrays["rp_t"][k], *q), up6))
    up3 = "0.04 0 0 0.05 0 0.06"
    lines.append("EDGE_SE3_XYZ 0.5 A3 L1 %s %s" % (_f(*rng.standard_normal(3)), up3))
    for k in range(len(arrays["rg_w"])):
        lines.append("EDGE_RANGE %.3f A%d L%d %s" % (k * 0.1, int(arrays["rg_a"][k]), int(arrays["rg_b"][k]) - n,
                                                     _f(arrays["rg_r"][k], 0.09)))
    if with_priors:
        lines.append("VERTEX_SE3:QUAT:PRIOR 0.0 A0 0 0 0 0 0 0 1 %s" % up6)
        lines.append("VERTEX_XYZ:PRIOR 0.0 L0 %s %s" % (_f(*L[0]), up3))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("with_priors", [False, True])
def test_se3_file_against_oracle_parser(lib, with_priors):
    from cora_b200 import capi
    text = _se3_text(with_priors=with_priors)
    p = co.parse_pyfg(text, from_text=True)
    d, n, l, A = capi.parse_pyfg(text, from_text=True)
    assert (d, n, l) == (3, p.n, p.l)
    if with_priors:
        assert n == 13      # the auto-added origin pose O0 (src/CORA_problem.cpp:80-86)
    _same_arrays(A, p.measurement_arrays())
    p.preconditioner = co.JACOBI
    p.update_problem_data()
    Q = capi.assemble(d, n, l, A)
    assert abs(Q - p.Q).max() <= 1e-12 * abs(p.Q).max()


def test_errors_match_the_reference(lib, tmp_path):
    from cora_b200 import capi
    with pytest.raises(capi.CoraB200Error):          # unknown keyword: std::runtime_error (:157-159)
        capi.parse_pyfg("VERTEX_SE2 0 A0 0 0 0\nBOGUS 1 2 3\n", from_text=True)
    with pytest.raises(capi.InvalidArgument):        # duplicate variable: std::invalid_argument
        capi.parse_pyfg("VERTEX_SE2 0 A0 0 0 0\nVERTEX_SE2 1 A0 0 0 0\n", from_text=True)
    with pytest.raises(capi.InvalidArgument):        # duplicate measurement
        capi.parse_pyfg("VERTEX_SE2 0 A0 0 0 0\nVERTEX_XY L0 1 1\nEDGE_RANGE 0 A0 L0 1.0 0.1\nEDGE_RANGE 1 A0 L0 1.1 0.1\n",
                        from_text=True)
    with pytest.raises(capi.CoraB200Error):          # missing file
        capi.parse_pyfg(str(tmp_path / "nope.pyfg"))
    f = tmp_path / "ok.pyfg"
    f.write_text("VERTEX_SE2 0 A0 0 0 0\nVERTEX_SE2 1 A1 1 0 0\nEDGE_SE2 0 A0 A1 1 0 0.1 0.01 0 0 0.01 0 0.001\n")
    d, n, l, A = capi.parse_pyfg(str(f))
    assert (d, n, l) == (2, 2, 0) and A["rp_tau"][0] == pytest.approx(2 / 0.02) and A["rot_kappa"][0] == pytest.approx(1000.0)


@pytest.mark.parametrize("name", ["plaza2", "single_drone"])
def test_real_datasets_when_the_reference_tree_is_present(lib, name):
    """The committed plaza2 / single_drone stacks (tests/golden/*.npz, made from the reference's
    examples/data by tests/golden/make_golden.py) against the C++ parser reading the same file.  Only where
    /root/reference exists (this container); skipped on the GPU box."""
    import os
    from cora_b200 import capi
    path = "/root/reference/examples/data/%s.pyfg" % name
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    d, n, l, A = capi.parse_pyfg(path)
    assert (d, n, l) == (int(g["d"]), int(g["n"]), int(g["l"]))
    _same_arrays(A, {k: g[k] for k in A})
