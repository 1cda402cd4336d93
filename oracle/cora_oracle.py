"""CPU oracle for the CORA Riemannian-staircase inner loop (NumPy / SciPy).

TEST INFRASTRUCTURE ONLY.  This module is a plain CPU restatement of the
reference algorithm (MarineRoboticsGroup/cora @ 015dc43).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it; nothing under ``cora_b200/`` does, and the
product path fails loudly when its CUDA library is missing.

Parity status
-------------
* The reference itself cannot be built here (needs Eigen3 + SuiteSparse, neither
  is installed), so the oracle is pinned by the reference's own
  golden fixtures (``tests/data/*/*.mm`` -> ``tests/golden/*.npz``): Q assembly
  and all seven sub-matrices, f, egrad, rgrad, Hess-vec, the certificate matrix
  S, Q*X_gt = 0, and the min-eigenpair known answers of
  ``tests/test_certification.cpp``; STPCG/TNT by the known answers in
  ``libs/Optimization/tests``.
* PARITY UNPINNED (no reference test asserts them): the polar retraction, the
  TNT trajectory on CORA problems, and the eigen-search inside
  ``fast_verification`` (the reference uses SYM-ILDL-preconditioned LOBPCG; the
  oracle uses the same LOBPCG but an exact-factorisation preconditioner).

Every function cites the reference file:line it follows (paths relative to the
reference root).  Row form is used throughout: Y is N x r, pose i owns rows
[d*i, d*i+d), range factor k owns row d*n+k, translations own the last n+l rows
(src/CORA_problem.cpp:964-1021).
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --------------------------------------------------------------------------- #
#  Measurements (include/CORA/Measurements.h)                                 #
# --------------------------------------------------------------------------- #


def rot_precision(cov: np.ndarray) -> float:
    """RelativePoseMeasurement::getRotPrecision, Measurements.h:79-93."""
    if cov.shape[0] == 6:
        return 1.5 / (cov[3, 3] + cov[4, 4] + cov[5, 5])
    if cov.shape[0] == 3:
        return 1.0 / cov[2, 2]
    raise RuntimeError("getRotPrecision() only implemented for 2D and 3D rotations")


def trans_precision(cov: np.ndarray, dim: int) -> float:
    """getTransPrecision, Measurements.h:109-112,134-137."""
    return float(dim) / float(np.trace(cov[:dim, :dim]))


@dataclass
class RelPose:  # RelativePoseMeasurement / PosePrior (first = origin)
    a: str
    b: str
    R: np.ndarray
    t: np.ndarray
    cov: np.ndarray


@dataclass
class RelPoseLandmark:  # RelativePoseLandmarkMeasurement / LandmarkPrior
    a: str
    b: str
    t: np.ndarray
    cov: np.ndarray


@dataclass
class Range:
    a: str
    b: str
    r: float
    cov: float


# --------------------------------------------------------------------------- #
#  Problem (src/CORA_problem.cpp)                                             #
# --------------------------------------------------------------------------- #

JACOBI = 1
BLOCK_CHOLESKY = 2  # broken in the reference; not implemented
REG_CHOLESKY = 3


EXPLICIT, IMPLICIT = 0, 1  # include/CORA/CORA_types.h:52-56


def pose_major_order(d, n, l, m):
    """A permutation of the reference rows [d n | m | n | l]: range rows, then per pose its d rotation rows and its
    translation row, then the landmarks (the last reference row -- the pinned translation -- stays last)."""
    dn = d * n
    poses = np.concatenate([dn * 0 + (np.arange(n)[:, None] * d + np.arange(d)[None, :]),
                            (dn + m + np.arange(n))[:, None]], axis=1).reshape(-1)
    lm = dn + m + n + np.arange(l)
    if l == 0:  # the pinned row is the last pose's translation: already last in `poses`
        return np.concatenate([dn + np.arange(m), poses])
    return np.concatenate([dn + np.arange(m), poses, lm])


class _PermutedLU:
    """splu of P M P^T in the given order without further column permutation; solve() is in the original order."""

    def __init__(self, M, order):
        self.order = np.asarray(order)
        Mp = M[self.order][:, self.order].tocsc()
        self.lu = spla.splu(Mp, permc_spec="NATURAL", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))

    def solve(self, B):
        out = np.empty_like(B)
        out[self.order] = self.lu.solve(np.ascontiguousarray(B[self.order]))
        return out


class Problem:
    """Restatement of CORA::Problem (explicit formulation; the translation-implicit one through
    set_formulation(IMPLICIT), src/CORA_problem.cpp:714-757)."""

    def __init__(self, dim: int, rank: int, preconditioner: int = REG_CHOLESKY):
        assert rank >= dim  # CORA_problem.h:203
        self.formulation = EXPLICIT
        self.d = dim
        self.rank = rank
        self.preconditioner = preconditioner
        self.pose_idx: dict = {}
        self.landmark_idx: dict = {}
        self.rpms: List[RelPose] = []
        self.pose_priors: List[RelPose] = []
        self.rplms: List[RelPoseLandmark] = []
        self.landmark_priors: List[RelPoseLandmark] = []
        self.ranges: List[Range] = []
        self._pairs = set()
        self.has_priors = False
        self.Q: Optional[sp.csr_matrix] = None
        self.up_to_date = False
        self.reg_chol_max_cond = 1e6
        self.lambda_reg: Optional[float] = None
        self._chol = None
        self._jacobi = None
        self._arrays = None
        self._sizes = None

    @classmethod
    def from_arrays(cls, d, n, l, arrays, rank=None, preconditioner=REG_CHOLESKY):
        """Problem over pre-flattened measurement arrays (see measurement_arrays);
        used for the committed plaza2 / single_drone fixtures and the synthetic
        generator.  Variables are implied: poses 0..n-1, landmarks 0..l-1."""
        p = cls(int(d), int(rank if rank is not None else d), preconditioner)
        p._arrays = {k: np.asarray(v) for k, v in arrays.items()}
        p._sizes = (int(n), int(l), int(len(p._arrays["rg_w"])))
        return p

    # -- construction (CORA_problem.cpp:24-113), O(1) duplicate checks -------
    def add_pose(self, s: str):
        if s in self.pose_idx:
            raise ValueError("Pose variable already exists")
        self.pose_idx[s] = len(self.pose_idx)
        self.up_to_date = False

    def add_landmark(self, s: str):
        if s in self.landmark_idx:
            raise ValueError("Landmark variable already exists")
        self.landmark_idx[s] = len(self.landmark_idx)

    def _check_pair(self, kind, a, b, msg):
        key = (kind, a, b) if a <= b else (kind, b, a)
        if key in self._pairs:
            raise ValueError(msg)
        self._pairs.add(key)

    def add_range(self, m: Range):
        self._check_pair("r", m.a, m.b, "Range measurement already exists")
        self.ranges.append(m)
        self.up_to_date = False

    def add_rel_pose(self, m: RelPose):
        self._check_pair("p", m.a, m.b, "Relative pose measurement already exists")
        self.rpms.append(m)
        self.up_to_date = False

    def add_rel_pose_landmark(self, m: RelPoseLandmark):
        self._check_pair("pl", m.a, m.b, "Relative pose landmark measurement already exists")
        self.rplms.append(m)
        self.up_to_date = False

    def _add_origin(self):  # CORA_problem.cpp:80-86
        self.add_pose("O0")

    def add_pose_prior(self, sym, R, t, cov):
        self._check_pair("pp", sym, sym, "Pose prior already exists")
        self.pose_priors.append(RelPose("O0", sym, R, t, cov))
        self.up_to_date = False
        if not self.has_priors:
            self.has_priors = True
            self._add_origin()

    def add_landmark_prior(self, sym, p, cov):
        self._check_pair("lp", sym, sym, "Landmark prior already exists")
        self.landmark_priors.append(RelPoseLandmark("O0", sym, p, cov))
        self.up_to_date = False
        if not self.has_priors:
            self.has_priors = True
            self._add_origin()

    # -- sizes (CORA_problem.h:292-323, CORA_problem.cpp:940-942) -----------
    @property
    def n(self):
        return self._sizes[0] if self._sizes else len(self.pose_idx)

    @property
    def l(self):
        return self._sizes[1] if self._sizes else len(self.landmark_idx)

    @property
    def m(self):
        return self._sizes[2] if self._sizes else len(self.ranges)

    @property
    def dn(self):
        return self.d * self.n

    @property
    def N(self):
        return self.n * (self.d + 1) + self.l + self.m

    def rotation_idx(self, s):  # :964-974
        if s not in self.pose_idx:
            raise ValueError("Unknown pose symbol: " + s)
        return self.pose_idx[s]

    def translation_idx(self, s):  # :998-1021 (absolute row)
        off = self.dn + self.m
        if s in self.pose_idx:
            return off + self.pose_idx[s]
        if s in self.landmark_idx:
            return off + self.n + self.landmark_idx[s]
        raise ValueError("Unknown translation symbol")

    # -- measurement arrays -------------------------------------------------
    def measurement_arrays(self):
        """Flatten the measurement lists into index/value arrays in the order
        the reference stacks them (pose-pose, pose priors, pose-landmark,
        landmark priors; CORA_problem.cpp:190-294)."""
        if self._arrays is not None:
            return self._arrays
        d = self.d
        off = self.dn + self.m
        tr = lambda s: self.translation_idx(s) - off
        rp_i, rp_j, rp_t, rp_tau = [], [], [], []
        rot_i, rot_j, rot_R, rot_kappa = [], [], [], []
        for mm in list(self.rpms) + list(self.pose_priors):
            rp_i.append(tr(mm.a)); rp_j.append(tr(mm.b))
            rp_t.append(mm.t); rp_tau.append(trans_precision(mm.cov, d))
            rot_i.append(self.rotation_idx(mm.a)); rot_j.append(self.rotation_idx(mm.b))
            rot_R.append(mm.R); rot_kappa.append(rot_precision(mm.cov))
        for mm in list(self.rplms) + list(self.landmark_priors):
            rp_i.append(tr(mm.a)); rp_j.append(tr(mm.b))
            rp_t.append(mm.t); rp_tau.append(trans_precision(mm.cov, d))
        # NB the reference stacks pose priors *between* pose-pose and
        # pose-landmark rows; the order above (pp, priors, pl, lpriors) matches.
        rg_a = [tr(mm.a) for mm in self.ranges]
        rg_b = [tr(mm.b) for mm in self.ranges]
        rg_r = [mm.r for mm in self.ranges]
        rg_w = [1.0 / mm.cov for mm in self.ranges]  # Measurements.h:151
        A = lambda x, dt=float: np.asarray(x, dtype=dt)
        return dict(
            rp_i=A(rp_i, np.int64), rp_j=A(rp_j, np.int64),
            rp_t=A(rp_t).reshape(-1, d), rp_tau=A(rp_tau),
            rot_i=A(rot_i, np.int64), rot_j=A(rot_j, np.int64),
            rot_R=A(rot_R).reshape(-1, d, d), rot_kappa=A(rot_kappa),
            rg_a=A(rg_a, np.int64), rg_b=A(rg_b, np.int64), rg_r=A(rg_r), rg_w=A(rg_w))

    def submatrices(self):
        """The seven sub-matrices of fillRangeSubmatrices / fillRelPoseSubmatrices
        / fillRotConnLaplacian (CORA_problem.cpp:115-377), for the golden checks."""
        a = self.measurement_arrays()
        return build_submatrices(self.d, self.n, self.l, a)

    def update_problem_data(self):  # CORA_problem.cpp:500-510
        a = self.measurement_arrays()
        self.Q = assemble_Q(self.d, self.n, self.l, a)
        self._update_preconditioner()
        if self.formulation == IMPLICIT:  # :506-508
            self._fill_implicit()
        self.up_to_date = True

    # -- implicit formulation (CORA_problem.cpp:714-741) ----------------------
    @property
    def rot_and_range_size(self):
        return self.d * self.n + self.m

    @property
    def expected_variable_size(self):  # :944-954
        return self.N if self.formulation == EXPLICIT else self.rot_and_range_size

    def set_formulation(self, formulation):  # CORA_problem.h:338
        if formulation not in (EXPLICIT, IMPLICIT):
            raise ValueError("Unknown formulation")
        self.formulation = formulation
        if self.Q is not None and formulation == IMPLICIT:
            self._fill_implicit()

    def _fill_implicit(self):
        k = self.rot_and_range_size
        nt = self.n + self.l
        Q = self.Q.tocsr()
        self._Qmain = Q[:k, :k].tocsr()
        self._Tred = Q[:k, k:k + nt - 1].tocsr()
        self._Ltrans = spla.splu(Q[k:k + nt - 1, k:k + nt - 1].tocsc())

    def translation_explicit_solution(self, Y):  # :1168-1197
        if Y.shape[0] != self.rot_and_range_size:
            raise ValueError("expected %d rows" % self.rot_and_range_size)
        X = np.zeros((self.N, Y.shape[1]))
        X[: Y.shape[0]] = Y
        X[Y.shape[0]: self.N - 1] = -self._Ltrans.solve(np.ascontiguousarray(self._Tred.T @ Y))
        return X

    # -- preconditioner (CORA_problem.cpp:512-623) ---------------------------
    def _update_preconditioner(self):
        Q = self.Q
        if self.preconditioner == JACOBI:
            self._jacobi = 1.0 / Q.diagonal()  # :616-618
        elif self.preconditioner == REG_CHOLESKY:
            if self.lambda_reg is None:
                # :556-591.  The reference estimates ||Q||_2 with a random-start
                # LOBPCG to 1e-2; the oracle uses the converged value.
                if Q.shape[0] <= 3:
                    dn = float(np.linalg.eigvalsh(Q.toarray())[-1])
                else:
                    dn = float(spla.eigsh(Q, k=1, which="LA", tol=1e-6,
                                          return_eigenvectors=False)[0])
                self.lambda_reg = dn / (self.reg_chol_max_cond - 1.0)
            N = Q.shape[0]
            M = (Q + self.lambda_reg * sp.identity(N, format="csr")).tocsc()
            # pin_last_translation_ is const true (CORA_problem.h:72, :602-609)
            if getattr(self, "chol_ordering", "colamd") == "pose_major":
                # Same matrix, same solution; only the elimination order differs.  SuperLU's default column ordering
                # takes minutes on the 100k-pose chain (dense landmark columns); ranges first, then the poses in
                # chain order (rotation rows + translation row together), landmarks last is banded + border.
                self._chol = _PermutedLU(M[: N - 1, : N - 1].tocsc(), pose_major_order(self.d, self.n, self.l, self.m)[:-1])
            else:
                self._chol = spla.splu(M[: N - 1, : N - 1].tocsc())
        else:
            raise ValueError("The desired preconditioner is not implemented")

    # -- operators (CORA_problem.cpp:742-938) --------------------------------
    def _check(self, Y):
        if not self.up_to_date:
            raise RuntimeError("The data matrix must be constructed first")
        if Y.shape[0] != self.expected_variable_size:
            raise ValueError("expected matrix of shape (%d, %d) but got %s"
                             % (self.expected_variable_size, Y.shape[1], Y.shape))

    def data_matrix_product(self, Y):  # :742-757
        self._check(Y)
        if self.formulation == EXPLICIT:
            return self.Q @ Y
        P2 = self._Ltrans.solve(np.ascontiguousarray(self._Tred.T @ Y))
        return self._Qmain @ Y - self._Tred @ P2

    def evaluate_objective(self, Y):  # :759-762
        return 0.5 * float(np.sum(Y * self.data_matrix_product(Y)))

    def euclidean_gradient(self, Y):  # :764-770
        return self.data_matrix_product(Y)

    def tangent_space_projection(self, Y, V):  # :782-820
        return tangent_projection(self.d, self.n, self.m, Y, V)

    def riemannian_gradient(self, Y, egrad=None):  # :772-780
        if egrad is None:
            egrad = self.euclidean_gradient(Y)
        return self.tangent_space_projection(Y, egrad)

    def hessvec(self, Y, egrad, Ydot):  # :822-867
        d, n, m = self.d, self.n, self.m
        W = self.data_matrix_product(Ydot)
        r = Y.shape[1]
        dn = d * n
        if n:
            Yb = Y[:dn].reshape(n, d, r)
            Gb = egrad[:dn].reshape(n, d, r)
            Db = Ydot[:dn].reshape(n, d, r)
            P = np.einsum("nir,njr->nij", Yb, Gb)
            S = 0.5 * (P + P.transpose(0, 2, 1))
            W[:dn] -= np.einsum("nij,njr->nir", S, Db).reshape(dn, r)
        if m:
            lam = np.sum(egrad[dn:dn + m] * Y[dn:dn + m], axis=1)
            W[dn:dn + m] -= lam[:, None] * Ydot[dn:dn + m]
        return tangent_projection(d, n, m, Y, W)

    def precondition(self, V):  # :869-903
        if self.preconditioner == JACOBI:
            # (the reference multiplies the N x N diagonal with the (d n + m)-row matrix in the implicit
            #  formulation, a size mismatch: the oracle uses the leading part of the diagonal)
            res = self._jacobi[: V.shape[0], None] * V
        elif self.formulation == IMPLICIT:  # :878-885: lift with zero translations, solve, keep the top rows
            lift = np.zeros((self.N, V.shape[1]))
            lift[: V.shape[0]] = V
            res = self._chol.solve(np.ascontiguousarray(lift[:-1]))[: V.shape[0]]
        else:
            res = np.zeros_like(V)
            res[:-1] = self._chol.solve(np.ascontiguousarray(V[:-1]))  # CORA_preconditioners.cpp:46-83
        if np.isnan(res).any():
            raise RuntimeError("NaNs in preconditioned vector")
        return res

    def project_to_manifold(self, A):  # :905-934
        return project_to_manifold(self.d, self.n, self.m, A)

    def retract(self, Y, V):  # :936-938
        return self.project_to_manifold(Y + V)

    def set_rank(self, r):
        self.rank = r

    def increment_rank(self):
        self.rank += 1

    def random_initial_guess(self, rng):  # :1023-1028 (reference is unseeded)
        return self.project_to_manifold(rng.uniform(-1, 1, size=(self.expected_variable_size, self.rank)))

    # -- certification (CORA_problem.cpp:1030-1166) ---------------------------
    def compute_lambda_blocks(self, Y):  # :1105-1131
        return lambda_blocks(self.d, self.n, self.m, Y, self.data_matrix_product(Y))

    def lambda_from_blocks(self, blocks, size=None):  # :1133-1160
        return lambda_matrix(self.d, self.n, self.m, blocks, self.N if size is None else size)

    def certificate_matrix(self, Y):  # :1162-1166
        return (self.Q - self.lambda_from_blocks(self.compute_lambda_blocks(Y))).tocsr()

    def certify_solution(self, Y, eta, nx, bootstrap, max_iters=500, rng=None):  # :1030-1103
        rng = rng or np.random.default_rng(0)
        sv = np.linalg.svd(Y, compute_uv=False)
        if sv[0] / sv[-1] > 1e6:  # :1039-1049
            return CertResults(True, 0.0, np.zeros(self.N), np.zeros((self.N, nx)), 0)
        S = self.certificate_matrix(Y)
        num = min(max(nx, Y.shape[1] + 2), S.shape[0])  # :1062-1063
        X0 = rng.uniform(-1, 1, size=(S.shape[0], num))
        if bootstrap is not None and bootstrap.size:
            # :1070-1071.  The reference writes all bootstrap columns (out of
            # bounds when the dense n<=100 branch returned N eigenvectors); the
            # oracle keeps the leading `num` columns.
            k = min(bootstrap.shape[1], num)
            X0[:, :k] = bootstrap[:, :k]
        res = fast_verification(S, eta, X0, max_iters)
        while math.isnan(res.theta):  # :1076-1083
            eta *= 2
            res = fast_verification(S, eta, X0, max_iters)
        if not res.is_certified and self.formulation == IMPLICIT:  # :1085-1100
            k = self.rot_and_range_size
            v = res.x[:k] / np.linalg.norm(res.x[:k])
            Lam = self.lambda_from_blocks(self.compute_lambda_blocks(Y), k)
            Sx = self.data_matrix_product(v[:, None])[:, 0] - Lam @ v
            res.x = v
            res.theta = float(v @ Sx)
        return res


# --------------------------------------------------------------------------- #
#  Assembly (vectorised; O(#factors))                                          #
# --------------------------------------------------------------------------- #


def build_submatrices(d, n, l, a):
    nt = n + l
    E = len(a["rp_tau"])
    m = len(a["rg_w"])
    Ep = len(a["rot_kappa"])
    ar = np.arange
    Arange = sp.coo_matrix((np.r_[-np.ones(m), np.ones(m)],
                            (np.r_[ar(m), ar(m)], np.r_[a["rg_a"], a["rg_b"]])), shape=(m, nt)).tocsr()
    OmegaRange = sp.diags(a["rg_w"], format="csr") if m else sp.csr_matrix((0, 0))
    RangeD = sp.diags(a["rg_r"], format="csr") if m else sp.csr_matrix((0, 0))
    Apose = sp.coo_matrix((np.r_[-np.ones(E), np.ones(E)],
                           (np.r_[ar(E), ar(E)], np.r_[a["rp_i"], a["rp_j"]])), shape=(E, nt)).tocsr()
    OmegaPose = sp.diags(a["rp_tau"], format="csr") if E else sp.csr_matrix((0, 0))
    # T: row e, columns d*i..d*i+d-1 = -t   (:208-213)
    rows = np.repeat(ar(E), d)
    cols = (a["rp_i"][:, None] * d + ar(d)[None, :]).ravel()
    T = sp.coo_matrix((-a["rp_t"].ravel(), (rows, cols)), shape=(E, d * n)).tocsr()
    L = rot_conn_laplacian(d, n, a)
    return dict(Arange=Arange, OmegaRange=OmegaRange, RangeDistances=RangeD, Apose=Apose,
                OmegaPose=OmegaPose, T=T, RotConLaplacian=L)


def rot_conn_laplacian(d, n, a):  # fillRotConnLaplacian, CORA_problem.cpp:297-377
    i, j, R, k = a["rot_i"], a["rot_j"], a["rot_R"], a["rot_kappa"]
    Ep = len(k)
    ar = np.arange(d)
    rows, cols, vals = [], [], []
    # diagonal blocks
    for idx in (i, j):
        rr = (idx[:, None] * d + ar[None, :]).ravel()
        rows.append(rr); cols.append(rr); vals.append(np.repeat(k, d))
    # (i,j) block: -kappa R ; (j,i) block: -kappa R^T
    bi = (i[:, None, None] * d + ar[None, :, None]) + np.zeros((1, 1, d), dtype=np.int64)
    bj = (j[:, None, None] * d + ar[None, None, :]) + np.zeros((1, d, 1), dtype=np.int64)
    v = -(k[:, None, None] * R)
    rows += [bi.ravel(), bj.ravel()]
    cols += [bj.ravel(), bi.ravel()]
    vals += [v.ravel(), v.ravel()]
    if Ep == 0:
        return sp.csr_matrix((d * n, d * n))
    return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(d * n, d * n)).tocsr()


def assemble_Q(d, n, l, a) -> sp.csr_matrix:
    """fillDataMatrix, CORA_problem.cpp:625-712, written as direct triplets
    (block formulas: SURVEY Appendix A / CORA_problem.h:147-184).  Exact zeros
    are dropped (the MatrixMarket goldens drop them too)."""
    nt = n + l
    m = len(a["rg_w"])
    dn = d * n
    N = dn + m + nt
    T0 = dn + m  # first translation row
    ar = np.arange(d)
    rows, cols, vals = [], [], []

    def add(r_, c_, v_):
        rows.append(np.asarray(r_).ravel()); cols.append(np.asarray(c_).ravel())
        vals.append(np.asarray(v_, dtype=float).ravel())

    # Q11 = L_rho + T^T Omega T
    L = rot_conn_laplacian(d, n, a).tocoo()
    add(L.row, L.col, L.data)
    i, j, t, tau = a["rp_i"], a["rp_j"], a["rp_t"], a["rp_tau"]
    E = len(tau)
    if E:
        bi = (i[:, None, None] * d + ar[None, :, None]) + np.zeros((1, 1, d), dtype=np.int64)
        bj = (i[:, None, None] * d + ar[None, None, :]) + np.zeros((1, d, 1), dtype=np.int64)
        add(bi, bj, tau[:, None, None] * t[:, :, None] * t[:, None, :])
        # Q13 = T^T Omega A_t: rows of pose i; col t_i: +tau t ; col t_j: -tau t
        ri = i[:, None] * d + ar[None, :]
        ci = np.repeat((T0 + i)[:, None], d, axis=1)
        cj = np.repeat((T0 + j)[:, None], d, axis=1)
        v = tau[:, None] * t
        add(ri, ci, v); add(ci, ri, v)
        add(ri, cj, -v); add(cj, ri, -v)
        # Q33 pose part: Laplacian
        add(T0 + i, T0 + i, tau); add(T0 + j, T0 + j, tau)
        add(T0 + i, T0 + j, -tau); add(T0 + j, T0 + i, -tau)
    if m:
        ka = np.arange(m)
        ra, rb, rho, w = a["rg_a"], a["rg_b"], a["rg_r"], a["rg_w"]
        add(dn + ka, dn + ka, w * rho * rho)           # Q22
        add(dn + ka, T0 + ra, -w * rho); add(T0 + ra, dn + ka, -w * rho)  # Q23
        add(dn + ka, T0 + rb, w * rho); add(T0 + rb, dn + ka, w * rho)
        add(T0 + ra, T0 + ra, w); add(T0 + rb, T0 + rb, w)   # Q33 range part
        add(T0 + ra, T0 + rb, -w); add(T0 + rb, T0 + ra, -w)
    if not rows:
        return sp.csr_matrix((N, N))
    Q = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(N, N)).tocsr()
    Q.sum_duplicates()
    Q.eliminate_zeros()
    Q.sort_indices()
    return Q


# --------------------------------------------------------------------------- #
#  Manifold geometry (src/StiefelProduct.cpp, src/ObliqueManifold.cpp)         #
# --------------------------------------------------------------------------- #


def tangent_projection(d, n, m, Y, V):
    """proj_Y(V): StiefelProduct.h:79-81 + StiefelProduct.cpp:38-55 on pose
    rows, ObliqueManifold.cpp:16-27 on range rows, identity on translations."""
    r = Y.shape[1]
    dn = d * n
    out = np.array(V, dtype=float, copy=True)
    if n:
        Yb = Y[:dn].reshape(n, d, r)
        Vb = out[:dn].reshape(n, d, r)
        P = np.einsum("nir,njr->nij", Yb, Vb)
        S = 0.5 * (P + P.transpose(0, 2, 1))
        out[:dn] = (Vb - np.einsum("nij,njr->nir", S, Yb)).reshape(dn, r)
    if m:
        y = Y[dn:dn + m]
        v = out[dn:dn + m]
        out[dn:dn + m] = v - np.sum(y * v, axis=1)[:, None] * y
    return out


def project_to_manifold(d, n, m, A):
    """Polar factor per pose block (StiefelProduct.cpp:8-36: thin SVD -> U V^T),
    row normalisation per range row (ObliqueManifold.cpp:6-14)."""
    r = A.shape[1]
    dn = d * n
    out = np.array(A, dtype=float, copy=True)
    if n:
        B = out[:dn].reshape(n, d, r)
        U, _, Vt = np.linalg.svd(B, full_matrices=False)
        out[:dn] = (U @ Vt).reshape(dn, r)
    if m:
        v = out[dn:dn + m]
        out[dn:dn + m] = v / np.linalg.norm(v, axis=1)[:, None]
    return out


def lambda_blocks(d, n, m, Y, QY):  # CORA_problem.cpp:1105-1131
    r = Y.shape[1]
    dn = d * n
    Yb = Y[:dn].reshape(n, d, r)
    Gb = QY[:dn].reshape(n, d, r)
    P = np.einsum("nir,njr->nij", Gb, Yb)
    st = 0.5 * (P + P.transpose(0, 2, 1))  # n x d x d
    ob = np.sum(Y[dn:dn + m] * QY[dn:dn + m], axis=1)
    return st, ob


def lambda_matrix(d, n, m, blocks, size):  # CORA_problem.cpp:1133-1160
    st, ob = blocks
    dn = d * n
    ar = np.arange(d)
    base = (np.arange(n) * d)[:, None, None]
    rows = (base + ar[None, :, None] + np.zeros((1, 1, d), dtype=np.int64)).ravel()
    cols = (base + ar[None, None, :] + np.zeros((1, d, 1), dtype=np.int64)).ravel()
    rr = np.r_[rows, dn + np.arange(m)]
    cc = np.r_[cols, dn + np.arange(m)]
    vv = np.r_[st.ravel(), ob]
    return sp.coo_matrix((vv, (rr, cc)), shape=(size, size)).tocsr()


# --------------------------------------------------------------------------- #
#  STPCG (libs/Optimization/.../LinearAlgebra/IterativeSolvers.h:166-426)      #
# --------------------------------------------------------------------------- #


def stpcg(g, H, inner, Delta, max_iterations=1000, kappa_fgr=0.1, theta=0.5,
          P=None, epsilon=1e-8):
    """Steihaug-Toint truncated preconditioned CG.  Returns (s, ||s||_M, iters)."""
    if Delta <= 0:
        raise ValueError("Trust-region radius (Delta) must be a positive real value")
    s = 0 * g
    r = g.copy()
    v = r if P is None else P(r)  # :229-253
    p = -v
    sMp = 0.0
    sM2 = 0.0
    pM2 = inner(r, v)  # :266
    Delta2 = Delta * Delta
    r0 = math.sqrt(inner(r, v))
    target = r0 * min(kappa_fgr, r0 ** theta)  # :278-279
    it = 0
    while it < max_iterations:
        if math.sqrt(inner(r, v)) <= target:  # :290
            break
        Hp = H(p)  # :294
        kappa = inner(p, Hp)
        if math.sqrt(inner(Hp, Hp)) / math.sqrt(inner(p, p)) < epsilon:  # :305-338
            if inner(p, r) < 0:
                p = -p
                sMp = -sMp
            sigma = (-sMp + math.sqrt(sMp * sMp + pM2 * (Delta2 - sM2))) / pM2
            return s + sigma * p, Delta, it
        alpha = inner(r, v) / kappa  # :341
        sM2_next = sM2 + 2 * alpha * sMp + alpha * alpha * pM2
        if kappa <= 0 or sM2_next > Delta2:  # :347-362
            sigma = (-sMp + math.sqrt(sMp * sMp + pM2 * (Delta2 - sM2))) / pM2
            return s + sigma * p, Delta, it
        s = s + alpha * p  # :374
        r = r + alpha * Hp  # :377
        v = r if P is None else P(r)  # :386
        rv = inner(r, v)
        beta = rv / (alpha * kappa)  # :412
        sM2 = sM2_next
        sMp = beta * (sMp + alpha * pM2)
        pM2 = rv + beta * beta * pM2
        p = -v + beta * p  # :420
        it += 1
    return s, math.sqrt(sM2), it


# --------------------------------------------------------------------------- #
#  TNT (libs/Optimization/.../Riemannian/TNT.h:242-689)                        #
# --------------------------------------------------------------------------- #

STATUS = ["Gradient", "PreconditionedGradient", "RelativeDecrease", "Stepsize",
          "TrustRegion", "IterationLimit", "ElapsedTime", "UserFunction"]


@dataclass
class TNTParams:
    # SmoothOptimizerParams / TNTParams defaults (TNT.h:76-130, Base/Concepts.h)
    Delta0: float = 1.0
    eta1: float = 0.05
    eta2: float = 0.9
    alpha1: float = 0.25
    alpha2: float = 2.5
    max_TPCG_iterations: int = 1000
    kappa_fgr: float = 0.1
    theta: float = 0.5
    preconditioned_gradient_tolerance: float = 1e-6
    Delta_tolerance: float = 1e-6
    gradient_tolerance: float = 1e-6
    relative_decrease_tolerance: float = 1e-6
    stepsize_tolerance: float = 1e-6
    max_iterations: int = 1000
    max_computation_time: float = float("inf")


def cora_tnt_params(**kw) -> TNTParams:
    """The values solveCORA sets (src/CORA.cpp:95-109), time cap lifted."""
    p = TNTParams(Delta0=5, alpha2=3.0, max_TPCG_iterations=80, max_iterations=250,
                  preconditioned_gradient_tolerance=1e-6, gradient_tolerance=1e-6,
                  theta=0.8, Delta_tolerance=1e-5, relative_decrease_tolerance=1e-6,
                  stepsize_tolerance=1e-6)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@dataclass
class TNTResult:
    x: np.ndarray = None
    f: float = 0.0
    gradfx_norm: float = 0.0
    preconditioned_grad_f_x_norm: float = 0.0
    status: str = "IterationLimit"
    elapsed_time: float = 0.0
    objective_values: list = field(default_factory=list)
    gradient_norms: list = field(default_factory=list)
    preconditioned_gradient_norms: list = field(default_factory=list)
    inner_iterations: list = field(default_factory=list)
    update_step_norms: list = field(default_factory=list)
    update_step_M_norms: list = field(default_factory=list)
    gain_ratios: list = field(default_factory=list)
    trust_region_radius: list = field(default_factory=list)
    time: list = field(default_factory=list)


def tnt(f: Callable, QM: Callable, metric: Callable, retract: Callable, x0,
        precon: Optional[Callable] = None, params: TNTParams = TNTParams()) -> TNTResult:
    """Riemannian truncated-Newton trust region.  ``QM(x)`` returns
    ``(grad, hess)`` with ``hess(v)`` the Hessian-vector product at x."""
    sqrt_eps = math.sqrt(np.finfo(float).eps)
    res = TNTResult()
    x = x0
    fx = f(x)
    grad, Hess = QM(x)
    gnorm = math.sqrt(metric(grad, grad))
    if precon is not None:  # :383-392
        pg = precon(x, grad)
        pgnorm = math.sqrt(metric(pg, pg))
    else:
        pgnorm = gnorm
    Delta = params.Delta0
    t0 = time.perf_counter()
    for _ in range(params.max_iterations):
        el = time.perf_counter() - t0
        if el > params.max_computation_time:
            res.status = "ElapsedTime"
            break
        res.time.append(el); res.objective_values.append(fx)
        res.gradient_norms.append(gnorm); res.preconditioned_gradient_norms.append(pgnorm)
        res.trust_region_radius.append(Delta)
        if gnorm < params.gradient_tolerance:  # :474-481
            res.status = "Gradient"
            break
        if pgnorm < params.preconditioned_gradient_tolerance:
            res.status = "PreconditionedGradient"
            break
        P = (lambda v, x=x: precon(x, v)) if precon is not None else None
        h, hM, inner_its = stpcg(grad, Hess, metric, Delta, params.max_TPCG_iterations,
                                 params.kappa_fgr, params.theta, P)  # :489-492
        hnorm = math.sqrt(metric(h, h))
        xp = retract(x, h)  # :505
        fp = f(xp)
        dm = -metric(grad, h) - 0.5 * metric(h, Hess(h))  # :511-512
        df = fx - fp
        rel = df / (sqrt_eps + abs(fx))
        rho = df / dm if dm != 0 else float("nan")
        accepted = (not math.isnan(rho)) and rho > params.eta1  # :532
        res.inner_iterations.append(inner_its); res.update_step_norms.append(hnorm)
        res.update_step_M_norms.append(hM); res.gain_ratios.append(rho)
        if accepted:
            x = xp
            fx = fp
            if rel < params.relative_decrease_tolerance:  # :561-564
                res.status = "RelativeDecrease"
                break
            if hnorm < params.stepsize_tolerance:  # :567-570
                res.status = "Stepsize"
                break
            grad, Hess = QM(x)  # :573
            gnorm = math.sqrt(metric(grad, grad))
            if precon is not None:
                pg = precon(x, grad)
                pgnorm = math.sqrt(metric(pg, pg))
            else:
                pgnorm = gnorm
        if (not math.isnan(rho)) and rho >= params.eta2:  # :590-603
            Delta = max(params.alpha2 * hM, Delta)
        elif math.isnan(rho) or rho < params.eta1:
            Delta = params.alpha1 * hM
            if Delta < params.Delta_tolerance:
                res.status = "TrustRegion"
                break
    res.elapsed_time = time.perf_counter() - t0
    res.x = x; res.f = fx; res.gradfx_norm = gnorm; res.preconditioned_grad_f_x_norm = pgnorm
    res.time.append(res.elapsed_time); res.objective_values.append(fx)
    res.gradient_norms.append(gnorm); res.preconditioned_gradient_norms.append(pgnorm)
    res.trust_region_radius.append(Delta)
    return res


def problem_tnt(problem: Problem, X0, params: Optional[TNTParams] = None) -> TNTResult:
    """TNT wired with the closures solveCORA builds (src/CORA.cpp:52-122)."""
    params = params or cora_tnt_params()
    metric = lambda a, b: float(np.sum(a * b))  # :119-122

    def QM(Y):  # :58-75
        eg = problem.euclidean_gradient(Y)
        grad = problem.riemannian_gradient(Y, eg)
        return grad, (lambda V: problem.hessvec(Y, eg, V))

    precon = lambda Y, V: problem.tangent_space_projection(Y, problem.precondition(V))  # :89-92
    return tnt(problem.evaluate_objective, QM, metric, problem.retract, X0, precon, params)


# --------------------------------------------------------------------------- #
#  LOBPCG (libs/Optimization/.../LinearAlgebra/LOBPCG.h:53-337)                #
# --------------------------------------------------------------------------- #


def rayleigh_ritz(A, B):  # LOBPCG.h:53-62
    D = 1.0 / np.sqrt(np.diag(B))
    w, V = sla.eigh(D[:, None] * A * D[None, :], D[:, None] * B * D[None, :])
    return w, D[:, None] * V


def lobpcg(A: Callable, X0, nev, max_iters, T: Optional[Callable] = None, tau=1e-6,
           user: Optional[Callable] = None, seed=1):
    """Returns (Theta[:nev], X[:, :nev], num_iters, nc).  B = I."""
    m_, nx = X0.shape
    if nev > nx:
        raise ValueError("Block size nx must be >= nev")
    if nx > m_:
        raise ValueError("Block size nx must be <= problem dimension")
    rng = np.random.default_rng(seed)
    Om = rng.standard_normal((m_, nx))
    A2 = np.linalg.norm(A(Om)) / np.linalg.norm(Om)  # :213
    X = X0.copy()
    AX = A(X)
    Theta, C = rayleigh_ritz(X.T @ AX, X.T @ X)  # :222-223
    # NB: the reference updates AX, BX but *not* X here (:226-227); the first
    # search space is spanned by the same columns, so the Ritz pairs agree.
    X = X @ C
    AX = AX @ C
    R = AX - X * Theta[None, :]
    nc = 0
    Pm = None
    it = 1
    while it < max_iters:  # :237
        W = T(R) if T is not None else R
        blocks = [X, W[:, nc:]]
        if it > 1:
            blocks.append(Pm[:, nc:])
        S = np.concatenate(blocks, axis=1)
        AS = A(S)
        Theta, C = rayleigh_ritz(S.T @ AS, S.T @ S)
        X = S @ C[:, :nx]
        AX = A(X)
        R = AX - X * Theta[None, :nx]
        Pm = S[:, nx:] @ C[nx:, :nx]
        rn = np.linalg.norm(R, axis=0)
        tol = tau * (A2 + np.abs(Theta[:nx])) * np.linalg.norm(X, axis=0)
        conv = rn[:nev] <= tol[:nev]
        nc = 0
        while nc < nev and conv[nc]:
            nc += 1
        if user is not None and user(it, Theta[:nx], X, rn, nc):
            break
        if nc == nev:
            break
        it += 1
    return Theta[:nev], X[:, :nev], it, nc


# --------------------------------------------------------------------------- #
#  fast_verification (src/CORA_utils.cpp:17-186)                               #
# --------------------------------------------------------------------------- #


@dataclass
class CertResults:  # include/CORA/CORA_types.h:58-64
    is_certified: bool
    theta: float
    x: np.ndarray
    all_eigvecs: np.ndarray
    num_iters: int


def is_positive_definite(M) -> bool:
    """Stand-in for the CholmodSupernodalLLT success test (CORA_utils.cpp:36-51):
    dense Cholesky for small n, otherwise a symmetric-mode sparse LU without
    off-diagonal pivoting whose pivot signs give the inertia."""
    n = M.shape[0]
    if n <= 3000:
        try:
            np.linalg.cholesky(M.toarray() if sp.issparse(M) else M)
            return True
        except np.linalg.LinAlgError:
            return False
    try:
        lu = spla.splu(sp.csc_matrix(M), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                       options=dict(SymmetricMode=True))
    except RuntimeError:
        return False
    if not np.array_equal(lu.perm_r, lu.perm_c):
        return False
    return bool(np.all(lu.U.diagonal() > 0))


def fast_verification(S, eta, X0, max_iters=500) -> CertResults:
    S = sp.csr_matrix(S)
    n = S.shape[0]
    if np.isscalar(X0) or isinstance(X0, int):  # CORA_utils.h overload: random block
        X0 = np.random.default_rng(0).uniform(-1, 1, size=(n, int(X0)))
    if X0.ndim == 1:
        X0 = X0[:, None]
    M = (S + eta * sp.identity(n, format="csr")).tocsr()
    PSD = is_positive_definite(M)
    theta = 0.0
    num_iters = 0
    X = np.zeros((n, 0))
    if PSD:
        x = np.zeros(n)
    else:
        if n <= 100:  # :63-74
            w, V = np.linalg.eigh(S.toarray())
            return CertResults(False, float(w[0]), V[:, 0].copy(), V, 0)
        Mop = lambda Z: M @ Z
        stop = lambda i, Th, Xc, rn, nc: float(Xc[:, 0] @ (S @ Xc[:, 0])) < -eta / 2  # :90-99
        n1 = int(0.01 * max_iters)
        _, X, num_iters, _ = lobpcg(Mop, X0, 1, n1, None, 0.0, stop)
        x = X[:, 0]
        theta = float(x @ (S @ x))
        if theta >= -eta / 2:  # :127-176
            # The reference preconditions with SYM-ILDL (inertia-corrected
            # incomplete LDL^T).  PARITY UNPINNED: the oracle uses the exact
            # inverse of (S + (|lambda_shift|) I) shifted to be positive definite.
            shift = max(eta, 1e-3 * abs(spla.norm(S, 1)))
            lu = spla.splu((S + shift * sp.identity(n)).tocsc())
            Tm = lambda Rm: lu.solve(np.ascontiguousarray(Rm))
            n2 = int((1.0 - 0.01) * max_iters)
            _, X, it2, _ = lobpcg(Mop, X0, 1, n2, Tm, 0.0, stop)
            x = X[:, 0]
            theta = float(x @ (S @ x))
            num_iters = it2 + int(0.01 * it2)
    return CertResults(PSD, theta, x, X, num_iters)


# --------------------------------------------------------------------------- #
#  Staircase driver (src/CORA.cpp)                                             #
# --------------------------------------------------------------------------- #


def saddle_escape(problem: Problem, Y, theta, v, gtol=1e-4, pgtol=1e-4):  # CORA.cpp:245-350
    r = problem.rank
    if r != Y.shape[1] + 1:
        raise RuntimeError("Relaxation rank should be one greater than the number of columns in Y")
    Ya = np.zeros((Y.shape[0], r)); Ya[:, : r - 1] = Y
    FY = problem.evaluate_objective(Ya)
    Yd = np.zeros((Y.shape[0], r)); Yd[:, -1] = v
    amin = 1e-6
    alpha = max(16 * amin, 100 * gtol / abs(theta))
    alphas, fvals = [], []
    while alpha >= amin:
        Yt = problem.retract(Ya, alpha * Yd)
        Ft = problem.evaluate_objective(Yt)
        g = problem.riemannian_gradient(Yt)
        gn = np.linalg.norm(g)
        pgn = np.linalg.norm(problem.tangent_space_projection(Yt, problem.precondition(g)))
        alphas.append(alpha); fvals.append(Ft)
        if Ft < FY and gn > gtol and pgn > pgtol:
            return Yt
        alpha /= 2
    k = int(np.argmin(fvals))
    if fvals[k] < FY:
        return problem.retract(Ya, alphas[k] * Yd)
    return Ya


def project_to_SOd(M):  # CORA_utils.cpp:188-202
    U, _, Vt = np.linalg.svd(M)
    if np.linalg.det(U) * np.linalg.det(Vt) > 0:
        return U @ Vt
    U = U.copy(); U[:, -1] *= -1
    return U @ Vt


def project_solution(problem: Problem, Y):  # CORA.cpp:352-441
    d, n, m = problem.d, problem.n, problem.m
    U, s, _ = np.linalg.svd(Y, full_matrices=False)
    Yd = U[:, :d] * s[:d][None, :]
    dets = np.linalg.det(Yd[: d * n].reshape(n, d, d)) if n else np.zeros(0)
    ng0 = int(np.sum(dets > 0))
    if n > 0 and ng0 < n // 2:  # integer division as in the reference (:403)
        refl = np.eye(d); refl[-1, -1] = -1
        Yd = Yd @ refl
    for i in range(n):
        Yd[i * d:(i + 1) * d] = project_to_SOd(Yd[i * d:(i + 1) * d])
    dn = d * n
    if m:
        Yd[dn:dn + m] /= np.linalg.norm(Yd[dn:dn + m], axis=1)[:, None]
    return Yd


@dataclass
class CoraResult:
    result: TNTResult
    certified: bool
    theta: float
    eta: float
    final_rank: int
    lifted_f: float = float("nan")
    lifted_rank: int = 0
    total_cg_iterations: int = 0
    stages: list = field(default_factory=list)


def solve_cora(problem: Problem, x0, max_rank=20, params: Optional[TNTParams] = None,
               verbose=False, tnt_fn=None) -> CoraResult:
    """solveCORA, src/CORA.cpp:26-243.  tnt_fn(problem, X, params) -> TNTResult replaces the NumPy TNT (the C++
    restatement oracle/cpu_ref.cpp runs the same algorithm two orders of magnitude faster: oracle/cpu_solve.py)."""
    params = params or cora_tnt_params()
    problem_tnt_ = tnt_fn or problem_tnt
    if x0.shape[0] != problem.N:
        raise ValueError("solveCora::Explicit: bad x0 shape")
    X = problem.project_to_manifold(x0)
    boot = None
    cert = None
    first = True
    total = 0
    stages = []
    lifted_f, lifted_rank = float("nan"), 0
    while problem.rank <= max_rank:
        res = problem_tnt_(problem, X, params)
        total += int(sum(res.inner_iterations))
        eta = min(max(res.f * 5e-6, 1e-7), 1e-1)  # :154
        boot = res.x if first else cert.all_eigvecs
        first = False
        cert = problem.certify_solution(res.x, eta, 10, boot)
        stages.append(dict(rank=problem.rank, f=res.f, grad=res.gradfx_norm, status=res.status,
                           outer=len(res.inner_iterations), cg=int(sum(res.inner_iterations)),
                           certified=cert.is_certified, theta=cert.theta, eta=eta))
        if verbose:
            print(stages[-1])
        if math.isnan(cert.theta):
            raise RuntimeError("Theta is NaN")
        lifted_f, lifted_rank = res.f, problem.rank
        if cert.is_certified:
            X = res.x
            break
        problem.increment_rank()
        X = saddle_escape(problem, res.x, cert.theta, cert.x, 1e-4, 1e-4)
    if X.shape[1] > problem.d:  # :200-233
        X = project_solution(problem, X)
        problem.set_rank(problem.d)
        res = problem_tnt_(problem, X, params)
        total += int(sum(res.inner_iterations))
        eta = min(max(res.f * 5e-6, 1e-7), 1e-1)
        cert = problem.certify_solution(res.x, eta, 10, boot)
        stages.append(dict(rank=problem.rank, f=res.f, grad=res.gradfx_norm, status=res.status,
                           outer=len(res.inner_iterations), cg=int(sum(res.inner_iterations)),
                           certified=cert.is_certified, theta=cert.theta, eta=eta, refine=True))
        if verbose:
            print(stages[-1])
    return CoraResult(res, cert.is_certified, cert.theta, eta, problem.rank, lifted_f,
                      lifted_rank, total, stages)


# --------------------------------------------------------------------------- #
#  PyFG text parser (src/pyfg_text_parser.cpp:112-401)                         #
# --------------------------------------------------------------------------- #


def _from_angle(th):  # :323-328
    c, s = math.cos(th), math.sin(th)
    return np.array([[c, -s], [s, c]])


def _from_quat(qx, qy, qz, qw):  # :330-338 (Eigen toRotationMatrix, no normalisation)
    tx, ty, tz = 2 * qx, 2 * qy, 2 * qz
    twx, twy, twz = tx * qw, ty * qw, tz * qw
    txx, txy, txz = tx * qx, ty * qx, tz * qx
    tyy, tyz, tzz = ty * qy, tz * qy, tz * qz
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def _read_symmetric(tok, dim):  # :385-401
    cov = np.zeros((dim, dim))
    k = 0
    for i in range(dim):
        for j in range(i, dim):
            cov[i, j] = cov[j, i] = float(tok[k]); k += 1
    return cov, k


_TYPES = {"VERTEX_SE2": 2, "VERTEX_SE3:QUAT": 3, "VERTEX_XY": 2, "VERTEX_XYZ": 3}


def parse_pyfg(path_or_text: str, from_text=False) -> Problem:
    if from_text:
        lines = path_or_text.splitlines()
    else:
        with open(path_or_text) as fh:
            lines = fh.read().splitlines()
    if not lines:
        raise RuntimeError("Could not read item type from line ")
    first = lines[0].split()
    if not first or first[0] not in _TYPES:
        raise RuntimeError("Could not determine dimension from first line " + lines[0])
    d = _TYPES[first[0]]
    prob = Problem(d, d, REG_CHOLESKY)  # :116-120
    for line in lines:
        tok = line.split()
        if not tok:
            raise RuntimeError("Could not read item type from line " + line)
        kind = tok[0]
        F = lambda a: np.array([float(x) for x in a])
        if kind in ("VERTEX_SE2", "VERTEX_SE3:QUAT"):
            prob.add_pose(tok[2])
        elif kind in ("VERTEX_XY", "VERTEX_XYZ"):
            prob.add_landmark(tok[1])
        elif kind == "EDGE_SE2":
            cov, _ = _read_symmetric(tok[7:], 3)
            prob.add_rel_pose(RelPose(tok[2], tok[3], _from_angle(float(tok[6])), F(tok[4:6]), cov))
        elif kind == "EDGE_SE3:QUAT":
            cov, _ = _read_symmetric(tok[11:], 6)
            prob.add_rel_pose(RelPose(tok[2], tok[3], _from_quat(*[float(x) for x in tok[7:11]]),
                                      F(tok[4:7]), cov))
        elif kind == "EDGE_SE2_XY":
            cov, _ = _read_symmetric(tok[6:], 2)
            prob.add_rel_pose_landmark(RelPoseLandmark(tok[2], tok[3], F(tok[4:6]), cov))
        elif kind == "EDGE_SE3_XYZ":
            cov, _ = _read_symmetric(tok[7:], 3)
            prob.add_rel_pose_landmark(RelPoseLandmark(tok[2], tok[3], F(tok[4:7]), cov))
        elif kind == "EDGE_RANGE":
            prob.add_range(Range(tok[2], tok[3], float(tok[4]), float(tok[5])))
        elif kind == "VERTEX_SE2:PRIOR":
            cov, _ = _read_symmetric(tok[6:], 3)
            prob.add_pose_prior(tok[2], _from_angle(float(tok[5])), F(tok[3:5]), cov)
        elif kind == "VERTEX_SE3:QUAT:PRIOR":
            cov, _ = _read_symmetric(tok[10:], 6)
            prob.add_pose_prior(tok[2], _from_quat(*[float(x) for x in tok[6:10]]), F(tok[3:6]), cov)
        elif kind == "VERTEX_XY:PRIOR":
            cov, _ = _read_symmetric(tok[5:], 2)
            prob.add_landmark_prior(tok[2], F(tok[3:5]), cov)
        elif kind == "VERTEX_XYZ:PRIOR":
            cov, _ = _read_symmetric(tok[6:], 3)
            prob.add_landmark_prior(tok[2], F(tok[3:6]), cov)
        else:
            raise RuntimeError("Unknown item type " + kind)
    return prob
