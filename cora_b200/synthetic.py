"""Synthetic RA-SLAM pose graphs of the BASELINE.json shapes.

Poses A0..A(n-1) on a smooth random walk (step 1 m, per-step rotation Exp(N(0, 0.1^2 I))),
odometry A_i -> A_{i+1} with noise sigma_t = 0.05 m, sigma_R = 0.01 rad, l landmarks uniform in
the trajectory bounding box, m ranges from a uniformly sampled pose subset to landmark (k mod l)
with sigma_r = 0.3 m.  Returns the flattened measurement arrays the assembly consumes
(same keys as the reference-ordered stacks of src/CORA_problem.cpp:190-294).
"""
from __future__ import annotations

import numpy as np


def _exp_so3(w):
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -w[..., 2], w[..., 1]
    K[..., 1, 0], K[..., 1, 2] = w[..., 2], -w[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -w[..., 1], w[..., 0]
    th2 = th * th
    small = th < 1e-8
    a = np.where(small, 1 - th2 / 6, np.sin(th) / np.where(small, 1, th))
    b = np.where(small, 0.5 - th2 / 24, (1 - np.cos(th)) / np.where(small, 1, th2))
    return np.eye(3) + a * K + b * (K @ K)


def _exp_so2(w):
    c, s = np.cos(w), np.sin(w)
    R = np.zeros(w.shape + (2, 2))
    R[..., 0, 0], R[..., 0, 1], R[..., 1, 0], R[..., 1, 1] = c, -s, s, c
    return R


def make_arrays(n, l, m, d=3, seed=42, sigma_t=0.05, sigma_R=0.01, sigma_r=0.3, loop_closures=None):
    """Returns (arrays, ground_truth) with ground_truth = (R_gt [n,d,d], t_gt [n,d], landmarks [l,d])."""
    rng = np.random.default_rng(seed)
    # relative motions (ground truth): step 1 m along the body x axis, small rotation
    if d == 3:
        dR = _exp_so3(rng.normal(0.0, 0.1, size=(n - 1, 3)))
    else:
        dR = _exp_so2(rng.normal(0.0, 0.1, size=(n - 1,)))
    dt = np.zeros((n - 1, d)); dt[:, 0] = 1.0
    R = np.empty((n, d, d)); t = np.empty((n, d))
    R[0] = np.eye(d); t[0] = 0.0
    for i in range(n - 1):  # sequential composition (setup only)
        t[i + 1] = t[i] + R[i] @ dt[i]
        R[i + 1] = R[i] @ dR[i]
    lo, hi = t.min(axis=0), t.max(axis=0)
    L = rng.uniform(lo, hi, size=(l, d)) if l else np.zeros((0, d))
    # noisy odometry
    if d == 3:
        Rn = dR @ _exp_so3(rng.normal(0.0, sigma_R, size=(n - 1, 3)))
    else:
        Rn = dR @ _exp_so2(rng.normal(0.0, sigma_R, size=(n - 1,)))
    tn = dt + rng.normal(0.0, sigma_t, size=(n - 1, d))
    ii = np.arange(n - 1, dtype=np.int64)
    jj = ii + 1
    tau = np.full(n - 1, 1.0 / sigma_t ** 2)             # d / tr(sigma_t^2 I_d)
    kap = np.full(n - 1, (1.0 / (2 * sigma_R ** 2)) if d == 3 else 1.0 / sigma_R ** 2)
    if loop_closures:
        lc = np.asarray(loop_closures, dtype=np.int64).reshape(-1, 2)
        a, b = lc[:, 0], lc[:, 1]
        Rab = np.einsum("nji,njk->nik", R[a], R[b])
        tab = np.einsum("nji,nj->ni", R[a], t[b] - t[a])
        ii, jj = np.r_[ii, a], np.r_[jj, b]
        Rn = np.concatenate([Rn, Rab]); tn = np.concatenate([tn, tab])
        tau = np.r_[tau, np.full(len(a), 1.0 / sigma_t ** 2)]
        kap = np.r_[kap, np.full(len(a), kap[0])]
    # ranges: pose subset -> landmark (k mod l)
    if m and l:
        pa = rng.choice(n, size=m, replace=(m > n))
        lb = np.arange(m) % l
        # duplicates (same pose, same landmark) are not allowed by the reference
        key = pa * l + lb
        _, first = np.unique(key, return_index=True)
        first.sort()
        pa, lb = pa[first], lb[first]
        rho = np.linalg.norm(t[pa] - L[lb], axis=1) + rng.normal(0.0, sigma_r, size=len(pa))
        rho = np.abs(rho)
        rg_a, rg_b = pa.astype(np.int64), (n + lb).astype(np.int64)
        rg_w = np.full(len(pa), 1.0 / sigma_r ** 2)
    else:
        rg_a = rg_b = np.zeros(0, dtype=np.int64)
        rho = rg_w = np.zeros(0)
    arrays = dict(rp_i=ii, rp_j=jj, rp_t=tn, rp_tau=tau, rot_i=ii.copy(), rot_j=jj.copy(), rot_R=Rn,
                  rot_kappa=kap, rg_a=rg_a, rg_b=rg_b, rg_r=rho, rg_w=rg_w)
    return arrays, (R, t, L)


def ground_truth_matrix(d, n, l, arrays, gt):
    """X_gt (N x d) in the reference row order: rows of pose i are R_i^T (row form: Y_i = R_i),
    range rows are the unit bearings (from the second to the first id), translations follow."""
    R, t, L = gt
    m = len(arrays["rg_w"])
    X = np.zeros((d * n + m + n + l, d))
    X[: d * n] = np.transpose(R, (0, 2, 1)).reshape(d * n, d)
    T = np.concatenate([t, L]) if l else t
    if m:
        diff = T[arrays["rg_a"]] - T[arrays["rg_b"]]  # Q23 row k: -w rho at a, +w rho at b
        X[d * n: d * n + m] = diff / np.linalg.norm(diff, axis=1)[:, None]
    X[d * n + m:] = T
    return X


def odometry_initialization(d, n, l, arrays, rank, seed=0, as_reference=False):
    """x0 (N x rank, reference row order) by composing the odometry chain, as the reference's paper
    experiments initialise (examples/paper_experiments.cpp:426-534): pose i gets R_i^T / t_i of the
    chained odometry, landmarks 10*U[-1,1]^d, range rows the normalised translation differences,
    everything embedded in rank columns and rotated by a random SO(rank) element so that it is
    generically dense.  Single chain A0 -> A1 -> ....

    Sign of the range rows: the data matrix has Q23 = +D Omega_r A_r with A_r = -1 at the first id,
    +1 at the second (src/CORA_problem.cpp:142-145,663), i.e. the cost w*||rho*y + (t_second -
    t_first)||^2 is minimised by y = (t_first - t_second)/rho -- the sign of the reference's X_gt
    fixtures.  paper_experiments.cpp:500-506 initialises with (second - first), the antipode;
    `as_reference=True` reproduces that, the default uses the sign consistent with Q."""
    rng = np.random.default_rng(seed)
    m = len(arrays["rg_w"])
    N = d * n + m + n + l
    ii, jj = np.asarray(arrays["rot_i"]), np.asarray(arrays["rot_j"])
    Rm, tm = np.asarray(arrays["rot_R"]), np.asarray(arrays["rp_t"])
    R = np.empty((n, d, d)); t = np.empty((n, d))
    R[0] = np.eye(d); t[0] = 0.0
    nxt = {int(a): k for k, (a, b) in enumerate(zip(ii, jj)) if b == a + 1}
    for i in range(n - 1):
        k = nxt[i]
        t[i + 1] = t[i] + R[i] @ tm[k]
        R[i + 1] = R[i] @ Rm[k]
    X = np.zeros((N, rank))
    X[: d * n, :d] = np.transpose(R, (0, 2, 1)).reshape(d * n, d)
    T = np.concatenate([t, 10.0 * rng.uniform(-1, 1, size=(l, d))]) if l else t
    X[d * n + m:, :d] = T
    if m:
        diff = T[arrays["rg_b"]] - T[arrays["rg_a"]] if as_reference else T[arrays["rg_a"]] - T[arrays["rg_b"]]
        nrm = np.linalg.norm(diff, axis=1)
        bad = nrm < 1e-5
        diff[bad] = rng.uniform(-1, 1, size=(int(bad.sum()), d))
        X[d * n: d * n + m, :d] = diff / np.linalg.norm(diff, axis=1)[:, None]
    Qr, _ = np.linalg.qr(rng.uniform(-1, 1, size=(rank, rank)))
    if np.linalg.det(Qr) < 0:
        Qr[:, -1] *= -1
    return np.asfortranarray(X @ Qr)


def perturbed_ground_truth(d, n, l, arrays, gt, rank, seed=0, sigma_R=0.05, sigma_t=0.5):
    """Warm start: the ground-truth configuration with every rotation perturbed by Exp(N(0, sigma_R^2)),
    every position by N(0, sigma_t^2), embedded in `rank` columns and rotated by a random SO(rank)
    element (reference row order, N x rank; on the manifold up to rounding).  This is the regime in
    which STPCG runs its full iteration budget (the trust region is not binding), i.e. the CG loop the
    BASELINE metric is about; `odometry_initialization` is the cold start of the reference's paper
    experiments."""
    rng = np.random.default_rng(seed)
    R, t, L = gt
    m = len(arrays["rg_w"])
    N = d * n + m + n + l
    dR = _exp_so3(rng.normal(0.0, sigma_R, size=(n, 3))) if d == 3 else _exp_so2(rng.normal(0.0, sigma_R, size=(n,)))
    Rp = R @ dR
    T = np.concatenate([t + rng.normal(0.0, sigma_t, size=t.shape), L + rng.normal(0.0, sigma_t, size=L.shape)]) if l \
        else t + rng.normal(0.0, sigma_t, size=t.shape)
    X = np.zeros((N, rank))
    X[: d * n, :d] = np.transpose(Rp, (0, 2, 1)).reshape(d * n, d)
    X[d * n + m:, :d] = T
    if m:
        diff = T[arrays["rg_a"]] - T[arrays["rg_b"]]   # sign of Q23, see odometry_initialization
        X[d * n: d * n + m, :d] = diff / np.linalg.norm(diff, axis=1)[:, None]
    Qr, _ = np.linalg.qr(rng.uniform(-1, 1, size=(rank, rank)))
    if np.linalg.det(Qr) < 0:
        Qr[:, -1] *= -1
    return np.asfortranarray(X @ Qr)
