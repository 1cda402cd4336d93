"""ctypes binding of the C-ABI in include/cora_b200.h.

This is the binding the tests and bench.py use; it holds no arithmetic.  There is no
CPU fallback: if the shared library is missing, or no CUDA device is present, the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CORA_B200_LIB") or os.path.join(_HERE, "lib", "libcora_b200.so")  # (override: development builds)

PRECON_NONE, PRECON_JACOBI, PRECON_BLOCK_CHOLESKY, PRECON_REG_CHOLESKY = 0, 1, 2, 3
FORMULATION_EXPLICIT, FORMULATION_IMPLICIT = 0, 1  # include/CORA/CORA_types.h:52-56
TNT_STATUS = ["Gradient", "PreconditionedGradient", "RelativeDecrease", "Stepsize", "TrustRegion",
              "IterationLimit", "ElapsedTime", "UserFunction"]
EINVAL, ERUNTIME, ECUDA, ENOTIMPL = 1, 2, 3, 4


class CoraB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class InvalidArgument(CoraB200Error, ValueError):
    """std::invalid_argument / MatrixShapeException of the reference."""


class NotImplementedInReference(CoraB200Error, NotImplementedError):
    pass


class TntParams(C.Structure):
    _fields_ = [("Delta0", C.c_double), ("eta1", C.c_double), ("eta2", C.c_double),
                ("alpha1", C.c_double), ("alpha2", C.c_double),
                ("max_TPCG_iterations", C.c_int32), ("max_iterations", C.c_int32),
                ("kappa_fgr", C.c_double), ("theta", C.c_double),
                ("preconditioned_gradient_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("relative_decrease_tolerance", C.c_double), ("stepsize_tolerance", C.c_double),
                ("Delta_tolerance", C.c_double), ("max_computation_time", C.c_double),
                ("verbose", C.c_int32), ("reserved", C.c_int32)]


_PD = C.POINTER(C.c_double)


class TntResultC(C.Structure):
    _fields_ = [("f", C.c_double), ("gradfx_norm", C.c_double), ("preconditioned_gradfx_norm", C.c_double),
                ("elapsed_time", C.c_double), ("device_time", C.c_double),
                ("status", C.c_int32), ("num_outer", C.c_int32), ("total_inner", C.c_int64),
                ("kernel_launches", C.c_int64), ("trace_capacity", C.c_int32), ("reserved", C.c_int32),
                ("objective_values", _PD), ("gradient_norms", _PD), ("preconditioned_gradient_norms", _PD),
                ("trust_region_radius", _PD), ("time", _PD), ("update_step_norms", _PD),
                ("update_step_M_norms", _PD), ("gain_ratios", _PD),
                ("inner_iterations", C.POINTER(C.c_int32))]


class StageC(C.Structure):
    _fields_ = [("rank", C.c_int32), ("status", C.c_int32), ("num_outer", C.c_int32), ("certified", C.c_int32),
                ("cg_iterations", C.c_int64), ("f", C.c_double), ("gradfx_norm", C.c_double),
                ("theta", C.c_double), ("eta", C.c_double), ("tnt_seconds", C.c_double),
                ("cert_seconds", C.c_double), ("cert_branch", C.c_int32), ("reserved", C.c_int32)]


CERT_BRANCH = ["none", "sv_ratio", "psd", "eigenpair", "inconclusive"]


class SolveResultC(C.Structure):
    _fields_ = [("f", C.c_double), ("lifted_f", C.c_double), ("final_rank", C.c_int32),
                ("lifted_rank", C.c_int32), ("certified", C.c_int32), ("num_stages", C.c_int32),
                ("total_cg_iterations", C.c_int64), ("seconds", C.c_double),
                ("stage_capacity", C.c_int32), ("refined_certified", C.c_int32), ("stages", C.POINTER(StageC))]


@dataclass
class TntResult:
    """Mirror of Optimization::Riemannian::TNTResult (TNT.h:168-194)."""
    x: Optional[np.ndarray] = None
    f: float = 0.0
    gradfx_norm: float = 0.0
    preconditioned_grad_f_x_norm: float = 0.0
    status: str = "IterationLimit"
    elapsed_time: float = 0.0
    device_time: float = 0.0
    kernel_launches: int = 0
    objective_values: List[float] = field(default_factory=list)
    gradient_norms: List[float] = field(default_factory=list)
    preconditioned_gradient_norms: List[float] = field(default_factory=list)
    trust_region_radius: List[float] = field(default_factory=list)
    time: List[float] = field(default_factory=list)
    inner_iterations: List[int] = field(default_factory=list)
    update_step_norms: List[float] = field(default_factory=list)
    update_step_M_norms: List[float] = field(default_factory=list)
    gain_ratios: List[float] = field(default_factory=list)


@dataclass
class CertResults:
    """Mirror of CORA::CertResults (include/CORA/CORA_types.h:58-64)."""
    is_certified: bool
    theta: float
    x: np.ndarray
    all_eigvecs: np.ndarray
    num_iters: int


_lib = None

# every symbol include/cora_b200.h declares (checked by the CPU test-suite)
SYMBOLS = [
    "cora_b200_last_error", "cora_b200_version", "cora_b200_device_count", "cora_b200_create",
    "cora_b200_destroy", "cora_b200_size", "cora_b200_set_preconditioner", "cora_b200_get_reg_lambda",
    "cora_b200_set_reg_lambda", "cora_b200_data_matrix_product", "cora_b200_objective", "cora_b200_egrad",
    "cora_b200_rgrad", "cora_b200_hessvec", "cora_b200_tangent_proj", "cora_b200_precondition",
    "cora_b200_retract", "cora_b200_project", "cora_b200_lambda_blocks", "cora_b200_certificate_product",
    "cora_b200_tnt_default_params", "cora_b200_tnt", "cora_b200_set_iterate", "cora_b200_get_iterate",
    "cora_b200_tnt_resident", "cora_b200_spmm_resident", "cora_b200_certify", "cora_b200_saddle_escape",
    "cora_b200_project_solution", "cora_b200_solve", "cora_b200_gather_best", "cora_b200_layout_roundtrip",
    "cora_b200_strip_layout_roundtrip", "cora_b200_effective_preconditioner", "cora_b200_last_cert_branch", "cora_b200_phase_profile_ctas", "cora_b200_gather_best_resident",
    "cora_b200_odometry_initialization", "cora_b200_save_solution", "cora_b200_debug_min_eigenpair", "cora_b200_psd_test",
    "cora_b200_assemble", "cora_b200_snapshot_iterate", "cora_b200_restore_iterate", "cora_b200_profile_hessvec",
    "cora_b200_profile_read", "cora_b200_debug_chain_host", "cora_b200_debug_factor_stats",
    "cora_b200_set_formulation", "cora_b200_variable_rows", "cora_b200_translation_explicit_solution",
    "cora_b200_device_vectors", "cora_b200_row_order",
    "cora_b200_peer_create", "cora_b200_peer_connect", "cora_b200_peer_product", "cora_b200_peer_destroy", "cora_b200_phase_profile", "cora_b200_get_work_vector",
    "cora_b200_pyfg_parse", "cora_b200_pyfg_sizes", "cora_b200_pyfg_arrays", "cora_b200_pyfg_free",
    "cora_b200_select_best", "cora_b200_nccl_unique_id", "cora_b200_nccl_init", "cora_b200_nccl_destroy",
]


def load():
    """Load libcora_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CoraB200Error(ECUDA, "libcora_b200.so is not built (%s): run `python cora_b200/build.py`; "
                            "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.cora_b200_last_error.restype = C.c_char_p
    for s in SYMBOLS:
        getattr(lib, s)
    _lib = lib
    return lib


def _check(code):
    if code == 0:
        return
    msg = load().cora_b200_last_error().decode()
    if code == EINVAL:
        raise InvalidArgument(code, msg)
    if code == ENOTIMPL:
        raise NotImplementedInReference(code, msg)
    raise CoraB200Error(code, msg)


def select_best(f, certified) -> int:
    """The winner rule of cora_b200_gather_best (pure host function)."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    c = np.ascontiguousarray(certified, dtype=np.int32)
    w = C.c_int(0)
    _check(load().cora_b200_select_best(C.c_int(len(f)), _p(f), c.ctypes.data_as(C.POINTER(C.c_int)), C.byref(w)))
    return w.value


_nccl_preloaded = False


def _preload_nccl():
    """The library resolves NCCL with dlopen("libnccl.so.2") at first use.  In a Python process that also imports
    torch, the copy torch was built against (site-packages/nvidia/nccl) must be the one behind that SONAME: the
    dynamic loader shares one libnccl.so.2 per process, and an older system copy loaded first breaks a later
    `import torch` (undefined symbol ncclDevCommCreate).  Load torch's copy first when it exists."""
    global _nccl_preloaded
    if _nccl_preloaded or os.environ.get("CORA_B200_NCCL_LIB"):
        return
    _nccl_preloaded = True
    try:
        import glob
        import importlib.util
        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            for path in sorted(glob.glob(os.path.join(base, "nccl", "lib", "libnccl.so*"))):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass  # fall back to the loader's search path


def nccl_unique_id() -> bytes:
    _preload_nccl()
    buf = C.create_string_buffer(128)
    _check(load().cora_b200_nccl_unique_id(buf))
    return buf.raw


class NcclComm:
    """ncclComm_t created from a unique id (cora_b200_nccl_init)."""

    def __init__(self, device, world_size, rank, unique_id: bytes):
        _preload_nccl()
        self._c = C.c_void_p()
        _check(load().cora_b200_nccl_init(C.byref(self._c), C.c_int(device), C.c_int(world_size), C.c_int(rank),
                                          C.c_char_p(unique_id)))

    def close(self):
        if self._c:
            load().cora_b200_nccl_destroy(self._c)
            self._c = C.c_void_p()


class PeerProduct:
    """Row-partitioned product over peer-mapped memory (cora_b200_peer_*): create -> exchange `handles` -> connect ->
    product(reps) -> close."""

    def __init__(self, handle, r, n_landmark_rows):
        self._lib = load()
        self._p = C.c_void_p()
        buf = C.create_string_buffer(192)
        _check(self._lib.cora_b200_peer_create(handle._h, C.c_int(r), C.c_int(n_landmark_rows), C.byref(self._p), buf))
        self.handles = buf.raw

    def connect(self, world, rank, all_handles: bytes, ghost_peer, ghost_src, ghost_dst, lm_rows):
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        gp, gs, gd, lm = i32(ghost_peer), i32(ghost_src), i32(ghost_dst), i32(lm_rows)
        P = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        assert len(all_handles) == 192 * world
        _check(self._lib.cora_b200_peer_connect(self._p, C.c_int(world), C.c_int(rank), C.c_char_p(all_handles),
                                                C.c_int(len(gp)), P(gp), P(gs), P(gd), P(lm)))

    def product(self, reps=1):
        ms = C.c_float()
        _check(self._lib.cora_b200_peer_product(self._p, C.c_int(reps), C.byref(ms)))
        return ms.value

    def close(self):
        if self._p:
            self._lib.cora_b200_peer_destroy(self._p)
            self._p = C.c_void_p()


def device_count() -> int:
    n = C.c_int(0)
    _check(load().cora_b200_device_count(C.byref(n)))
    return n.value


def default_tnt_params(**kw) -> TntParams:
    p = TntParams()
    _check(load().cora_b200_tnt_default_params(C.byref(p)))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _f64(a, shape=None):
    """Column-major float64 view/copy (the buffer of an Eigen::MatrixXd)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    a = np.asfortranarray(a)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise InvalidArgument(EINVAL, "expected matrix of shape %s but got %s" % (tuple(shape), tuple(a.shape)))
    return a


def _p(a):
    return a.ctypes.data_as(_PD)


def layout_roundtrip(d, n, m, nt, Q, strips=False):
    """Test hook: CSR -> internal layout (-> strip layout of the streaming kernels) -> CSR on the host (no GPU)."""
    import scipy.sparse as sp
    Q = sp.csr_matrix(Q)
    rp = np.ascontiguousarray(Q.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(Q.indices, dtype=np.int32)
    va = np.ascontiguousarray(Q.data, dtype=np.float64)
    nnz = int(Q.nnz)
    orp = np.zeros(Q.shape[0] + 1, dtype=np.int32)
    oci = np.zeros(max(nnz, 1), dtype=np.int32)
    ova = np.zeros(max(nnz, 1), dtype=np.float64)
    stats = np.zeros(8, dtype=np.int64)
    i32 = C.POINTER(C.c_int32)
    fn = load().cora_b200_strip_layout_roundtrip if strips else load().cora_b200_layout_roundtrip
    _check(fn(
        C.c_int(d), C.c_int(n), C.c_int(m), C.c_int(nt), rp.ctypes.data_as(i32), ci.ctypes.data_as(i32),
        _p(va), C.c_int64(nnz), orp.ctypes.data_as(i32), oci.ctypes.data_as(i32), _p(ova),
        stats.ctypes.data_as(C.POINTER(C.c_int64))))
    k = int(stats[7])
    out = sp.csr_matrix((ova[:k], oci[:k], orp), shape=Q.shape)
    names = ["num_tiles", "max_slots", "nnz_block", "nnz_spill", "nnz_hub", "num_hub_groups",
             "block_values_stored", "nnz_out"]
    return out, dict(zip(names, (int(x) for x in stats)))


def debug_chain_host(d, n, m, nt, Q, shift, pin_last=True, V=None):
    """Test hook: host execution of the chain Cholesky.  Returns (pos_def, M^-1 V or None)."""
    import scipy.sparse as sp
    Q = sp.csr_matrix(Q)
    rp = np.ascontiguousarray(Q.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(Q.indices, dtype=np.int32)
    va = np.ascontiguousarray(Q.data, dtype=np.float64)
    i32 = C.POINTER(C.c_int32)
    pd = C.c_int(0)
    out = None
    r = 0
    if V is not None:
        V = _f64(V)
        r = V.shape[1]
        out = np.zeros_like(V, order="F")
    _check(load().cora_b200_debug_chain_host(C.c_int(d), C.c_int(n), C.c_int(m), C.c_int(nt),
                                             rp.ctypes.data_as(i32), ci.ctypes.data_as(i32), _p(va),
                                             C.c_int64(Q.nnz), C.c_double(shift), C.c_int(int(pin_last)),
                                             C.c_int(r), _p(V) if V is not None else None,
                                             _p(out) if out is not None else None, C.byref(pd)))
    return bool(pd.value), out


def debug_factor_stats(d, n, m, nt, Q):
    """Test hook: structure of the pose-system factorisation (chain levels or general sparse block Cholesky)."""
    import scipy.sparse as sp
    Q = sp.csr_matrix(Q)
    rp = np.ascontiguousarray(Q.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(Q.indices, dtype=np.int32)
    va = np.ascontiguousarray(Q.data, dtype=np.float64)
    i32 = C.POINTER(C.c_int32)
    st = np.zeros(12, dtype=np.int64)
    _check(load().cora_b200_debug_factor_stats(C.c_int(d), C.c_int(n), C.c_int(m), C.c_int(nt), rp.ctypes.data_as(i32),
                                               ci.ctypes.data_as(i32), _p(va), C.c_int64(Q.nnz),
                                               st.ctypes.data_as(C.POINTER(C.c_int64))))
    names = ("chain", "couplings", "l_blocks", "etree_height", "clusters", "levels", "max_column", "poses",
             "cluster_inverse_blocks", "max_cluster_row_blocks", "longest_row", "rows_over_64")
    return dict(zip(names, (int(x) for x in st)))


def parse_pyfg(path_or_text, from_text=False):
    """PyFG file (or text) -> (d, n_poses, n_landmarks, measurement stacks) through the C++ parser."""
    lib = load()
    g = C.c_void_p()
    _check(lib.cora_b200_pyfg_parse(C.c_char_p(path_or_text.encode()), C.c_int(int(from_text)), C.byref(g)))
    try:
        d, n, l = C.c_int(), C.c_int(), C.c_int()
        E, Ep, m = C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib.cora_b200_pyfg_sizes(g, C.byref(d), C.byref(n), C.byref(l), C.byref(E), C.byref(Ep), C.byref(m)))
        d, n, l, E, Ep, m = d.value, n.value, l.value, E.value, Ep.value, m.value
        i64, f64 = np.int64, np.float64
        A = dict(rp_i=np.zeros(E, i64), rp_j=np.zeros(E, i64), rp_t=np.zeros((E, d), f64), rp_tau=np.zeros(E, f64),
                 rot_i=np.zeros(Ep, i64), rot_j=np.zeros(Ep, i64), rot_R=np.zeros((Ep, d, d), f64),
                 rot_kappa=np.zeros(Ep, f64), rg_a=np.zeros(m, i64), rg_b=np.zeros(m, i64), rg_r=np.zeros(m, f64),
                 rg_w=np.zeros(m, f64))
        pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        _check(lib.cora_b200_pyfg_arrays(g, pi(A["rp_i"]), pi(A["rp_j"]), _p(A["rp_t"]), _p(A["rp_tau"]),
                                         pi(A["rot_i"]), pi(A["rot_j"]), _p(A["rot_R"]), _p(A["rot_kappa"]),
                                         pi(A["rg_a"]), pi(A["rg_b"]), _p(A["rg_r"]), _p(A["rg_w"])))
        return d, n, l, A
    finally:
        lib.cora_b200_pyfg_free(g)


def odometry_initialization(d, n, l, arrays, rank, seed=0, reference_sign=False):
    """getOdomInitialization (examples/paper_experiments.cpp:426-534) from measurement stacks; N x rank, F order."""
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    rp_i, rp_j, rp_t = i64(arrays["rp_i"]), i64(arrays["rp_j"]), f64(arrays["rp_t"])
    rot_i, rot_j, rot_R = i64(arrays["rot_i"]), i64(arrays["rot_j"]), f64(arrays["rot_R"])
    rg_a, rg_b = i64(arrays["rg_a"]), i64(arrays["rg_b"])
    m = len(rg_a)
    N = d * n + m + n + l
    out = np.zeros((N, rank), order="F")
    pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    _check(load().cora_b200_odometry_initialization(
        C.c_int(d), C.c_int(n), C.c_int(l), C.c_int64(len(rp_i)), pi(rp_i), pi(rp_j), _p(rp_t),
        C.c_int64(len(rot_i)), pi(rot_i), pi(rot_j), _p(rot_R), C.c_int64(m), pi(rg_a), pi(rg_b), C.c_int(rank),
        C.c_uint64(seed), C.c_int(int(reference_sign)), _p(out)))
    return out


def save_solution(path, X, d, n, m, nt, fmt="tum", first=0, count=None):
    """saveSolnToTum / saveSolnToG20 (src/CORA_utils.cpp:234-350) of a rounded N x d solution."""
    X = np.asfortranarray(np.asarray(X, dtype=np.float64))
    count = n - first if count is None else count
    _check(load().cora_b200_save_solution(str(path).encode(), C.c_int({"tum": 0, "g2o": 1}[fmt]), C.c_int(d), C.c_int(n),
                                          C.c_int(m), C.c_int(nt), _p(X), C.c_int64(first), C.c_int64(count)))


def assemble(d, n, l, arrays):
    """Data matrix Q (scipy CSR, reference row order) from flattened measurement stacks."""
    import scipy.sparse as sp
    lib = load()
    i64, f64 = np.int64, np.float64
    A = {k: np.ascontiguousarray(arrays[k], dtype=(i64 if k in ("rp_i", "rp_j", "rot_i", "rot_j", "rg_a", "rg_b") else f64))
         for k in ("rp_i", "rp_j", "rp_t", "rp_tau", "rot_i", "rot_j", "rot_R", "rot_kappa", "rg_a", "rg_b", "rg_r", "rg_w")}
    E, Ep, m = len(A["rp_tau"]), len(A["rot_kappa"]), len(A["rg_w"])
    pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    nnz = C.c_int64(0)
    args = [C.c_int(d), C.c_int(n), C.c_int(l), C.c_int64(E), pi(A["rp_i"]), pi(A["rp_j"]), _p(A["rp_t"]),
            _p(A["rp_tau"]), C.c_int64(Ep), pi(A["rot_i"]), pi(A["rot_j"]), _p(A["rot_R"]), _p(A["rot_kappa"]),
            C.c_int64(m), pi(A["rg_a"]), pi(A["rg_b"]), _p(A["rg_r"]), _p(A["rg_w"]), C.byref(nnz)]
    _check(lib.cora_b200_assemble(*args, None, None, None))
    N = d * n + m + n + l
    rp = np.zeros(N + 1, dtype=np.int32)
    ci = np.zeros(max(nnz.value, 1), dtype=np.int32)
    va = np.zeros(max(nnz.value, 1), dtype=np.float64)
    i32 = C.POINTER(C.c_int32)
    _check(lib.cora_b200_assemble(*args, rp.ctypes.data_as(i32), ci.ctypes.data_as(i32), _p(va)))
    return sp.csr_matrix((va[: nnz.value], ci[: nnz.value], rp), shape=(N, N))


class Handle:
    """One problem on one GPU: the device side of CORA::Problem after updateProblemData()."""

    def __init__(self, d, n_poses, n_ranges, n_trans, Q, preconditioner=PRECON_JACOBI, device=0,
                 stream=None, reg_chol_max_cond=0.0):
        import scipy.sparse as sp
        lib = load()
        Q = sp.csr_matrix(Q)
        Q.sort_indices()
        N = d * n_poses + n_ranges + n_trans
        if Q.shape != (N, N):
            raise InvalidArgument(EINVAL, "data matrix has shape %s, expected (%d, %d)" % (Q.shape, N, N))
        self.d, self.n, self.m, self.nt, self.N = int(d), int(n_poses), int(n_ranges), int(n_trans), int(N)
        self.rows = self.N  # getExpectedVariableSize(): d n + m after set_formulation(FORMULATION_IMPLICIT)
        self.nnz = int(Q.nnz)
        rp = np.ascontiguousarray(Q.indptr, dtype=np.int32)
        ci = np.ascontiguousarray(Q.indices, dtype=np.int32)
        va = np.ascontiguousarray(Q.data, dtype=np.float64)
        self._h = C.c_void_p()
        i32 = C.POINTER(C.c_int32)
        _check(lib.cora_b200_create(C.byref(self._h), C.c_int(device), C.c_void_p(stream or 0), C.c_int(d),
                                    C.c_int(n_poses), C.c_int(n_ranges), C.c_int(n_trans),
                                    rp.ctypes.data_as(i32), ci.ctypes.data_as(i32), _p(va), C.c_int64(self.nnz),
                                    C.c_int(preconditioner), C.c_double(reg_chol_max_cond)))
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.cora_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- configuration ------------------------------------------------------
    @property
    def effective_preconditioner(self):
        """The preconditioner applied (RegularizedCholesky falls back to Jacobi on non-chain graphs)."""
        v = C.c_int(0)
        _check(self._lib.cora_b200_effective_preconditioner(self._h, C.byref(v)))
        return v.value

    @property
    def last_cert_branch(self):
        v = C.c_int(0)
        _check(self._lib.cora_b200_last_cert_branch(self._h, C.byref(v)))
        return CERT_BRANCH[v.value]

    def set_preconditioner(self, preconditioner, reg_chol_max_cond=0.0):
        _check(self._lib.cora_b200_set_preconditioner(self._h, C.c_int(preconditioner),
                                                      C.c_double(reg_chol_max_cond)))

    def set_formulation(self, formulation):
        """Problem::setFormulation: FORMULATION_EXPLICIT / FORMULATION_IMPLICIT (translations marginalised)."""
        _check(self._lib.cora_b200_set_formulation(self._h, C.c_int(formulation)))
        v = C.c_int64(0)
        _check(self._lib.cora_b200_variable_rows(self._h, C.byref(v)))
        self.rows = int(v.value)

    def translation_explicit_solution(self, Y):
        """Problem::getTranslationExplicitSolution: (d n + m) x r -> N x r."""
        Y = self._mat(Y)
        out = np.empty((self.N, Y.shape[1]), order="F")
        _check(self._lib.cora_b200_translation_explicit_solution(self._h, Y.shape[1], _p(Y), _p(out)))
        return out

    @property
    def reg_lambda(self):
        v = C.c_double()
        _check(self._lib.cora_b200_get_reg_lambda(self._h, C.byref(v)))
        return v.value

    @reg_lambda.setter
    def reg_lambda(self, lam):
        _check(self._lib.cora_b200_set_reg_lambda(self._h, C.c_double(lam)))

    # -- tier 1 ---------------------------------------------------------------
    def _mat(self, A, r=None, rows=None):
        A = _f64(A)
        rows = self.rows if rows is None else rows
        if A.shape[0] != rows or (r is not None and A.shape[1] != r):
            raise InvalidArgument(EINVAL, "expected matrix of shape (%d, %s) but got (%d, %d)"
                                  % (rows, r if r is not None else "r", A.shape[0], A.shape[1]))
        return A

    def data_matrix_product(self, Y):
        Y = self._mat(Y)
        out = np.empty_like(Y, order="F")
        _check(self._lib.cora_b200_data_matrix_product(self._h, Y.shape[1], _p(Y), _p(out)))
        return out

    def evaluate_objective(self, Y):
        Y = self._mat(Y)
        f = C.c_double()
        _check(self._lib.cora_b200_objective(self._h, Y.shape[1], _p(Y), C.byref(f)))
        return f.value

    def euclidean_gradient(self, Y):
        Y = self._mat(Y)
        out = np.empty_like(Y, order="F")
        _check(self._lib.cora_b200_egrad(self._h, Y.shape[1], _p(Y), _p(out)))
        return out

    def riemannian_gradient(self, Y, egrad=None):
        Y = self._mat(Y)
        G = self._mat(egrad, Y.shape[1]) if egrad is not None else None
        out = np.empty_like(Y, order="F")
        _check(self._lib.cora_b200_rgrad(self._h, Y.shape[1], _p(Y), _p(G) if G is not None else None, _p(out)))
        return out

    def hessvec(self, Y, egrad, Ydot):
        Y = self._mat(Y)
        r = Y.shape[1]
        G = self._mat(egrad, r) if egrad is not None else None
        D = self._mat(Ydot, r)
        out = np.empty_like(Y, order="F")
        _check(self._lib.cora_b200_hessvec(self._h, r, _p(Y), _p(G) if G is not None else None, _p(D), _p(out)))
        return out

    def tangent_space_projection(self, Y, V):
        Y = self._mat(Y)
        V = self._mat(V, Y.shape[1])
        out = np.empty_like(Y, order="F")
        _check(self._lib.cora_b200_tangent_proj(self._h, Y.shape[1], _p(Y), _p(V), _p(out)))
        return out

    def precondition(self, V):
        V = self._mat(V)
        out = np.empty_like(V, order="F")
        _check(self._lib.cora_b200_precondition(self._h, V.shape[1], _p(V), _p(out)))
        return out

    def retract(self, Y, V):
        Y = self._mat(Y)
        V = self._mat(V, Y.shape[1])
        out = np.empty_like(Y, order="F")
        _check(self._lib.cora_b200_retract(self._h, Y.shape[1], _p(Y), _p(V), _p(out)))
        return out

    def project_to_manifold(self, A):
        A = self._mat(A)
        out = np.empty_like(A, order="F")
        _check(self._lib.cora_b200_project(self._h, A.shape[1], _p(A), _p(out)))
        return out

    def compute_lambda_blocks(self, Y):
        Y = self._mat(Y)
        st = np.zeros((self.d, self.d * self.n), order="F")
        ob = np.zeros(max(self.m, 1))
        _check(self._lib.cora_b200_lambda_blocks(self._h, Y.shape[1], _p(Y), _p(st), _p(ob)))
        return st, ob[: self.m]

    def certificate_product(self, Y, x):
        Y = self._mat(Y)
        x = self._mat(x, rows=self.N)  # S is the translation-explicit certificate matrix in either formulation
        out = np.empty_like(x, order="F")
        _check(self._lib.cora_b200_certificate_product(self._h, Y.shape[1], _p(Y), x.shape[1], _p(x), _p(out)))
        return out

    # -- tier 2 ---------------------------------------------------------------
    @staticmethod
    def _alloc_result(cap):
        res = TntResultC()
        keep = {}
        res.trace_capacity = cap
        for name in ("objective_values", "gradient_norms", "preconditioned_gradient_norms",
                     "trust_region_radius", "time", "update_step_norms", "update_step_M_norms", "gain_ratios"):
            keep[name] = np.zeros(cap)
            setattr(res, name, _p(keep[name]))
        keep["inner_iterations"] = np.zeros(cap, dtype=np.int32)
        res.inner_iterations = keep["inner_iterations"].ctypes.data_as(C.POINTER(C.c_int32))
        return res, keep

    @staticmethod
    def _unpack_result(res, keep, x=None):
        k = res.num_outer
        out = TntResult(x=x, f=res.f, gradfx_norm=res.gradfx_norm,
                        preconditioned_grad_f_x_norm=res.preconditioned_gradfx_norm,
                        status=TNT_STATUS[res.status], elapsed_time=res.elapsed_time,
                        device_time=res.device_time, kernel_launches=res.kernel_launches)
        for name in ("objective_values", "gradient_norms", "preconditioned_gradient_norms",
                     "trust_region_radius", "time"):
            setattr(out, name, keep[name][: k + 1].tolist())
        for name in ("update_step_norms", "update_step_M_norms", "gain_ratios", "inner_iterations"):
            setattr(out, name, keep[name][:k].tolist())
        return out

    def tnt(self, X0, params: Optional[TntParams] = None) -> TntResult:
        X0 = self._mat(X0)
        params = params or default_tnt_params()
        res, keep = self._alloc_result(params.max_iterations + 2)
        out = np.empty_like(X0, order="F")
        _check(self._lib.cora_b200_tnt(self._h, X0.shape[1], _p(X0), C.byref(params), _p(out), C.byref(res)))
        return self._unpack_result(res, keep, out)

    def set_iterate(self, X):
        X = self._mat(X)
        _check(self._lib.cora_b200_set_iterate(self._h, X.shape[1], _p(X)))

    def set_iterate_ptr(self, r, host_ptr):
        """X given as a raw host pointer (e.g. a pinned torch tensor, column-major N x r)."""
        _check(self._lib.cora_b200_set_iterate(self._h, C.c_int(r), C.cast(host_ptr, _PD)))

    def get_iterate(self, r):
        out = np.empty((self.rows, r), order="F")
        _check(self._lib.cora_b200_get_iterate(self._h, r, _p(out)))
        return out

    def get_iterate_ptr(self, r, host_ptr):
        _check(self._lib.cora_b200_get_iterate(self._h, C.c_int(r), C.cast(host_ptr, _PD)))

    def tnt_resident(self, params: Optional[TntParams] = None) -> TntResult:
        params = params or default_tnt_params()
        res, keep = self._alloc_result(params.max_iterations + 2)
        _check(self._lib.cora_b200_tnt_resident(self._h, C.byref(params), C.byref(res)))
        return self._unpack_result(res, keep)

    def snapshot_iterate(self):
        _check(self._lib.cora_b200_snapshot_iterate(self._h))

    def restore_iterate(self):
        _check(self._lib.cora_b200_restore_iterate(self._h))

    def profile_hessvec(self, max_samples):
        _check(self._lib.cora_b200_profile_hessvec(self._h, C.c_int(max_samples)))

    def profile_read(self, capacity=65536):
        ms = np.zeros(capacity, dtype=np.float32)
        n = C.c_int(0)
        _check(self._lib.cora_b200_profile_read(self._h, C.c_int(capacity),
                                                ms.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n)))
        return ms[: n.value].astype(np.float64)

    PHASES = ["hub", "grad", "hess", "update", "pupdate", "retract", "precond", "cginit", "sync", "misc",
              "q.wait", "q.qx", "q.epi", "q.store", "ch.pre", "ch.fwd", "ch.bwd", "ch.border", "ch.post", "smid", "reduce"]

    def phase_profile(self):
        """In-kernel phase profile of the last persistent TNT call: {phase: (total_us, count)}, grid, barriers."""
        tot = np.zeros(24)
        cnt = np.zeros(24, dtype=np.int64)
        n, grid, bars = C.c_int(0), C.c_int(0), C.c_int64(0)
        _check(self._lib.cora_b200_phase_profile(self._h, C.c_int(24), _p(tot),
                                                 cnt.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(n),
                                                 C.byref(grid), C.byref(bars)))
        prof = {name: (float(tot[i]), int(cnt[i])) for i, name in enumerate(self.PHASES[: n.value])}
        return prof, grid.value, bars.value

    def phase_profile_ctas(self):
        """{phase: (avg us on the slowest CTA, avg us on the median CTA)} of the last persistent TNT call."""
        mx, md = np.zeros(24), np.zeros(24)
        _check(self._lib.cora_b200_phase_profile_ctas(self._h, C.c_int(24), _p(mx), _p(md)))
        return {name: (float(mx[i]), float(md[i])) for i, name in enumerate(self.PHASES) if mx[i] > 0}

    def get_work_vector(self, which, r):
        out = np.empty((self.rows, r), order="F")
        _check(self._lib.cora_b200_get_work_vector(self._h, C.c_int(which), C.c_int(r), _p(out)))
        return out

    def device_vectors(self, r):
        """(x_ptr, qx_ptr): raw device addresses of the resident iterate and of Q*X, N x r row-major in the internal
        row order (row_order()); marks rank r resident."""
        x, qx = C.c_void_p(), C.c_void_p()
        _check(self._lib.cora_b200_device_vectors(self._h, C.c_int(r), C.byref(x), C.byref(qx)))
        return x.value, qx.value

    def row_order(self):
        """internal_to_reference[i] = reference row held by internal row i."""
        out = np.zeros(self.N, dtype=np.int32)
        _check(self._lib.cora_b200_row_order(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def spmm_resident(self, reps):
        ms = C.c_float()
        _check(self._lib.cora_b200_spmm_resident(self._h, C.c_int(reps), C.byref(ms)))
        return ms.value

    def certify_solution(self, Y, eta, nx, bootstrap=None, max_iters=500) -> CertResults:
        Y = self._mat(Y)
        r = Y.shape[1]
        B = _f64(bootstrap) if bootstrap is not None and np.size(bootstrap) else None
        cap = max(nx, r + 2)
        ev = np.zeros((self.rows, cap), order="F")
        x = np.zeros(self.rows)
        cert, ncols = C.c_int(), C.c_int()
        theta, iters = C.c_double(), C.c_int64()
        _check(self._lib.cora_b200_certify(self._h, r, _p(Y), C.c_double(eta), C.c_int(nx),
                                           _p(B) if B is not None else None,
                                           C.c_int(B.shape[1] if B is not None else 0), C.c_int(max_iters),
                                           C.byref(cert), C.byref(theta), _p(x), _p(ev), C.c_int(cap),
                                           C.byref(ncols), C.byref(iters)))
        return CertResults(bool(cert.value), theta.value, x, ev[:, : ncols.value].copy(), iters.value)

    def saddle_escape(self, Y, theta, v, gradient_tolerance=1e-4, preconditioned_gradient_tolerance=1e-4):
        Y = self._mat(Y)
        v = np.ascontiguousarray(v, dtype=np.float64).ravel()
        if v.shape[0] != self.rows:
            raise InvalidArgument(EINVAL, "v must have N entries")
        out = np.empty((self.rows, Y.shape[1] + 1), order="F")
        _check(self._lib.cora_b200_saddle_escape(self._h, Y.shape[1] + 1, _p(Y), C.c_double(theta), _p(v),
                                                 C.c_double(gradient_tolerance),
                                                 C.c_double(preconditioned_gradient_tolerance), _p(out)))
        return out

    def project_solution(self, Y):
        Y = self._mat(Y)
        out = np.empty((self.rows, self.d), order="F")
        _check(self._lib.cora_b200_project_solution(self._h, Y.shape[1], _p(Y), _p(out)))
        return out

    def gather_best(self, comm: "NcclComm", world_size, rank, f, certified, X):
        """All ranks call this with their own (f, certified, X [N x r_max]); returns (winner_rank, winner_f, X_winner)."""
        X = np.array(self._mat(X), order="F", copy=True)
        w, wf = C.c_int(0), C.c_double(0)
        _check(self._lib.cora_b200_gather_best(comm._c, self._h, C.c_int(world_size), C.c_int(rank),
                                               C.c_int(X.shape[1]), C.c_double(f), C.c_int(int(certified)), _p(X),
                                               C.byref(w), C.byref(wf)))
        return w.value, wf.value, X

    def psd_test(self, eta, Y=None, r=None):
        """Is S(Y) + eta I positive definite (Cholesky on the device)?  Y None: the resident iterate of rank r."""
        v = C.c_int(0)
        if Y is None:
            _check(self._lib.cora_b200_psd_test(self._h, C.c_int(r), None, C.c_double(eta), C.byref(v)))
        else:
            Y = self._mat(Y)
            _check(self._lib.cora_b200_psd_test(self._h, C.c_int(Y.shape[1]), _p(Y), C.c_double(eta), C.byref(v)))
        return bool(v.value)

    def debug_min_eigenpair(self, max_iters=200):
        """Test hook: (theta, x, steps) of the smallest eigenpair of the handle's matrix by the device Lanczos."""
        th, st = C.c_double(0), C.c_int(0)
        x = np.zeros(self.N)
        _check(self._lib.cora_b200_debug_min_eigenpair(self._h, C.c_int(max_iters), C.byref(th), _p(x), C.byref(st)))
        return th.value, x, st.value

    def gather_best_resident(self, comm: "NcclComm", world_size, rank, f, certified):
        """Device-resident variant: the winner's resident iterate becomes every rank's resident iterate."""
        w, wf = C.c_int(0), C.c_double(0)
        _check(self._lib.cora_b200_gather_best_resident(comm._c, self._h, C.c_int(world_size), C.c_int(rank),
                                                        C.c_double(f), C.c_int(int(certified)), C.byref(w), C.byref(wf)))
        return w.value, wf.value

    def solve(self, X0, max_rank=20, params: Optional[TntParams] = None, verbose=False):
        X0 = self._mat(X0)
        params = params or default_tnt_params()
        cap = 2 * (max_rank + 2)
        stages = (StageC * cap)()
        res = SolveResultC()
        res.stage_capacity = cap
        res.stages = C.cast(stages, C.POINTER(StageC))
        out = np.empty((self.rows, self.d), order="F")
        _check(self._lib.cora_b200_solve(self._h, X0.shape[1], _p(X0), C.c_int(max_rank), C.byref(params),
                                         C.c_int(int(verbose)), _p(out), C.byref(res)))
        st = []
        for i in range(min(res.num_stages, cap)):
            s = stages[i]
            st.append(dict(rank=s.rank, status=TNT_STATUS[s.status], outer=s.num_outer,
                           certified=bool(s.certified), cg=s.cg_iterations, f=s.f, grad=s.gradfx_norm,
                           theta=s.theta, eta=s.eta, tnt_seconds=s.tnt_seconds, cert_seconds=s.cert_seconds,
                           cert_branch=CERT_BRANCH[s.cert_branch]))
        return dict(x=out, f=res.f, lifted_f=res.lifted_f, final_rank=res.final_rank,
                    lifted_rank=res.lifted_rank, certified=bool(res.certified),
                    refined_certified=bool(res.refined_certified),
                    total_cg_iterations=res.total_cg_iterations, seconds=res.seconds, stages=st)
