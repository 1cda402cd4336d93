"""ctypes loader of oracle/cpu_ref.cpp -- TEST / BASELINE INFRASTRUCTURE ONLY.

The C++ restatement of the reference's CPU path (Jacobi-preconditioned TNT + STPCG with the
reference's data layouts and operation counts).  Built with `-O3 -march=native` as the reference is
(CMakeLists.txt:20-23,55-58), therefore per host CPU: the shared object lives in
oracle/_build/<cpu signature>/ and is (re)built on first use on a new machine.
Only tests/, bench.py's cpu_baseline / `--impl reference` legs and __graft_entry__.smoke() may import this.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _cpu_signature():
    try:
        with open("/proc/cpuinfo") as f:
            lines = [l for l in f if l.startswith(("model name", "flags"))][:2]
    except OSError:
        lines = []
    return hashlib.sha1("".join(lines).encode()).hexdigest()[:12]


def build():
    out = os.path.join("_build", _cpu_signature())
    subprocess.run(["make", "-s", "-C", _HERE, "OUT=" + out], check=True)
    return os.path.join(_HERE, out, "libcora_cpu_ref.so")


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.cpu_ref_last_error.restype = C.c_char_p
        lib.cpu_ref_spmm_count.restype = C.c_int64
        lib.cpu_ref_set_reg_cholesky.argtypes = [C.c_void_p, C.c_double]
        _lib = lib
    return _lib


_PD = C.POINTER(C.c_double)


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


class CpuRef:
    """One problem on the host: CSR data matrix in the reference row order + Jacobi preconditioner."""

    def __init__(self, d, n_poses, n_ranges, n_trans, Q, preconditioner=1, threads=None, reg_lambda=None):
        import scipy.sparse as sp
        lib = load()
        Q = sp.csr_matrix(Q)
        Q.sort_indices()
        self.N = d * n_poses + n_ranges + n_trans
        assert Q.shape == (self.N, self.N)
        rp = np.ascontiguousarray(Q.indptr, dtype=np.int32)
        ci = np.ascontiguousarray(Q.indices, dtype=np.int32)
        va = np.ascontiguousarray(Q.data, dtype=np.float64)
        self._h = C.c_void_p()
        i32 = C.POINTER(C.c_int32)
        if threads:
            lib.cpu_ref_set_threads(C.c_int(int(threads)))
        rc = lib.cpu_ref_create(C.byref(self._h), C.c_int(d), C.c_int(n_poses), C.c_int(n_ranges), C.c_int(n_trans),
                                rp.ctypes.data_as(i32), ci.ctypes.data_as(i32), va.ctypes.data_as(_PD),
                                C.c_int64(Q.nnz), C.c_int(1 if preconditioner == 3 else preconditioner))
        if rc:
            raise RuntimeError(lib.cpu_ref_last_error().decode())
        self._lib = lib
        if preconditioner == 3:   # RegularizedCholesky: lambda = ||Q||_2 / (c - 1), src/CORA_problem.cpp:556-591
            if reg_lambda is None:
                raise ValueError("RegularizedCholesky needs reg_lambda")
            self.set_reg_cholesky(reg_lambda)

    @property
    def threads(self):
        return int(self._lib.cpu_ref_threads())

    def close(self):
        if self._h:
            self._lib.cpu_ref_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, r, *mats):
        out = np.empty((self.N, r), order="F")
        args = [m.ctypes.data_as(_PD) for m in mats]
        rc = getattr(self._lib, name)(self._h, C.c_int(r), *args, out.ctypes.data_as(_PD))
        if rc:
            raise RuntimeError(self._lib.cpu_ref_last_error().decode())
        return out

    def data_matrix_product(self, Y):
        Y = _f(Y)
        return self._call("cpu_ref_data_matrix_product", Y.shape[1], Y)

    def hessvec(self, Y, G, Yd):
        Y, G, Yd = _f(Y), _f(G), _f(Yd)
        return self._call("cpu_ref_hessvec", Y.shape[1], Y, G, Yd)

    def tangent_space_projection(self, Y, V):
        Y, V = _f(Y), _f(V)
        return self._call("cpu_ref_tangent_proj", Y.shape[1], Y, V)

    def project_to_manifold(self, A):
        A = _f(A)
        return self._call("cpu_ref_project", A.shape[1], A)

    def set_reg_cholesky(self, lam):
        """Select RegularizedCholesky with regularisation `lam` (= ||Q||_2 / (c - 1) of the reference)."""
        rc = self._lib.cpu_ref_set_reg_cholesky(self._h, C.c_double(float(lam)))
        if rc:
            raise RuntimeError(self._lib.cpu_ref_last_error().decode())

    def precondition(self, V):
        V = _f(V)
        return self._call("cpu_ref_precondition", V.shape[1], V)

    def spmm_count(self):
        return int(self._lib.cpu_ref_spmm_count(self._h))

    def tnt(self, X0, params):
        """params: cora_b200.capi.TntParams (the same C struct).  Returns cora_b200.capi.TntResult."""
        from cora_b200 import capi   # struct definitions only (no library call)
        X0 = _f(X0)
        res, keep = capi.Handle._alloc_result(params.max_iterations + 2)
        out = np.empty_like(X0, order="F")
        rc = self._lib.cpu_ref_tnt(self._h, C.c_int(X0.shape[1]), X0.ctypes.data_as(_PD), C.byref(params),
                                   out.ctypes.data_as(_PD), C.byref(res))
        if rc:
            raise RuntimeError(self._lib.cpu_ref_last_error().decode())
        return capi.Handle._unpack_result(res, keep, out)


def chain_posdef(d, n_poses, n_ranges, n_trans, S, shift):
    """Cholesky test of S + shift I (S: symmetric sparse, reference row order) on a chain + landmark graph."""
    import scipy.sparse as sp
    lib = load()
    S = sp.csr_matrix(S)
    S.sort_indices()
    rp = np.ascontiguousarray(S.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(S.indices, dtype=np.int32)
    va = np.ascontiguousarray(S.data, dtype=np.float64)
    pd = C.c_int(0)
    i32 = C.POINTER(C.c_int32)
    rc = lib.cpu_ref_chain_posdef(C.c_int(d), C.c_int(n_poses), C.c_int(n_ranges), C.c_int(n_trans),
                                  rp.ctypes.data_as(i32), ci.ctypes.data_as(i32), va.ctypes.data_as(_PD),
                                  C.c_double(float(shift)), C.byref(pd))
    if rc:
        raise RuntimeError(lib.cpu_ref_last_error().decode())
    return bool(pd.value)
