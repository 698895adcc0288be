"""ctypes binding of oracle/eloc_oracle.c (liboracle.so) — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this.  build() compiles the C restatement with gcc; see eloc_oracle.c for the
reference lines it follows.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
SRC = os.path.join(HERE, "eloc_oracle.c")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        p, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        L.oracle_table_create.restype = p
        L.oracle_table_create.argtypes = [p, p, p, i64, i32, i32, i32, i32]
        L.oracle_table_destroy.argtypes = [p]
        L.oracle_table_groups.restype = i64
        L.oracle_table_groups.argtypes = [p]
        L.oracle_restricted_index.argtypes = [p, p, i64, p]
        L.oracle_hij_dense.argtypes = [p, p, i64, p]
        L.oracle_rows_count.argtypes = [p, p, i64, p]
        L.oracle_rows_fill.argtypes = [p, p, i64, p, p, p]
        L.oracle_eloc.restype = i32
        L.oracle_eloc.argtypes = [p, p, p, i64, p, p, i64, p, i32]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _keys(x, W):
    a = np.ascontiguousarray(x)
    if a.dtype != np.uint64:
        a = a.astype(np.int64).astype(np.uint64)
    return np.ascontiguousarray(a.reshape(-1, W))


class COracleTable:
    """Term table handle of the C oracle (n_alpha=None => no sector filter)."""

    def __init__(self, xy, yz, coeff, n_qubits, n_alpha=None, n_beta=None):
        self.W = 1 if n_qubits <= 63 else 2
        self.n_qubits = n_qubits
        xy, yz = _keys(xy, self.W), _keys(yz, self.W)
        c = np.ascontiguousarray(coeff, np.float64).reshape(-1)
        self.K = len(c)
        na = -1 if n_alpha is None else int(n_alpha)
        nb = -1 if n_beta is None else int(n_beta)
        self._h = lib().oracle_table_create(_ptr(xy), _ptr(yz), _ptr(c), self.K, self.W, n_qubits, na, nb)
        self.G = lib().oracle_table_groups(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_table_destroy(self._h)
            self._h = None

    def restricted_index(self, keys):
        k = _keys(keys, self.W)
        out = np.empty(len(k), np.int64)
        lib().oracle_restricted_index(self._h, _ptr(k), len(k), _ptr(out))
        return out

    def hij_dense(self, states):
        s = _keys(states, self.W)
        out = np.empty(len(s) * self.G, np.float64)
        lib().oracle_hij_dense(self._h, _ptr(s), len(s), _ptr(out))
        return out

    def rows(self, states):
        """-> (indptr[M+1], col_keys[nnz, W], vals[nnz]); columns in ascending unique-XY order."""
        s = _keys(states, self.W)
        counts = np.empty(len(s), np.int64)
        lib().oracle_rows_count(self._h, _ptr(s), len(s), _ptr(counts))
        indptr = np.zeros(len(s) + 1, np.int64)
        np.cumsum(counts, out=indptr[1:])
        cols = np.empty((int(indptr[-1]), self.W), np.uint64)
        vals = np.empty(int(indptr[-1]), np.float64)
        lib().oracle_rows_fill(self._h, _ptr(s), len(s), _ptr(indptr), _ptr(cols), _ptr(vals))
        return indptr, cols, vals

    def local_energy(self, states, psi, table_keys=None, table_psi=None, reference_order=True):
        s = _keys(states, self.W)
        p = np.ascontiguousarray(psi).astype(np.complex128)
        tk = s if table_keys is None else _keys(table_keys, self.W)
        tp = p if table_psi is None else np.ascontiguousarray(table_psi).astype(np.complex128)
        out = np.empty(len(s), np.complex128)
        lib().oracle_eloc(self._h, _ptr(s), _ptr(p), len(s), _ptr(tk), _ptr(tp), len(tk), _ptr(out),
                          0 if reference_order else 1)
        return out
