"""DeviceTermTable — Python owner of a device-resident Pauli term table (naqs_table_t).

The fused Level-1 entry (`local_energy`) is what `calculate_local_energy` calls; `rows`,
`hij_dense`, `coupled_state_set`, `restricted_index` and `stats` expose the other C-ABI entries.
Semantics follow the reference exactly (see include/naqs_eloc.h for file:line citations).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import LOOKUP_AUTO, LOOKUP_DENSE, LOOKUP_HASH  # noqa: F401


class DeviceTermTable:
    def __init__(self, xy, yz, coeff, n_qubits, n_alpha=None, n_beta=None, device=None):
        """xy, yz: packed masks in REFERENCE TERM ORDER (ints, integer arrays or [K, words] uint64);
        coeff: float64 [K] (src/optimizer/hamiltonian.py:373-430).  n_alpha=None => no sector filter."""
        self.device = _lib.require_cuda(device)
        self.n_qubits = int(n_qubits)
        self.words = _lib.n_words(self.n_qubits)
        self.n_alpha = None if n_alpha is None else int(n_alpha)
        self.n_beta = None if n_beta is None else int(n_beta)
        xy = _lib.keys_to_numpy(xy, self.words)
        yz = _lib.keys_to_numpy(yz, self.words)
        c = np.ascontiguousarray(np.asarray(coeff).reshape(-1), np.float64)
        if not (len(xy) == len(yz) == len(c)):
            raise ValueError("xy, yz and coeff must have the same length")
        self._h = C.c_void_p()
        lib = _lib.load()
        _lib.check(lib.naqs_table_create(C.byref(self._h), _lib.ptr(xy), _lib.ptr(yz), _lib.ptr(c), len(c), self.words,
                                         self.n_qubits, -1 if n_alpha is None else self.n_alpha,
                                         -1 if n_beta is None else self.n_beta, self.device.index), "naqs_table_create")
        info = np.zeros(8, np.int64)
        _lib.check(lib.naqs_table_info(self._h, _lib.ptr(info)))
        self.K, self.Kxy, self.Kyz = int(info[0]), int(info[1]), int(info[2])
        self._lookup_built = False
        self.precision = np.dtype(np.float64)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().naqs_table_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def set_algo(self, algo):
        """Kernel formulation of the fused path: "sliced" (nibble-sliced parity + group LUT, default) or "direct"
        (AND/POPC walk).  Both accumulate every H_ij in reference order; they differ only in speed."""
        code = {"sliced": 0, "direct": 1}[algo]
        _lib.check(_lib.load().naqs_table_set_algo(self._h, code), "naqs_table_set_algo")
        return self

    def set_precision(self, dtype):
        """Accumulation type of H_ij: np.float64 (default; what the reference's experiments use, _base.py:234) or np.float32
        (the constructor default of PauliHamiltonian.get, hamiltonian.py:48): coefficients must already be float32 values and
        every partial sum is rounded to float32 like __inner_int64_float does, so H_ij equals the reference's float32 matrix
        bit for bit.  Any other type (np.float128) -> TypeError: it has no device type."""
        dt = np.dtype(dtype)
        bits = {np.dtype(np.float64): 64, np.dtype(np.float32): 32}.get(dt, dt.itemsize * 8 if dt.kind == "f" else 0)
        _lib.check(_lib.load().naqs_table_set_precision(self._h, bits), "naqs_table_set_precision")
        self.precision = np.dtype(np.float32 if bits == 32 else np.float64)
        return self

    # ------------------------------------------------------------------ helpers
    def _keys(self, x):
        return _lib.keys_to_device(x, self.words, self.device)

    def _stream(self):
        return _lib.stream_ptr(self.device)

    # ------------------------------------------------------------------ amplitude lookup
    def build_lookup(self, keys, psi, kind=LOOKUP_AUTO, assume_unique=False, duplicates_equal=False):
        """(key, psi) pairs of the sampled batch -> device lookup structure (duplicates summed).
        assume_unique=True is the reference's own contract at the call site (energy.py:245 passes assume_unique=True): it
        lets the dense table keep complex64 amplitudes as 8-byte entries.  duplicates_equal=True: keys may repeat but every
        copy carries the same amplitude (psi is a function of the state, e.g. all-gathered shards of several ranks) — copies
        are dropped instead of summed."""
        k = self._keys(keys)
        p, code = _lib.psi_to_device(psi, self.device)
        if p.shape[0] != k.shape[0]:
            raise ValueError("keys and psi must have the same length")
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().naqs_lookup_build(self._h, _lib.ptr(k), _lib.ptr(p), code, k.shape[0],
                                                     kind | (_lib.LOOKUP_ASSUME_UNIQUE if assume_unique else 0) |
                                                     (_lib.LOOKUP_DUPLICATES_EQUAL if duplicates_equal else 0), self._stream()),
                       "naqs_lookup_build")
        self._lookup_built = True
        return self

    def attach_dense32(self, table_tensor):
        """Use a caller-owned complex64 direct-address table (float32 [2^N, 2] CUDA tensor, absent = -0.0) as the lookup."""
        if table_tensor.dtype not in (torch.float32, torch.int32) or table_tensor.numel() != 2 * (1 << self.n_qubits):
            raise ValueError("dense32 table must be float32/int32 [2^N, 2]")
        _lib.check(_lib.load().naqs_lookup_attach_dense32(self._h, _lib.ptr(table_tensor), 1 << self.n_qubits), "naqs_lookup_attach_dense32")
        self._dense32_keepalive = table_tensor
        self._lookup_built = True
        return self

    # ------------------------------------------------------------------ fused E_loc (Level-1)
    def local_energy(self, states, psi, table_keys=None, table_psi=None, kind=LOOKUP_AUTO, out=None, rebuild_lookup=True,
                     assume_unique=False):
        """E_loc of every state in `states` (src/optimizer/energy.py:245-248), complex128.

        states/psi on the host (numpy / CPU tensors) or on the device (CUDA tensors).  The lookup table is
        the batch itself unless (table_keys, table_psi) are given.  Returns a CUDA float64 tensor [M, 2]
        (re, im) — use `_lib.complex_from_pairs` / `.cpu()` to bring it back."""
        k = self._keys(states)
        p, code = _lib.psi_to_device(psi, self.device)
        M = k.shape[0]
        if p.shape[0] != M:
            raise ValueError("states and psi must have the same length")
        if rebuild_lookup or not self._lookup_built:
            if table_keys is None:
                self.build_lookup(k, torch.view_as_complex(p), kind, assume_unique)
            else:
                self.build_lookup(table_keys, table_psi, kind, assume_unique)
        if out is None:
            out = torch.empty((M, 2), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().naqs_eloc(self._h, _lib.ptr(k), _lib.ptr(p), code, M, _lib.ptr(out), self._stream()), "naqs_eloc")
        return out

    def apply_H(self, states, v, out=None):
        """Matrix-free (H v)[rows]: H restricted to the basis `states`, v complex on that basis -> CUDA float64 [M, 2].
        No matrix is formed; one fused kernel per application (naqs_lookup_build holds v, naqs_apply_h walks the rows)."""
        k = self._keys(states)
        self.build_lookup(k, v)
        if out is None:
            out = torch.empty((k.shape[0], 2), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().naqs_apply_h(self._h, _lib.ptr(k), k.shape[0], _lib.ptr(out), self._stream()), "naqs_apply_h")
        return out

    def _host_keys(self, x):
        """-> (contiguous numpy array, itemsize code) without copying when the caller already holds int16/int32/uint64 keys."""
        if torch.is_tensor(x):
            x = x.detach().cpu().numpy()
        a = np.asarray(x)
        if self.words == 1 and a.dtype in (np.int16, np.uint16, np.int32, np.uint32) and a.flags.c_contiguous:
            return a.reshape(-1), a.dtype.itemsize
        return _lib.keys_to_numpy(a, self.words), 8

    def local_energy_host(self, states, psi, table_keys=None, table_psi=None, out=None, kind=LOOKUP_AUTO, assume_unique=False,
                          out_dtype=np.complex128, wait=True):
        """Host-buffer path (numpy / CPU tensors in, numpy out) through naqs_eloc_host: upload -> lookup build -> fused
        kernel -> download, synchronous.  Keys may be the reference's int16 / int32 state indices or uint64 words; psi
        complex64 / complex128; out_dtype complex128, or complex64 = the float32 pairs the reference returns to torch.
        Page-locked inputs / `out` (e.g. numpy views of pinned torch tensors) make the copies run at full PCIe rate.
        wait=False: only enqueue (naqs_eloc_host_begin) — call `local_energy_host_wait()` before reading `out`; with one table
        per batch in flight the copies of one batch overlap the kernel of another."""
        k, ksz = self._host_keys(states)
        n = len(k)
        if torch.is_tensor(psi):
            psi = (torch.view_as_complex(psi.detach().contiguous()) if not psi.is_complex() else psi.detach()).numpy()
        p = np.ascontiguousarray(psi).reshape(-1)
        if p.dtype not in (np.complex64, np.complex128):
            p = p.astype(np.complex128)
        code = _lib.NAQS_C64 if p.dtype == np.complex64 else _lib.NAQS_C128
        out_dtype = np.dtype(out_dtype)
        if out_dtype not in (np.dtype(np.complex64), np.dtype(np.complex128)):
            raise TypeError("out_dtype must be complex64 or complex128")
        if out is None:
            out = np.empty(n, out_dtype)
        elif out.dtype != out_dtype or out.shape != (n,) or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous array of out_dtype with one entry per state")
        tk = tp = None
        T = 0
        if table_keys is not None:
            tk, tksz = self._host_keys(table_keys)
            if tksz != ksz:
                tk, k, ksz = _lib.keys_to_numpy(tk, self.words), _lib.keys_to_numpy(k, self.words), 8
            tp = np.ascontiguousarray(table_psi).astype(p.dtype).reshape(-1)
            T = len(tk)
        fn = _lib.load().naqs_eloc_host if wait else _lib.load().naqs_eloc_host_begin
        _lib.check(fn(self._h, _lib.ptr(k), ksz, _lib.ptr(p), code, n, _lib.ptr(tk), _lib.ptr(tp), T,
                      kind | (_lib.LOOKUP_ASSUME_UNIQUE if assume_unique else 0), _lib.ptr(out),
                      _lib.NAQS_C64 if out_dtype == np.dtype(np.complex64) else _lib.NAQS_C128), "naqs_eloc_host")
        if not wait:
            self._host_keepalive = (k, p, tk, tp, out)  # the buffers must outlive the asynchronous copies
        return out

    def local_energy_host_wait(self):
        _lib.check(_lib.load().naqs_eloc_host_end(self._h), "naqs_eloc_host_end")
        self._host_keepalive = None

    def check(self):
        """Raise IndexError if a key outside [0, 2^n_qubits) was passed since the last check (the reference raises it at
        hilbert.py:607-640 / hamiltonian.py:94).  Synchronises the current stream; the host-buffer entry checks itself."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().naqs_table_check(self._h, self._stream()), "naqs_table_check")

    # ------------------------------------------------------------------ stored rows (CSR / coupled sets)
    def rows(self, states, with_restricted_index=True):
        """Stored couplings of each state (src/optimizer/hamiltonian.py:301-363):
        -> (indptr [M+1] int64, col_keys [nnz, words] int64 bit patterns, col_ridx [nnz] int64 or None, vals [nnz] float64),
        all CUDA tensors; columns in ascending unique-XY order."""
        lib = _lib.load()
        k = self._keys(states)
        M = k.shape[0]
        counts = torch.empty(M, dtype=torch.int64, device=self.device)
        indptr = torch.empty(M + 1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            st = self._stream()
            _lib.check(lib.naqs_rows_count(self._h, _lib.ptr(k), M, _lib.ptr(counts), st), "naqs_rows_count")
            _lib.check(lib.naqs_exclusive_scan(self._h, _lib.ptr(counts), M, _lib.ptr(indptr), st), "naqs_exclusive_scan")
            nnz = int(indptr[-1].item())
            cols = torch.empty((nnz, self.words), dtype=torch.int64, device=self.device)
            ridx = torch.empty(nnz, dtype=torch.int64, device=self.device) if with_restricted_index else None
            vals = torch.empty(nnz, dtype=torch.float64, device=self.device)
            _lib.check(lib.naqs_rows_fill(self._h, _lib.ptr(k), M, _lib.ptr(indptr), _lib.ptr(cols), _lib.ptr(ridx), _lib.ptr(vals), st),
                       "naqs_rows_fill")
        return indptr, cols, ridx, vals

    def hij_dense(self, states):
        """[M, Kxy] float64: exactly get_Hij_cy's output (src_cpp/hamiltonian_math.pyx:200-288)."""
        k = self._keys(states)
        out = torch.empty((k.shape[0], self.Kxy), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().naqs_hij_dense(self._h, _lib.ptr(k), k.shape[0], _lib.ptr(out), self._stream()), "naqs_hij_dense")
        return out

    def unique_keys(self, keys):
        """Sorted unique keys (device radix sort + adjacent unique): [U, words] int64 bit patterns."""
        k = self._keys(keys)
        out = torch.empty_like(k)
        n_unique = C.c_int64(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().naqs_unique_keys(self._h, _lib.ptr(k), k.shape[0], _lib.ptr(out), C.byref(n_unique), self._stream()),
                       "naqs_unique_keys")
        return out[: n_unique.value]

    def coupled_state_set(self, states):
        """get_coupled_state_idxs(return_unique=True) (src/optimizer/hamiltonian.py:122-132), as keys."""
        _, cols, _, _ = self.rows(states, with_restricted_index=False)
        return self.unique_keys(cols)

    def restricted_index(self, keys):
        """full2restricted_idx (src/utils/hilbert.py:607-640) by combinatorial ranking: int64 [n], -1 outside the sector."""
        k = self._keys(keys)
        out = torch.empty(k.shape[0], dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().naqs_restricted_index(self._h, _lib.ptr(k), k.shape[0], _lib.ptr(out), self._stream()),
                       "naqs_restricted_index")
        return out

    def stats(self, eloc, weights=None):
        """[sum w, sum w Re E, sum w Im E, sum w (Re E)^2, n] in fp64 (src/optimizer/energy.py:328,372-375)."""
        e = eloc if torch.is_tensor(eloc) else torch.from_numpy(np.ascontiguousarray(eloc))
        if e.is_complex():
            e = torch.view_as_real(e.to(torch.complex128))
        e = e.to(self.device, torch.float64).contiguous()
        w = None
        if weights is not None:
            w = (weights if torch.is_tensor(weights) else torch.from_numpy(np.asarray(weights))).to(self.device, torch.float64).reshape(-1).contiguous()
        out = torch.empty(5, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().naqs_eloc_stats(self._h, _lib.ptr(e), _lib.ptr(w), e.shape[0], _lib.ptr(out), self._stream()),
                       "naqs_eloc_stats")
        return out
