"""PauliHamiltonian — B200-backed mirror of src/optimizer/hamiltonian.py.

`PauliHamiltonian.get(hilbert, qubit_hamiltonian, ...)` keeps the reference's signature and returns
an object with the reference's methods (update_H / get_H / get_restricted_H / get_coupled_state_idxs /
freeze_H / unfreeze_H / is_frozen / save / load), so src/optimizer/energy.py runs unchanged on top of it.
Rows are produced by the CUDA rows kernels (naqs_rows_count / naqs_rows_fill) instead of the numpy
broadcast + popcount_parity + get_Hij_cy pipeline; the scipy CSR cache semantics are kept.

New, and what `calculate_local_energy` should call: `local_energy(states_idx, psi)` — the fused,
stateless device path (no H cache, no M x Kyz / M x Kxy intermediates).
"""
import os

import numpy as np
import torch
from scipy.sparse import csr_matrix, load_npz, save_npz

from . import _lib
from .hilbert import Encoding
from .pauli import pack_terms
from .table import DeviceTermTable

# Quirk q1 of SURVEY.md §8a: the reference's get_H returns get_restricted_H() — rows AND columns in restricted order —
# whenever the batch has as many states as the sector (hamiltonian.py:100-105), whatever order the batch is in, so H and
# psi are misaligned for a full sample (LiH: every VMC iteration samples all 225 sector states, in ascending key order).
# Default: NOT reproduced (the caller's order is kept, E_loc is the physical one).  NAQS_ELOC_REFERENCE_QUIRKS=1 (or
# install(reference_quirks=True)) reproduces it, for bitwise trajectory parity with the reference on such batches.
REFERENCE_QUIRKS = os.environ.get("NAQS_ELOC_REFERENCE_QUIRKS", "0") not in ("", "0")


class PauliHamiltonian:

    @staticmethod
    def _format_fnames(fname, fname_H=None):
        f, ext = os.path.splitext(fname)
        if f[-5:] != "_info":
            fbase, fname = f, f"{f}_info"
        else:
            fbase, fname = f[:-5], f
        return f"{fname}.npz", f"{fbase}.npz"

    @staticmethod
    def get(hilbert, qubit_hamiltonian, hamiltonian_fname=None, restricted_idxs=None, n_excitations_max=None,
            verbose=False, dtype=np.float32, device=None):
        """Same arguments as the reference's factory (hamiltonian.py:48-62).  A saved Hamiltonian
        (`*_info.npz` + `.npz`) is loaded into the cache exactly like the reference does."""
        ph = PauliHamiltonianB200(hilbert, qubit_hamiltonian, restricted_idxs, n_excitations_max, verbose, dtype, device)
        if hamiltonian_fname is not None:
            info_name, npz_name = PauliHamiltonian._format_fnames(hamiltonian_fname)
            if os.path.exists(info_name):
                ph.load(info_name)
            elif os.path.exists(npz_name):
                ph.load_H(npz_name)
                ph.freeze_H()
        return ph


class PauliHamiltonianB200:
    # defaults for objects assembled without __init__ (tests drive the mirror from a packed table)
    track_seen = True
    _seen_chunks, _seen_count = (), 0

    def __init__(self, hilbert, qubit_hamiltonian, restricted_idxs=None, n_excitations_max=None, verbose=False,
                 dtype=np.float32, device=None):
        self.hilbert = hilbert
        assert self.hilbert.encoding.name == Encoding.SIGNED.name, "PauliCouplings requires Encoding.SIGNED."
        self.qubit_hamiltonian = qubit_hamiltonian
        self.restricted_idxs = self.hilbert.full2restricted_idx(restricted_idxs)
        self._restricted_full_idxs = None if restricted_idxs is None else self.hilbert.to_idx_array(restricted_idxs).reshape(-1)
        self.n_excitations_max = n_excitations_max
        self.dtype = dtype
        self.verbose = verbose

        N = self.hilbert.N
        n_alpha, n_beta = getattr(hilbert, "N_alpha", None), getattr(hilbert, "N_beta", None)
        if n_alpha is not None and np.ndim(n_alpha) > 0:
            raise NotImplementedError("partially restricted Hilbert spaces (several (N_alpha, N_beta) sectors) are not supported on the device path")
        xy, yz, c = pack_terms(qubit_hamiltonian.terms, N, n_occ=getattr(hilbert, "N_occ", 0) or 0,
                               n_excitations_max=n_excitations_max)
        # reference attributes (hamiltonian.py:246-252)
        idt = self.hilbert.get_idx_dtype("np") if hasattr(self.hilbert, "get_idx_dtype") else np.int64
        self.XY_sites_idx = xy[:, 0].view(np.int64).astype(idt)
        self.YZ_sites_idx = yz[:, 0].view(np.int64).astype(idt)
        self.couplings = c.astype(dtype).reshape(-1, 1)
        self._unique_XY_sites_idx, self._unique2all_XY_sites_idx = np.unique(self.XY_sites_idx, return_inverse=True)
        self._unique_YZ_sites_idx, self._unique2all_YZ_sites_idx = np.unique(self.YZ_sites_idx, return_inverse=True)
        # float64 (experiments/_base.py:234) or float32 (this constructor's default, as in the reference): the device holds
        # the coefficients the reference would hold in `dtype` and rounds every partial sum of H_ij to it
        if np.dtype(dtype) not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError(f"PauliHamiltonian on the device supports dtype float32 / float64, got {np.dtype(dtype)} (no device type).")
        self.table = DeviceTermTable(xy, yz, self.couplings.reshape(-1).astype(np.float64), N, n_alpha, n_beta, device)
        self.table.set_precision(dtype)

        d = self.hilbert.size
        self.H = csr_matrix(([], ([], [])), shape=(d, d), dtype=self.dtype)
        self._cached_idxs = np.array([], dtype=idt)
        self._frozen_H = False
        self._restricted_H = None
        # states that went through the fused (stateless) local_energy path but are not in the CSR cache: the reference's
        # update_H would have cached their rows as a side effect (energy.py:245), and its solve_H / save rely on that
        # (energy.py:559,777,447).  Their rows are computed on demand by get_H(idxs) / save(); see _materialize_seen.
        self.track_seen = True
        self._seen_chunks, self._seen_count = [], 0
        if verbose:
            print(f"Pauli Hamiltonian has K={self.table.K} terms, {self.table.Kxy} unique XY masks, {self.table.Kyz} unique YZ masks "
                  f"(device {self.table.device}).")

    # ------------------------------------------------------------------ fused path
    def local_energy(self, states_idx, psi, ret_numpy=True, assume_unique=True, track=None):
        """E_loc (complex128) of the sampled batch, only couplings inside the batch contribute
        (energy.py:247-248).  Stateless: nothing is cached.  assume_unique=True as in the reference's call
        update_H(states_idx, check_unseen=True, assume_unique=True) (energy.py:245)."""
        on_device = (torch.is_tensor(states_idx) and states_idx.is_cuda) or (torch.is_tensor(psi) and psi.is_cuda)
        if (self.track_seen if track is None else track) and not self._frozen_H:
            self._note_seen(states_idx)  # a CUDA batch is kept as a device tensor and only copied if get_H() / save() need it
        if REFERENCE_QUIRKS and self._is_full_sample(states_idx):
            # q1: the reference pairs psi[j] with the j-th sector state in RESTRICTED order here (see REFERENCE_QUIRKS)
            states_idx = self._restricted_full_idxs
        if ret_numpy and not on_device:
            # host-resident batch (the reference's situation: sampler output is moved to the CPU, nade.py:727-733):
            # one C-ABI call that uploads, computes and downloads (naqs_eloc_host)
            idx = states_idx.detach().numpy() if torch.is_tensor(states_idx) else np.asarray(states_idx)
            return self.table.local_energy_host(np.ascontiguousarray(idx.reshape(-1)), psi, assume_unique=assume_unique)
        out = self.table.local_energy(np.asarray(states_idx).reshape(-1) if not torch.is_tensor(states_idx) else states_idx.reshape(-1), psi,
                                      assume_unique=assume_unique)
        return _lib.complex_from_pairs(out) if ret_numpy else out

    def local_energy_full(self, states_idx, psi, psi_fn):
        """E_loc with the amplitudes of ALL coupled states, not only the sampled ones — the mode the reference leaves
        unimplemented (`set_unsampled_states_to_zero=False` raises NotImplementedError, energy.py:250-258).
        The sorted unique coupled set comes from the device (rows kernel + radix sort + unique, i.e.
        get_coupled_state_idxs(return_unique=True) of hamiltonian.py:122-132); `psi_fn(keys int64 [U]) -> complex [U]`
        evaluates the wavefunction on it (one batched network call); the fused kernel then runs against that table."""
        keys = np.asarray(states_idx).reshape(-1).astype(np.int64) if not torch.is_tensor(states_idx) else states_idx.reshape(-1).cpu().numpy().astype(np.int64)
        coupled = self.table.coupled_state_set(keys)[:, 0].cpu().numpy()
        psi_c = np.asarray(psi_fn(coupled))
        out = self.table.local_energy(keys, psi, table_keys=coupled, table_psi=psi_c, assume_unique=True)
        return _lib.complex_from_pairs(out)

    def linear_operator(self, states_idx):
        """scipy LinearOperator of H restricted to the given basis states, applied matrix-free on the device."""
        from scipy.sparse.linalg import LinearOperator
        keys = self.table._keys(np.asarray(states_idx).reshape(-1) if not torch.is_tensor(states_idx) else states_idx.reshape(-1))
        n = keys.shape[0]

        def matvec(v):
            v = np.asarray(v).reshape(-1)
            out = _lib.complex_from_pairs(self.table.apply_H(keys, v.astype(np.complex128)))
            return out if np.iscomplexobj(v) else out.real

        return LinearOperator((n, n), matvec=matvec, dtype=np.float64)

    def solve_H(self, states_idx, k=1, tol=1e-10):
        """Lowest eigenpairs of H on the sub-space spanned by `states_idx` (what OptimizerBase.solve_H computes with
        scipy eigs on the cached CSR, energy.py:762-786) by Lanczos on the matrix-free operator."""
        from scipy.sparse.linalg import eigsh
        return eigsh(self.linear_operator(states_idx), k=k, which="SA", tol=tol)

    def _is_full_sample(self, idxs):
        r = getattr(self, "_restricted_full_idxs", None)
        try:
            return r is not None and len(idxs) == len(r)
        except TypeError:
            return False

    # ------------------------------------------------------------------ reference API
    def _seen_host(self):
        return [self.hilbert.to_idx_array(c.cpu() if torch.is_tensor(c) else c).reshape(-1) for c in self._seen_chunks]

    def _note_seen(self, states_idx):
        if torch.is_tensor(states_idx) and states_idx.is_cuda:
            a = states_idx.detach().reshape(-1).clone()
        else:
            a = self.hilbert.to_idx_array(states_idx).reshape(-1).copy()
        self._seen_chunks = list(self._seen_chunks) + [a]
        self._seen_count += len(a)
        if len(self._seen_chunks) > 64 or self._seen_count > 4 * max(len(self._cached_idxs), 1 << 16):
            u = np.unique(np.concatenate(self._seen_host()))
            self._seen_chunks, self._seen_count = [u], len(u)

    def _materialize_seen(self):
        """Rows of every state that went through the fused path since the last call -> CSR cache (what the reference's
        update_H side effect at energy.py:245 would have produced)."""
        if self._seen_chunks and not self._frozen_H:
            seen = np.unique(np.concatenate(self._seen_host()))
            self._seen_chunks, self._seen_count = [], 0
            self.update_H(seen, check_unseen=True, assume_unique=True)

    def _rows_csr(self, state_i_idx):
        indptr, cols, ridx, vals = self.table.rows(state_i_idx.astype(np.int64), with_restricted_index=True)
        indptr, ridx, vals = indptr.cpu().numpy(), ridx.cpu().numpy(), vals.cpu().numpy()
        i_idx = np.asarray(self.hilbert.full2restricted_idx(state_i_idx)).astype(np.int64)
        rows = np.repeat(i_idx, np.diff(indptr))
        return csr_matrix((vals.astype(self.dtype), (rows, ridx)), shape=(self.hilbert.size, self.hilbert.size))

    def update_H(self, state_idx, check_unseen=True, assume_unique=False):
        """hamiltonian.py:272-370: add the rows of not-yet-cached states to the sparse H and return it."""
        if not self._frozen_H:
            state_i_idx = self.hilbert.to_idx_array(state_idx).reshape(-1)
            if check_unseen:
                state_i_idx = np.setdiff1d(state_i_idx, self._cached_idxs, assume_unique=assume_unique)
                if len(state_i_idx) == 0:
                    return self.get_H()
            H_new = self._rows_csr(state_i_idx)
            self._cached_idxs = np.concatenate((self._cached_idxs, state_i_idx))
            self.H = self.H + H_new
        return self.H

    def __get_new_H_subspace(self, idxs):
        idxs = np.asarray(idxs).reshape(-1)
        return self.H[idxs[:, np.newaxis], idxs]

    def get_H(self, idxs=None):
        """hamiltonian.py:96-111.  (The reference's full-sector shortcut returns rows in restricted order even
        when `idxs` is permuted — quirk q1 of SURVEY.md §8a; here the caller's order is kept unless REFERENCE_QUIRKS is set.)"""
        if idxs is not None:
            idxs = idxs.detach().cpu().numpy() if torch.is_tensor(idxs) else np.asarray(idxs)
            if not self._frozen_H:
                # the fused local_energy path caches nothing, so the rows the reference would find in its cache (filled by
                # the update_H of calculate_local_energy, energy.py:245) are computed here on demand — solve_H
                # (energy.py:559,777) calls get_H without update_H
                self.update_H(idxs.reshape(-1), check_unseen=True)
            if REFERENCE_QUIRKS and self._is_full_sample(idxs):
                return self.get_restricted_H()  # hamiltonian.py:100-105 (q1)
            return self.__get_new_H_subspace(self.hilbert.full2restricted_idx(idxs))
        self._materialize_seen()
        return self.H

    def get_restricted_H(self):
        if self._frozen_H and (self._restricted_H is not None):
            return self._restricted_H
        r = self.restricted_idxs
        H = self.__get_new_H_subspace(r.detach().cpu().numpy() if torch.is_tensor(r) else r)
        if self._frozen_H:
            self._restricted_H = H
        return H

    def get_coupled_state_idxs(self, state_idxs, return_unique=False):
        """hamiltonian.py:122-132 (rows of the cached H, i.e. restricted indices)."""
        coupled = [self.H.indices[self.H.indptr[i]:self.H.indptr[i + 1]] for i in state_idxs]
        if return_unique:
            coupled = np.unique(np.concatenate(coupled))
        return coupled

    def coupled_state_set(self, states_idx):
        """Device version of get_coupled_state_idxs(return_unique=True) for FULL state indices, without the cache:
        sorted unique coupled keys (radix sort + unique on the GPU)."""
        keys = self.table.coupled_state_set(np.asarray(states_idx).reshape(-1))
        return keys[:, 0].cpu().numpy()

    def freeze_H(self):
        self._frozen_H = True

    def unfreeze_H(self):
        self._frozen_H = False
        self._restricted_H = None

    def is_frozen(self):
        return self._frozen_H

    # ------------------------------------------------------------------ npz persistence (hamiltonian.py:146-198)
    def load_H(self, fname):
        if os.path.splitext(fname)[-1] != ".npz":
            fname += ".npz"
        self.H = load_npz(fname)

    def save_H(self, fname):
        if os.path.splitext(fname)[-1] != ".npz":
            fname += ".npz"
        d = os.path.dirname(fname)
        if d:
            os.makedirs(d, exist_ok=True)
        data = self.H.data
        data[np.abs(self.H.data) < 1e-12] = 0
        self.H.data = data
        self.H.eliminate_zeros()
        save_npz(fname, self.H)

    def load(self, fname, fname_H=None):
        fname, fname_H = PauliHamiltonian._format_fnames(fname, fname_H)
        with np.load(fname, allow_pickle=True) as f_in:
            info = f_in["info"]
        self._cached_idxs = info[0]
        self._frozen_H = info[1]
        self.load_H(fname_H)

    def save(self, fname, fname_H=None):
        self._materialize_seen()  # the reference saves the rows of every state seen so far (energy.py:447)
        fname, fname_H = PauliHamiltonian._format_fnames(fname, fname_H)
        d = os.path.dirname(fname)
        if d:
            os.makedirs(d, exist_ok=True)
        info = np.array([self._cached_idxs, self._frozen_H, fname_H], dtype=object)
        np.savez(fname, info=info)
        self.save_H(fname_H)
