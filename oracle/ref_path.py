"""The reference's E_loc CPU path, runnable WITHOUT /root/reference — TEST INFRASTRUCTURE, NOT PRODUCT.

The arithmetic kernels are the reference's own compiled Cython modules from oracle/_ref/
(popcount_parity, get_Hij_cy, sparse_dense_mv — built by oracle/build_ref.py from the
sources under /root/reference/src_cpp, shipped to the GPU box as .so).  The numpy / scipy
orchestration between them is restated here call for call from

    src/optimizer/hamiltonian.py:241-252   (np.unique groupings)
    src/optimizer/hamiltonian.py:272-370   (update_H)
    src/optimizer/hamiltonian.py:93-111    (get_H sub-matrix)
    src/optimizer/energy.py:245-248        (calculate_local_energy)

so that bench.py can time "the reference Cython path on the GPU box's own host cores"
 where the reference's Python tree is absent.  The full->restricted LUT
(hilbert.py:429-434) is the same 2^N-entry array gather the reference performs for N <= 30.
"""
import numpy as np
from scipy.sparse import csr_matrix

from . import eloc_oracle as eo
from . import ref_harness


class ReferencePath:
    def __init__(self, xy, yz, coeff, n_qubits, n_alpha=None, n_beta=None):
        k = ref_harness.reference_kernels_only()
        self.get_Hij_cy = k.hamiltonian_math.get_Hij_cy
        self.popcount_parity = k.hamiltonian_math.popcount_parity
        self.sparse_dense_mv = k.sparse_math.sparse_dense_mv
        assert n_qubits <= 62, "the reference path holds keys in int64 (hilbert.py:405-410)"
        self.N = n_qubits
        self.idt = eo.idx_dtype(n_qubits)
        self.XY = np.asarray(xy).reshape(-1).astype(np.int64).astype(self.idt)
        self.YZ = np.asarray(yz).reshape(-1).astype(np.int64).astype(self.idt)
        self.couplings = np.asarray(coeff, np.float64).reshape(-1, 1)
        self.uXY, self.u2aXY = np.unique(self.XY, return_inverse=True)          # :248
        self.uYZ, self.u2aYZ = np.unique(self.YZ, return_inverse=True)          # :249
        self.uXY = self.uXY.astype(self.idt)
        self.u2aXY = self.u2aXY.astype(self.idt)
        if n_alpha is None:
            self.size = 2 ** n_qubits
            self.lut = None
        else:
            sec = eo.sector_keys(n_qubits, n_alpha, n_beta)[:, 0].astype(np.int64)
            self.size = len(sec)
            lut = -1 * np.ones(2 ** n_qubits)                                    # hilbert.py:432
            lut[sec] = np.arange(len(sec))
            self.lut = lut.astype(self.idt)
        self.reset()

    def reset(self):
        """Cold cache (hamiltonian.py:86-88)."""
        self.H = csr_matrix(([], ([], [])), shape=(self.size, self.size), dtype=np.float64)
        self.cached = np.array([], dtype=self.idt)

    def full2restricted(self, idx):
        return idx if self.lut is None else self.lut[idx.astype(np.int64)]

    def update_H(self, state_idx, check_unseen=True, assume_unique=False):
        s = np.asarray(state_idx).astype(self.idt)
        if check_unseen:
            s = np.setdiff1d(s, self.cached, assume_unique=assume_unique)       # :294
            if len(s) == 0:
                return self.H
        P_bits = np.bitwise_and(s[:, None], self.uYZ[None, :])                  # :301
        P = self.popcount_parity(P_bits)                                        # :305
        Kxy = len(self.uXY)
        j_full = np.bitwise_xor(s[:, None], self.uXY[None, :]).ravel()          # :313
        i_idx = self.full2restricted(s)                                         # :321
        j_idx = self.full2restricted(j_full)                                    # :322
        mask = np.where(j_idx >= 0)[0]                                          # :328
        H_ij = self.get_Hij_cy(s, self.uXY, self.u2aXY, P, self.u2aYZ, self.couplings.squeeze())  # :335
        H_ij = H_ij[mask]                                                       # :343
        i_idx, j_idx = i_idx[mask // Kxy], j_idx[mask]                          # :344
        H_new = csr_matrix((H_ij, (i_idx, j_idx)), shape=(self.size, self.size))  # :350
        self.cached = np.concatenate((self.cached, s))                          # :357
        self.H = self.H + H_new                                                 # :363
        return self.H

    def get_H(self, idxs):
        r = self.full2restricted(np.asarray(idxs).astype(self.idt))
        return self.H[r[:, np.newaxis], r]                                      # :94

    def local_energy(self, states_idx, psi):
        """energy.py:245-248."""
        self.update_H(states_idx, check_unseen=True, assume_unique=True)
        return (self.sparse_dense_mv(self.get_H(states_idx), psi) / psi).conj()
