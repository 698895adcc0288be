"""Local-energy entry points — mirror of the hot-path methods of src/optimizer/energy.py.

`calculate_local_energy` has the signature of OptimizerBase.calculate_local_energy (energy.py:219-263)
and is what `install()` binds onto the reference's optimizer classes; `local_energy_statistics`
computes the scalars of energy.py:328,372-375 in fp64 on the device.
"""
import numpy as np
import torch

from . import _lib


def np_to_torch(x):
    """src/utils/complex.py:139-140: complex numpy -> float32 [..., 2] torch (the reference truncates here)."""
    return torch.FloatTensor(np.stack([np.real(x), np.imag(x)], -1))


def calculate_local_energy(self, states_idx, psi=None, set_unsampled_states_to_zero=True, ret_complex=False):
    """Drop-in for OptimizerBase.calculate_local_energy (energy.py:219-263).

    `self` is the optimizer (needs .pauli_hamiltonian with a `.local_energy` — i.e. a PauliHamiltonianB200 —
    and, when psi is None, .wavefunction / .hilbert).  psi: complex torch [M, 2] (float32 in the reference's
    loop) or complex numpy.  Returns float32 [M, 2] torch, or complex128 numpy if ret_complex — the same
    outputs as the reference (complex.py:139-140), computed by the fused sm_100a kernel."""
    with torch.no_grad():
        if psi is None:
            psi = self.wavefunction.psi(self.hilbert.idx2state(states_idx, use_restricted_idxs=False), ret_complex=True)
        elif torch.is_tensor(psi):
            psi = psi.detach()
        if not set_unsampled_states_to_zero:
            raise NotImplementedError()  # energy.py:250-251
        idx = states_idx.detach().cpu().numpy() if torch.is_tensor(states_idx) else np.asarray(states_idx)
        local_energy = self.pauli_hamiltonian.local_energy(idx.reshape(-1), psi, ret_numpy=True)
        if not ret_complex:
            local_energy = np_to_torch(local_energy)
        return local_energy


def local_energy_statistics(table, eloc, weights=None):
    """-> dict(sum_w, mean (complex), variance of Re E) from the five all-reducible fp64 sums
    [sum w, sum w Re E, sum w Im E, sum w (Re E)^2, n] (energy.py:328,372-375)."""
    s = table.stats(eloc, weights).cpu().numpy()
    return stats_from_sums(s)


def stats_from_sums(s):
    sw = s[0]
    mean_re, mean_im = s[1] / sw, s[2] / sw
    return {"sum_w": float(sw), "mean": complex(mean_re, mean_im), "variance": float(s[3] / sw - mean_re ** 2), "n": int(round(s[4]))}
