"""Pauli-sum pre-processing: Pauli strings -> packed (XY, YZ, coefficient) term table.

Mirror of _PauliHamiltonianDynamic.__calc_coupling_info (src/optimizer/hamiltonian.py:373-430) and an
openfermion-free reader for the reference's `<mol>_qubit_hamiltonian.pkl` files
(src/utils/system.py:29-33).  Host-side, once per run.
"""
import io
import pickle

import numpy as np


class QubitOperatorData:
    """What the reference's pickles hold: an object whose only attribute is
    terms = {((qubit, 'X'|'Y'|'Z'), ...): complex coefficient} (openfermion 0.11 QubitOperator)."""

    terms = None

    def many_body_order(self):
        return max((len(t) for t in self.terms), default=0)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("openfermion") and name == "QubitOperator":
            return QubitOperatorData
        return super().find_class(module, name)


def load_qubit_hamiltonian(path):
    """Read a `<mol>_qubit_hamiltonian.pkl` without openfermion -> object with `.terms`."""
    with open(path, "rb") as f:
        return _Unpickler(io.BytesIO(f.read())).load()


def pack_terms(terms, n_qubits, n_occ=0, n_excitations_max=None):
    """-> (xy [K, W] uint64, yz [K, W] uint64, coeff [K] float64) in the dict's iteration order.

    Per term (hamiltonian.py:383-416): XY bit for X or Y, YZ bit for Y or Z; the term is dropped if it
    flips a frozen qubit (q < n_occ) or flips more than n_excitations_max qubits; coefficient
    Re(i^nY) * coeff with the imaginary part discarded by the float cast (hamiltonian.py:416,424) —
    odd-nY terms therefore carry +-0.0, exactly as in the reference."""
    W = 1 if n_qubits <= 63 else 2
    items = terms.items() if hasattr(terms, "items") else terms
    xy, yz, cs = [], [], []
    for term, coeff in items:
        mx = mz = 0
        n_y = n_flip = 0
        valid = True
        for q, p in term:
            if p == "X" or p == "Y":
                mx |= 1 << q
                if p == "Y":
                    n_y += 1
                    mz |= 1 << q
                if q < n_occ:
                    valid = False
                    break
                if n_excitations_max is not None:
                    n_flip += 1
                    if n_flip > n_excitations_max:
                        valid = False
                        break
            elif p == "Z":
                mz |= 1 << q
        if valid:
            xy.append(mx)
            yz.append(mz)
            cs.append(complex((1j ** n_y).real * coeff).real)
    K = len(cs)
    out_xy = np.zeros((K, W), np.uint64)
    out_yz = np.zeros((K, W), np.uint64)
    m64 = (1 << 64) - 1
    for w in range(W):
        out_xy[:, w] = np.array([(v >> (64 * w)) & m64 for v in xy], dtype=np.uint64) if K else 0
        out_yz[:, w] = np.array([(v >> (64 * w)) & m64 for v in yz], dtype=np.uint64) if K else 0
    return out_xy, out_yz, np.array(cs, np.float64)
