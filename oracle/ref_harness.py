"""Drive the UNMODIFIED reference code as the primary oracle — TEST INFRASTRUCTURE, NOT PRODUCT.

Works only where /root/reference exists (the build container).  It imports the
reference's own Python (src/optimizer/hamiltonian.py, src/utils/hilbert.py, ...) straight
from /root/reference, with the reference's Cython kernels compiled by oracle/build_ref.py
into oracle/_ref/.  Nothing under /root/reference is modified or copied; the environment
drift listed in SURVEY.md §8c is bridged with sys.modules stubs only.

Used by tests/golden/make_golden.py (fixture generation) and by the not-gpu tests that
re-validate the restatement live when the reference is present.
"""
import importlib
import math
import os
import pickle
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_KERNELS = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("NAQS_REFERENCE_ROOT", "/root/reference")

# (n_qubits, n_alpha, n_beta): the .hdf5 metadata the harness cannot read without h5py
# (electrons from the molecules' closed-shell STO-3G / 6-31G setups).
MOLECULES = {
    "H2": (4, 1, 1), "LiH": (12, 2, 2), "H2O": (14, 5, 5), "BeH2": (14, 3, 3), "NH3": (16, 5, 5),
    "CH4": (18, 5, 5), "N2": (20, 7, 7), "C2": (20, 6, 6), "F2": (20, 9, 9), "HCl": (20, 9, 9),
    "LiF": (20, 6, 6), "H2S": (22, 9, 9), "PH3": (24, 9, 9), "H2O_6-31G": (26, 5, 5),
    "LiCl": (28, 10, 10), "H4O2": (28, 10, 10), "Li2O": (30, 7, 7),
}
for _g in ("0.75", "0.9", "1.05", "1.2", "1.35", "1.5", "1.65", "1.8", "1.95", "2.1", "2.25"):
    MOLECULES[f"N2_{_g}"] = (20, 7, 7)


class QubitOperator:
    """Stand-in for openfermion.ops._qubit_operator.QubitOperator: the pickles hold a plain
    object whose __dict__ is {'terms': {((q, 'X'|'Y'|'Z'), ...): complex128}}."""

    def many_body_order(self):
        return max((len(t) for t in self.terms), default=0)


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "optimizer"))


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_stubs():
    _stub("openfermion")
    _stub("openfermion.hamiltonians", MolecularData=type("MolecularData", (), {}))
    _stub("openfermion.transforms", get_fermion_operator=None, jordan_wigner=None)
    _stub("openfermion.ops")
    _stub("openfermion.ops._qubit_operator", QubitOperator=QubitOperator)
    import torch  # noqa: F401
    _stub("torch._six", inf=math.inf)
    if not hasattr(np, "long"):
        np.long = np.int64


def load_terms(molecule, reference_root=None):
    """-> dict of Pauli terms in the pickle's (insertion) order."""
    install_stubs()
    root = reference_root or REFERENCE_ROOT
    path = os.path.join(root, "molecules", molecule, f"{molecule}_qubit_hamiltonian.pkl")
    with open(path, "rb") as f:
        op = pickle.load(f)
    return op.terms


def load_operator(molecule):
    op = QubitOperator()
    op.terms = load_terms(molecule)
    return op


_ref_modules = None


def reference_modules():
    """Import the reference's src.* packages (from /root/reference) with oracle/_ref kernels."""
    global _ref_modules
    if _ref_modules is not None:
        return _ref_modules
    if not available():
        raise RuntimeError("reference tree not present")
    sys.path.insert(0, HERE)
    import build_ref
    if not build_ref.build(REFERENCE_ROOT):
        raise RuntimeError("could not build oracle/_ref")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    src_utils = importlib.import_module("src.utils")
    kp = os.path.join(REF_KERNELS, "src", "utils")
    if kp not in src_utils.__path__:
        src_utils.__path__.append(kp)
    mods = types.SimpleNamespace(
        hamiltonian_math=importlib.import_module("src.utils.hamiltonian_math"),
        sparse_math=importlib.import_module("src.utils.sparse_math"),
        hilbert_math=importlib.import_module("src.utils.hilbert_math"),
        hilbert=importlib.import_module("src.utils.hilbert"),
        hamiltonian=importlib.import_module("src.optimizer.hamiltonian"),
    )
    _ref_modules = mods
    return mods


def reference_kernels_only():
    """Import just the compiled reference kernels from oracle/_ref (works on the GPU box too)."""
    kp = os.path.join(REF_KERNELS)
    pkg = _stub("naqs_refk")
    pkg.__path__ = [os.path.join(kp, "src", "utils")]
    if not hasattr(np, "long"):
        np.long = np.int64
    import torch  # noqa: F401  (sparse_math.pyx:6 imports torch)
    return types.SimpleNamespace(
        hamiltonian_math=importlib.import_module("naqs_refk.hamiltonian_math"),
        sparse_math=importlib.import_module("naqs_refk.sparse_math"),
        hilbert_math=importlib.import_module("naqs_refk.hilbert_math"),
    )


def make_reference(molecule, restricted=True, dtype=np.float64):
    """-> (hilbert, pauli_hamiltonian) built by the reference's own classes
    (experiments/_base.py:113-123, src/optimizer/energy.py:115-125)."""
    mods = reference_modules()
    N, na, nb = MOLECULES[molecule]
    H = mods.hilbert
    if restricted:
        hilbert = H.Hilbert.get(N, na, nb, encoding=H.Encoding.SIGNED, make_basis=True)
        idxs = hilbert.get_subspace(ret_states=False, ret_idxs=True).numpy()
    else:
        hilbert = H.Hilbert.get(N, encoding=H.Encoding.SIGNED, make_basis=False)
        idxs = None
    ph = mods.hamiltonian.PauliHamiltonian.get(hilbert, load_operator(molecule), restricted_idxs=idxs, dtype=dtype)
    return hilbert, ph


def reference_local_energy(ph, states_idx, psi):
    """energy.py:245-248 verbatim: update_H(check_unseen=True, assume_unique=True);
    (sparse_dense_mv(get_H(idx), psi) / psi).conj().  states_idx: numpy int array."""
    mods = reference_modules()
    ph.update_H(states_idx, check_unseen=True, assume_unique=True)
    return (mods.sparse_math.sparse_dense_mv(ph.get_H(states_idx), psi) / psi).conj()
