"""naqs-for-quantum-chemistry_b200 — B200-native local-energy (E_loc) hot path of NAQS.

Hand-written sm_100a CUDA behind the C ABI of include/naqs_eloc.h (csrc/), plus the host-side mirror
of the reference's interface for this path: the three compiled modules
(hamiltonian_math / sparse_math / hilbert_math), PauliHamiltonian, calculate_local_energy and the
Hilbert encodings.  Importable as `naqs_b200` (alias package at the repository root).
No CPU fallback: compute entries raise without the CUDA library / a CUDA device.
"""
from . import _lib  # noqa: F401
from ._lib import NaqsError, launch_count  # noqa: F401
from . import distributed  # noqa: F401
from .energy import calculate_local_energy, local_energy_statistics, stats_from_sums  # noqa: F401
from .hamiltonian import PauliHamiltonian, PauliHamiltonianB200  # noqa: F401
from .hilbert import Encoding, Hilbert  # noqa: F401
from .install import install  # noqa: F401
from .pauli import load_qubit_hamiltonian, pack_terms  # noqa: F401
from .table import DeviceTermTable  # noqa: F401

__all__ = ["DeviceTermTable", "PauliHamiltonian", "PauliHamiltonianB200", "Hilbert", "Encoding", "calculate_local_energy",
           "local_energy_statistics", "stats_from_sums", "pack_terms", "load_qubit_hamiltonian", "install", "NaqsError",
           "launch_count"]
