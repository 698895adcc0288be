"""install() — make the UNMODIFIED reference tree use the B200 path.

After `install(reference_root)`:
  * `import src.utils.hamiltonian_math / sparse_math / hilbert_math` resolve to the device-backed Level-0
    twins in this package (the names fixed by src_cpp/setup.py:36-38), so the reference imports
    (hamiltonian.py:13, energy.py:27, hilbert.py:14) work without the Cython build;
  * `src.optimizer.hamiltonian.PauliHamiltonian.get` returns a PauliHamiltonianB200;
  * `OptimizerBase.calculate_local_energy` is the fused device version.
experiments/run.py and src/naqs then run unchanged.  The switch NAQS_ELOC_BACKEND=reference leaves the
reference path untouched (both can run in one process for parity / timing).
"""
import importlib
import os
import sys

from . import energy as _energy
from . import hamiltonian as _hamiltonian
from . import hamiltonian_math, hilbert_math, sparse_math


def install(reference_root=None, patch_level0=True, patch_level1=True, reference_quirks=None):
    """reference_quirks=True reproduces quirk q1 (hamiltonian.REFERENCE_QUIRKS) for bitwise parity with the reference on
    full-sector batches; None keeps the NAQS_ELOC_REFERENCE_QUIRKS environment setting (default off)."""
    if os.environ.get("NAQS_ELOC_BACKEND", "b200") == "reference":
        return False
    if reference_quirks is not None:
        _hamiltonian.REFERENCE_QUIRKS = bool(reference_quirks)
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    if patch_level0:
        for name, mod in (("hamiltonian_math", hamiltonian_math), ("sparse_math", sparse_math), ("hilbert_math", hilbert_math)):
            sys.modules[f"src.utils.{name}"] = mod
            pkg = sys.modules.get("src.utils")
            if pkg is not None:
                setattr(pkg, name, mod)
    if patch_level1:
        ref_h = importlib.import_module("src.optimizer.hamiltonian")
        ref_h.PauliHamiltonian.get = staticmethod(_hamiltonian.PauliHamiltonian.get)
        ref_e = importlib.import_module("src.optimizer.energy")
        ref_e.PauliHamiltonian = ref_h.PauliHamiltonian
        ref_e.OptimizerBase.calculate_local_energy = _energy.calculate_local_energy
    return True
