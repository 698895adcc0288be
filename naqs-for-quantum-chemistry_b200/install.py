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


def _patch_device_resident(reference_root):
    """Keep the sampler's output on the GPU: the networks' and wavefunctions' out_device default ("cpu",
    network/base.py:29,39 and wavefunction.py:27,38-42) becomes the model device, and the two Hilbert helpers the loop
    applies to the samples accept CUDA tensors: state2idx packs the int8 rows with the device kernel (naqs_state2idx,
    hilbert.py:573-581), to_idx_array copies to the host only where the reference needs numpy (energy.py:300)."""
    import functools

    import torch

    from . import _lib
    nb = importlib.import_module("src.naqs.network.base")
    wf = importlib.import_module("src.naqs.wavefunction")
    hil = importlib.import_module("src.utils.hilbert")

    def default_out_device(init):
        @functools.wraps(init)
        def wrapped(self, *a, **k):
            if k.get("out_device", "cpu") == "cpu" and torch.cuda.is_available() and k.get("device", None) in (None, "cuda"):
                k["out_device"] = "cuda"
            return init(self, *a, **k)
        return wrapped

    nb.ComplexAutoregressiveMachine_Base.__init__ = default_out_device(nb.ComplexAutoregressiveMachine_Base.__init__)
    wf._NAQSComplex_Base.__init__ = default_out_device(wf._NAQSComplex_Base.__init__)

    def patch_hilbert(cls):
        state2idx_host, to_idx_array_host = cls.state2idx, cls.to_idx_array

        def state2idx(self, state, use_restricted_idxs=False):
            if not (torch.is_tensor(state) and state.is_cuda):
                return state2idx_host(self, state, use_restricted_idxs)
            rows = state.to(torch.int8).reshape(-1, state.shape[-1]).contiguous()
            words = _lib.n_words(self.N)
            keys = torch.empty((rows.shape[0], words), dtype=torch.int64, device=rows.device)
            with torch.cuda.device(rows.device):
                _lib.check(_lib.load().naqs_state2idx(_lib.ptr(rows), rows.shape[0], self.N, words, _lib.ptr(keys), _lib.stream_ptr(rows.device)),
                           "naqs_state2idx")
            idxs = self.to_idx_tensor(keys[:, :1])
            if use_restricted_idxs:
                idxs = self.to_idx_tensor(self.full2restricted_idx(idxs.cpu())).to(rows.device)
            return idxs

        def to_idx_array(self, idx):
            if torch.is_tensor(idx) and idx.is_cuda:
                idx = idx.cpu()
            return to_idx_array_host(self, idx)

        cls.state2idx, cls.to_idx_array = state2idx, to_idx_array

    for name in dir(hil):
        cls = getattr(hil, name)
        if isinstance(cls, type) and "state2idx" in cls.__dict__ and "to_idx_array" in cls.__dict__:
            patch_hilbert(cls)


def install(reference_root=None, patch_level0=True, patch_level1=True, reference_quirks=None, device_resident=None, fused_loss=None):
    """reference_quirks=True reproduces quirk q1 (hamiltonian.REFERENCE_QUIRKS) for bitwise parity with the reference on
    full-sector batches; None keeps the NAQS_ELOC_REFERENCE_QUIRKS environment setting (default off).
    device_resident=True (or NAQS_ELOC_DEVICE_RESIDENT=1): the sampler's states / log_psi stay on the GPU and E_loc is computed
    from them without a host round trip (_patch_device_resident).
    fused_loss=True (or NAQS_ELOC_FUSED_LOSS=1): OptimizerBase._SGD_step is energy.sgd_step — the loss statistics of
    energy.py:316-329, 367-375 come from one fp64 device kernel and autograd receives a detached weight vector."""
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
        if fused_loss is None:
            fused_loss = os.environ.get("NAQS_ELOC_FUSED_LOSS", "0") not in ("", "0")
        if fused_loss:
            ref_e.OptimizerBase._SGD_step = _energy.sgd_step
    if device_resident is None:
        device_resident = os.environ.get("NAQS_ELOC_DEVICE_RESIDENT", "0") not in ("", "0")
    if device_resident:
        _patch_device_resident(reference_root)
    return True
