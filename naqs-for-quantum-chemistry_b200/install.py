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
    """Device-resident hand-off from the sampler to the E_loc path: whatever `wavefunction.sample` returns
    (states, counts, probs, log_psi) is put on the model's device before the loop sees it, and the two Hilbert helpers the
    loop applies to the samples accept CUDA tensors — state2idx packs the int8 rows with the device kernel (naqs_state2idx,
    hilbert.py:573-581), to_idx_array copies to the host only where the reference needs numpy (energy.py:300).  From there
    calculate_local_energy / sgd_step stay on the device.
    (The reference's NADE cannot simply be given out_device="cuda" — network/base.py:29,39: its sampler builds index tensors
    on the host, nade.py:585 — so its per-orbital outputs still pass through out_device; that is sampler code, outside this path.)"""
    import functools

    import torch

    from . import _lib
    wf = importlib.import_module("src.naqs.wavefunction")
    hil = importlib.import_module("src.utils.hilbert")

    def to_model_device(sample):
        @functools.wraps(sample)
        def wrapped(self, *a, **k):
            out = sample(self, *a, **k)
            dev = getattr(self, "device", "cpu")
            if not isinstance(out, (list, tuple)) or not torch.cuda.is_available() or str(dev) == "cpu":
                return out
            return [o.to(dev, non_blocking=True) if torch.is_tensor(o) else o for o in out]
        return wrapped

    for name in dir(wf):
        cls = getattr(wf, name)
        if isinstance(cls, type) and cls.__module__ == wf.__name__ and "sample" in cls.__dict__ and name != "_NAQSComplex_Base":
            cls.sample = to_model_device(cls.sample)

    def patch_state2idx(cls):
        state2idx_host = cls.state2idx

        @functools.wraps(state2idx_host)
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

        cls.state2idx = state2idx

    def patch_to_idx_array(cls):
        to_idx_array_host = cls.to_idx_array

        @functools.wraps(to_idx_array_host)
        def to_idx_array(self, idx):
            if torch.is_tensor(idx) and idx.is_cuda:
                idx = idx.cpu()
            return to_idx_array_host(self, idx)

        cls.to_idx_array = to_idx_array

    # state2idx is defined per Hilbert flavour (hilbert.py:344, 573, 833; abstract at :51), to_idx_array once in the base (:85)
    for name in dir(hil):
        cls = getattr(hil, name)
        if not isinstance(cls, type) or cls.__module__ != hil.__name__:
            continue
        if "state2idx" in cls.__dict__ and not getattr(cls.__dict__["state2idx"], "__isabstractmethod__", False):
            patch_state2idx(cls)
        if "to_idx_array" in cls.__dict__:
            patch_to_idx_array(cls)


def install(reference_root=None, patch_level0=True, patch_level1=True, reference_quirks=None, device_resident=None, fused_loss=None):
    """reference_quirks=True reproduces quirk q1 (hamiltonian.REFERENCE_QUIRKS) for bitwise parity with the reference on
    full-sector batches; None keeps the NAQS_ELOC_REFERENCE_QUIRKS environment setting (default off).
    device_resident=True (or NAQS_ELOC_DEVICE_RESIDENT=1): the sampler's output is handed to the loop on the GPU and state2idx,
    E_loc and the loss statistics are computed from it without a host round trip (_patch_device_resident; implies fused_loss).
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
        if device_resident is None:
            device_resident = os.environ.get("NAQS_ELOC_DEVICE_RESIDENT", "0") not in ("", "0")
        if fused_loss is None:
            fused_loss = os.environ.get("NAQS_ELOC_FUSED_LOSS", "0") not in ("", "0")
        if fused_loss or device_resident:  # the reference's _SGD_step mixes host and device tensors; the fused step does not
            ref_e.OptimizerBase._SGD_step = _energy.sgd_step
    if device_resident:
        _patch_device_resident(reference_root)
    return True
