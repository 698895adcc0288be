"""Run the reference's own VMC loop (experiments/_base._run) — TEST INFRASTRUCTURE, NOT PRODUCT.

The UNMODIFIED reference tree is imported from `tree_root()` — /root/reference in the build container, or the
git-ignored staged copy oracle/_ref/tree that oracle/build_ref.py makes so that the loop can also run on the GPU box
(where /root/reference does not exist).  Two backends, selected by NAQS_ELOC_BACKEND exactly as a user would:

  reference : the reference's compiled Cython kernels (oracle/_ref/src/utils/*.so) — the host path as shipped
  b200      : naqs_b200.install() — Level-0 modules, PauliHamiltonian.get and calculate_local_energy replaced by the
              sm_100a path; experiments/_base.py, src/optimizer/energy.py and src/naqs run unchanged

Environment drift bridged with stubs only: openfermion / h5py absent (MolecularData stub
carrying n_qubits and electron counts), matplotlib / seaborn absent (plot_training stub), scipy.random gone
(set_global_seed without the scipy line), torch._six gone, and — reference backend only — torch index tensors passed
to numpy-only get_H / get_restricted_H.

`run_vmc` records, per iteration: the sampled state indices, the E_loc tensor handed to the loss, and the wall-clock
split sample / state2idx / E_loc / backward+step (energy.py:996-1003, 310-377).

CLI (used by the tests and by bench.py through a subprocess so that both backends can run side by side):
    python -m oracle.ref_vmc --molecule LiH --iters 20 --seed 111 --backend b200 --out /tmp/x.npz
"""
import argparse
import importlib
import json
import os
import random
import shutil
import sys
import tempfile
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGED = os.path.join(HERE, "_ref", "tree")
STAGED_MOLECULES = ("LiH", "H2O", "N2", "Li2O")


def tree_root():
    """The reference tree to import: NAQS_REFERENCE_ROOT, else /root/reference, else the staged copy."""
    for c in (os.environ.get("NAQS_REFERENCE_ROOT"), "/root/reference", STAGED):
        if c and os.path.isdir(os.path.join(c, "src", "optimizer")) and os.path.isdir(os.path.join(c, "experiments")):
            return c
    return None


def stage_tree(reference="/root/reference", force=False):
    """Copy src/, experiments/ and the qubit-Hamiltonian pickles of STAGED_MOLECULES into oracle/_ref/tree (git-ignored;
    travels to the GPU box with the snapshot like the compiled kernels next to it).  Nothing is edited."""
    if os.path.isdir(os.path.join(STAGED, "src", "optimizer")) and not force:
        return True
    if not os.path.isdir(os.path.join(reference, "src", "optimizer")):
        return False
    shutil.rmtree(STAGED, ignore_errors=True)
    os.makedirs(STAGED, exist_ok=True)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "bash")
    for d in ("src", "experiments"):
        shutil.copytree(os.path.join(reference, d), os.path.join(STAGED, d), ignore=ignore)
    for m in STAGED_MOLECULES:
        os.makedirs(os.path.join(STAGED, "molecules", m), exist_ok=True)
        shutil.copy(os.path.join(reference, "molecules", m, f"{m}_qubit_hamiltonian.pkl"), os.path.join(STAGED, "molecules", m))
    return True


class _Molecule:
    """Stand-in for openfermion's MolecularData (the .hdf5 files need h5py): only what _run and the optimizer read."""

    def __init__(self, name, n_qubits, n_alpha, n_beta):
        self.name, self.n_qubits, self.n_orbitals = name, n_qubits, n_qubits // 2
        self._na, self._nb = n_alpha, n_beta
        self.n_electrons = n_alpha + n_beta
        self.hf_energy = self.mp2_energy = self.ccsd_energy = self.fci_energy = float("nan")

    def get_n_alpha_electrons(self):
        return self._na

    def get_n_beta_electrons(self):
        return self._nb


def _prepare(backend, root):
    """Stubs + imports.  -> namespace of the reference modules the harness touches."""
    from oracle import ref_harness
    ref_harness.install_stubs()
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "mpl_toolkits", "mpl_toolkits.axes_grid1", "mpl_toolkits.axes_grid1.inset_locator"):
        ref_harness._stub(name)
    sys.modules["matplotlib"].gridspec = types.ModuleType("gridspec")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_toolkits.axes_grid1.inset_locator"].InsetPosition = object
    if root not in sys.path:
        sys.path.insert(0, root)
    os.environ["NAQS_ELOC_BACKEND"] = backend
    if backend == "b200":
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import naqs_b200
        assert naqs_b200.install(root) is True
    else:
        src_utils = importlib.import_module("src.utils")
        kp = os.path.join(HERE, "_ref", "src", "utils")
        if kp not in src_utils.__path__:
            src_utils.__path__.append(kp)
    plotting = importlib.import_module("src.utils.plotting")
    system = importlib.import_module("src.utils.system")

    def set_global_seed(seed=-1):  # system.py:64-79 minus sp.random.seed (removed from scipy); same draw order otherwise
        import torch
        random.seed(seed)
        np.random.seed(random.randint(0, 2 ** 32 - 1))
        random.randint(0, 2 ** 32)  # the draw the reference spends on scipy
        torch.manual_seed(random.randint(0, 2 ** 32))
        for _ in range(2):
            random.randint(0, 2 ** 32)

    def load_molecule(fname, hamiltonian_fname=None, verbose=True):
        name = os.path.split(fname.rstrip("/"))[-1]
        n, na, nb = ref_harness.MOLECULES[name]
        op = ref_harness.QubitOperator()
        op.terms = ref_harness.load_terms(name, reference_root=root)
        return _Molecule(name, n, na, nb), op

    system.set_global_seed, system.load_molecule = set_global_seed, load_molecule
    fig = types.SimpleNamespace(savefig=lambda *a, **k: None)
    plotting.plot_training = lambda *a, **k: fig
    base = importlib.import_module("experiments._base")
    base.set_global_seed, base.load_molecule, base.plot_training = set_global_seed, load_molecule, plotting.plot_training
    energy = importlib.import_module("src.optimizer.energy")
    if backend == "reference":
        import torch
        ham = importlib.import_module("src.optimizer.hamiltonian")
        for cls in (ham._PauliHamiltonianDynamic,) if hasattr(ham, "_PauliHamiltonianDynamic") else ():
            for meth in ("get_H", "update_H", "get_coupled_state_idxs"):
                if hasattr(cls, meth):
                    def wrap(f):
                        def g(self, idxs=None, *a, **k):
                            if torch.is_tensor(idxs):
                                idxs = idxs.detach().cpu().numpy()
                            return f(self, idxs, *a, **k) if idxs is not None else f(self, *a, **k)
                        return g
                    setattr(cls, meth, wrap(getattr(cls, meth)))
        get_inner = ham.PauliHamiltonian.get

        def get(hilbert, qubit_hamiltonian, hamiltonian_fname=None, restricted_idxs=None, *a, **k):  # scipy no longer indexes with tensors
            if torch.is_tensor(restricted_idxs):
                restricted_idxs = restricted_idxs.detach().cpu().numpy()
            return get_inner(hilbert, qubit_hamiltonian, hamiltonian_fname, restricted_idxs, *a, **k)
        ham.PauliHamiltonian.get = staticmethod(get)
    return types.SimpleNamespace(base=base, energy=energy)


def run_vmc(molecule="LiH", iters=20, seed=111, backend="b200", n_samps=int(1e5), n_hid=64, n_hid_phase=512, n_layer_phase=2,
            n_unq_samps_min=int(1e4), n_unq_samps_max=int(1e5), final_solve=True, quiet=True, model_device=None, record_psi=False):
    """experiments/_base._run with the flags of experiments/bash/naqs/batch_train.sh:14 for `iters` iterations.
    -> dict(idx=[per-iteration int64 arrays], eloc=[per-iteration float32 [M, 2]], log_eloc, log_var, times={...}, eig).
    model_device: None = the reference's default (cuda when available, wavefunction.py:33-34); "cpu" makes the whole
    trajectory bitwise reproducible, so two backends can be compared iteration by iteration."""
    import torch
    root = tree_root()
    if root is None:
        raise RuntimeError("no reference tree (neither /root/reference nor oracle/_ref/tree)")
    mods = _prepare(backend, root)
    rec = {"idx": [], "eloc": [], "psi": [], "t_eloc": [], "t_sample": [], "t_state2idx": [], "t_step": [], "eig": None}
    OB = mods.energy.OptimizerBase
    inner = OB.calculate_local_energy

    def sync():
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def recording_eloc(self, states_idx, psi=None, *a, **k):
        sync()
        t = time.perf_counter()
        out = inner(self, states_idx, psi, *a, **k)
        sync()
        rec["t_eloc"].append(time.perf_counter() - t)
        rec["idx"].append(np.asarray(states_idx.detach().cpu().numpy() if torch.is_tensor(states_idx) else states_idx).astype(np.int64).reshape(-1))
        rec["eloc"].append(out.detach().cpu().numpy().copy() if torch.is_tensor(out) else np.asarray(out))
        if record_psi and psi is not None:
            rec["psi"].append(psi.detach().cpu().numpy().copy() if torch.is_tensor(psi) else np.asarray(psi))
        return out

    OB.calculate_local_energy = recording_eloc
    step_inner = OB._SGD_step

    def timed_step(self, *a, **k):
        sync()
        t = time.perf_counter()
        out = step_inner(self, *a, **k)
        sync()
        rec["t_step"].append(time.perf_counter() - t)
        return out

    OB._SGD_step = timed_step
    wf_mod = importlib.import_module("src.naqs.wavefunction")
    if model_device is not None:
        base_init = wf_mod._NAQSComplex_Base.__init__

        def init_on(self, *a, device=None, **k):
            return base_init(self, *a, device=model_device, **k)
        wf_mod._NAQSComplex_Base.__init__ = init_on
    hil_mod = importlib.import_module("src.utils.hilbert")

    def timed(cls, name, key):
        f = getattr(cls, name)

        def g(self, *a, **k):
            sync()
            t = time.perf_counter()
            out = f(self, *a, **k)
            sync()
            rec[key].append(time.perf_counter() - t)
            return out
        setattr(cls, name, g)

    timed(wf_mod.NAQSComplex_NADE_orbitals, "sample", "t_sample")
    for name in dir(hil_mod):
        cls = getattr(hil_mod, name)
        if isinstance(cls, type) and "state2idx" in cls.__dict__:
            timed(cls, "state2idx", "t_state2idx")
    captured = {}
    solve_inner = mods.energy.PartialSamplingOptimizer.solve_H

    def solve(self, *a, **k):
        captured["opt"] = self
        if not final_solve:
            return float("nan"), None, 0
        out = solve_inner(self, *a, **k)
        rec["eig"] = float(np.real(out[0]))
        return out

    mods.energy.PartialSamplingOptimizer.solve_H = solve
    out_dir = tempfile.mkdtemp(prefix="naqs_vmc_")
    cwd = os.getcwd()
    devnull = open(os.devnull, "w")
    stdout = sys.stdout
    try:
        os.chdir(out_dir)
        if quiet:
            sys.stdout = devnull
        mods.base._run(molecule_fname=os.path.join(root, "molecules", molecule), exp_name=os.path.join(out_dir, "exp"), num_experiments=1,
                       n_samps=n_samps, n_samps_max=1e12, n_unq_samps_min=n_unq_samps_min, n_unq_samps_max=n_unq_samps_max,
                       n_train=iters, n_pretrain=0, output_freq=25, save_freq=None, n_hid=n_hid, n_layer=1, n_hid_phase=n_hid_phase,
                       n_layer_phase=n_layer_phase, comb_amp_phase=False, use_amp_spin_sym=True, use_phase_spin_sym=False,
                       aggregate_phase=False, overwrite_pauli_hamiltonian=False, loadH=False, presolveH=False, use_restrictedH=True, seed=seed)
    finally:
        sys.stdout = stdout
        devnull.close()
        os.chdir(cwd)
        shutil.rmtree(out_dir, ignore_errors=True)
    opt = captured.get("opt")
    LogKey = importlib.import_module("src.optimizer.utils").LogKey
    rec["log_eloc"] = [e for _, e in opt.log[LogKey.E_LOC]] if opt is not None else []
    rec["log_var"] = [e for _, e in opt.log[LogKey.E_LOC_VAR]] if opt is not None else []
    rec["backend"] = backend
    return rec


def save_record(rec, path):
    arrays = {f"idx_{i}": a for i, a in enumerate(rec["idx"])}
    arrays.update({f"eloc_{i}": a for i, a in enumerate(rec["eloc"])})
    arrays.update({f"psi_{i}": a for i, a in enumerate(rec.get("psi", []))})
    meta = {k: rec[k] for k in ("t_eloc", "t_sample", "t_state2idx", "t_step", "eig", "log_eloc", "log_var", "backend")}
    np.savez(path, meta=np.array(json.dumps(meta)), n=np.array(len(rec["idx"])), **arrays)


def load_record(path):
    z = np.load(path)
    rec = json.loads(str(z["meta"]))
    n = int(z["n"])
    rec["idx"] = [z[f"idx_{i}"] for i in range(n)]
    rec["eloc"] = [z[f"eloc_{i}"] for i in range(n)]
    rec["psi"] = [z[f"psi_{i}"] for i in range(n) if f"psi_{i}" in z.files]
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--molecule", default="LiH")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--seed", type=int, default=111)
    ap.add_argument("--backend", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-samps", type=int, default=int(1e5))
    ap.add_argument("--n-unq-min", type=int, default=int(1e4))
    ap.add_argument("--n-unq-max", type=int, default=int(1e5))
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--model-device", default=None)
    ap.add_argument("--record-psi", action="store_true")
    ap.add_argument("--device-resident", action="store_true", help="b200 backend: install(device_resident=True)")
    ap.add_argument("--fused-loss", action="store_true", help="b200 backend: install(fused_loss=True)")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    if a.device_resident:
        os.environ["NAQS_ELOC_DEVICE_RESIDENT"] = "1"
    if a.fused_loss:
        os.environ["NAQS_ELOC_FUSED_LOSS"] = "1"
    rec = run_vmc(a.molecule, a.iters, a.seed, a.backend, n_samps=a.n_samps, n_unq_samps_min=a.n_unq_min, n_unq_samps_max=a.n_unq_max,
                  final_solve=not a.no_solve, quiet=not a.verbose, model_device=a.model_device, record_psi=a.record_psi)
    save_record(rec, a.out)
    print(json.dumps({"backend": a.backend, "iters": len(rec["log_eloc"]), "eloc_calls": len(rec["eloc"]), "last_E": rec["log_eloc"][-1] if rec["log_eloc"] else None,
                      "eig": rec["eig"], "ms_step": 1e3 * float(np.mean(rec["t_step"])) if rec["t_step"] else None,
                      "ms_eloc": 1e3 * float(np.mean(rec["t_eloc"])) if rec["t_eloc"] else None}))


if __name__ == "__main__":
    main()
