"""The reference's own VMC loop (experiments/_base._run, the path of experiments/run.py) on the replacement.

`oracle/ref_vmc.py` runs the UNMODIFIED loop from the reference tree (/root/reference here, the staged git-ignored
copy oracle/_ref/tree on the GPU box) once per backend in a subprocess — NAQS_ELOC_BACKEND=reference (compiled Cython
kernels) and b200 (naqs_b200.install()) — with the same seed and the model on the CPU, so the trajectory is bitwise
reproducible and the E_loc tensors handed to the loss can be compared call by call."""
import os
import subprocess
import sys

import numpy as np
import pytest
from conftest import ROOT

from oracle import ref_vmc

pytestmark = pytest.mark.gpu


def _run(backend, out, molecule="LiH", iters=20, extra=(), quirks=False, model_device="cpu"):
    env = dict(os.environ, NAQS_ELOC_BACKEND=backend, NAQS_ELOC_REFERENCE_QUIRKS="1" if quirks else "0")
    cmd = [sys.executable, "-m", "oracle.ref_vmc", "--molecule", molecule, "--iters", str(iters), "--seed", "111", "--backend", backend,
           "--out", out, "--record-psi", *extra]
    if model_device:
        cmd += ["--model-device", model_device]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return ref_vmc.load_record(out)


@pytest.mark.skipif(ref_vmc.tree_root() is None, reason="no reference tree (oracle/_ref/tree is staged by __graft_entry__.build())")
def test_lih_vmc_loop_reference_vs_b200(tmp_path):
    """Stock flags of experiments/bash/naqs/batch_train.sh:14 (1e5 samples): every iteration samples ALL 225 sector states, in
    ascending key order — the batch on which the reference's get_H returns H in restricted order (quirk q1).
    With NAQS_ELOC_REFERENCE_QUIRKS=1 the replacement reproduces that pairing and the two trajectories are bitwise equal."""
    ref = _run("reference", str(tmp_path / "ref.npz"))
    b200 = _run("b200", str(tmp_path / "b200.npz"), quirks=True)
    assert len(ref["eloc"]) == len(b200["eloc"]) == 20
    assert all(len(i) == 225 for i in ref["idx"])
    for i, (ir, ib, er, eb) in enumerate(zip(ref["idx"], b200["idx"], ref["eloc"], b200["eloc"])):
        assert np.array_equal(ir, ib), f"iteration {i}: the sampled batches differ"
        assert er.dtype == eb.dtype == np.float32 and er.shape == eb.shape
        # the reference hands float32 pairs to the loss (complex.py:139-140): equal after that rounding
        assert np.array_equal(er, eb), f"iteration {i}: max |dE_loc| = {np.abs(er - eb).max():.3e}"
    assert np.array_equal(np.array(ref["log_eloc"]), np.array(b200["log_eloc"]))
    assert np.array_equal(np.array(ref["log_var"]), np.array(b200["log_var"]))
    # solve_H at the end of _run (energy.py:762-786) goes through get_H(idxs) without update_H: the replacement computes
    # the rows on demand — LiH FCI energy of the sampled sub-space
    assert abs(ref["eig"] - b200["eig"]) < 1e-9 and abs(b200["eig"] - (-7.78446028)) < 1e-6


@pytest.mark.skipif(ref_vmc.tree_root() is None, reason="no reference tree")
def test_vmc_eloc_calls_match_reference_kernels_on_the_same_inputs(tmp_path):
    """Independent of the trajectory: every (states, psi) batch the b200 run saw, recomputed with the reference's compiled
    kernels and numpy orchestration (oracle/ref_path.py, pinned to the live reference in the CPU suite)."""
    from conftest import load_table
    from oracle import ref_path
    b200 = _run("b200", str(tmp_path / "b200.npz"), iters=6, extra=("--no-solve",))
    xy, yz, c, N, na, nb = load_table("LiH")
    ph = ref_path.ReferencePath(xy, yz, c, N, na, nb)
    assert len(b200["psi"]) == len(b200["eloc"]) == 6
    for idx, psi, eloc in zip(b200["idx"], b200["psi"], b200["eloc"]):
        psi_c = (psi[:, 0] + 1j * psi[:, 1]).astype(np.complex64)
        ref = ph.local_energy(idx, psi_c)
        ref32 = np.stack([ref.real, ref.imag], -1).astype(np.float32)
        assert np.array_equal(ref32, eloc)


@pytest.mark.skipif(ref_vmc.tree_root() is None, reason="no reference tree")
def test_lih_vmc_loop_partial_samples_default_mode(tmp_path):
    """400 samples per iteration: the batches are proper subsets of the sector (no quirk involved), default mode of the
    replacement; same batches and E_loc equal to float32 rounding for 20 iterations."""
    extra = ("--n-samps", "400", "--n-unq-min", "400", "--n-unq-max", "100000", "--no-solve")
    ref = _run("reference", str(tmp_path / "ref.npz"), extra=extra)
    b200 = _run("b200", str(tmp_path / "b200.npz"), extra=extra)
    assert len(ref["eloc"]) == len(b200["eloc"]) == 20
    assert all(0 < len(i) < 225 for i in ref["idx"])
    for i, (ir, ib, er, eb) in enumerate(zip(ref["idx"], b200["idx"], ref["eloc"], b200["eloc"])):
        assert np.array_equal(ir, ib), f"iteration {i}: the sampled batches differ"
        # complex128 E_loc agrees to ~1e-16 (the fused kernel adds the couplings of a row in XY-mask order, scipy's CSR product
        # in column order); after the float32 truncation of complex.py:139-140 a component that is itself ~1e-17 may differ
        # in its last bits, so: float32-ulp tolerance relative to |E_loc|, not bitwise
        mag = np.abs(er[:, 0] + 1j * er[:, 1])[:, None]
        assert np.all(np.abs(er - eb) <= 1.2e-7 * mag + 1e-30), f"iteration {i}: max |dE_loc| = {np.abs(er - eb).max():.3e}"
    assert np.allclose(np.array(ref["log_eloc"]), np.array(b200["log_eloc"]), rtol=1e-6, atol=0)


@pytest.mark.skipif(ref_vmc.tree_root() is None, reason="no reference tree")
def test_fused_loss_step_matches_reference_step(tmp_path):
    """install(fused_loss=True): _SGD_step with the loss statistics from one fp64 device kernel (naqs_loss_terms) and a
    detached weight vector for autograd.  Same model (CPU, seeded), same batches: the logged energies / variances of the
    reference's float32 tensor arithmetic (energy.py:316-329, 367-375) are reproduced to float32 rounding, and so is E_loc
    in every iteration — i.e. the parameter updates driven by the fused loss follow the reference's."""
    extra = ("--n-samps", "400", "--n-unq-min", "400", "--n-unq-max", "100000", "--no-solve")
    ref = _run("reference", str(tmp_path / "ref.npz"), iters=10, extra=extra)
    b200 = _run("b200", str(tmp_path / "b200.npz"), iters=10, extra=extra + ("--fused-loss",))
    assert len(ref["log_eloc"]) == len(b200["log_eloc"]) == 10
    assert np.allclose(ref["log_eloc"], b200["log_eloc"], rtol=2e-5, atol=0)
    assert np.allclose(ref["log_var"], b200["log_var"], rtol=1e-3, atol=1e-6)
    for i, (ir, ib) in enumerate(zip(ref["idx"], b200["idx"])):
        assert np.array_equal(ir, ib), f"iteration {i}: the sampled batches differ"


@pytest.mark.skipif(ref_vmc.tree_root() is None, reason="no reference tree")
def test_device_resident_loop_first_iteration_matches_host_loop(tmp_path):
    """install(device_resident=True): sampler output stays on the GPU, state2idx runs on the device, E_loc is computed from
    CUDA tensors and handed back as one.  The model sits on the GPU in both runs (its kernels are not bitwise
    reproducible across runs), so the comparison is the first iteration — same seed, same initial weights — plus the
    logged energy trajectory to a loose tolerance."""
    extra = ("--n-samps", "400", "--n-unq-min", "400", "--n-unq-max", "100000", "--no-solve")
    host = _run("b200", str(tmp_path / "host.npz"), iters=5, extra=extra, model_device=None)
    dev = _run("b200", str(tmp_path / "dev.npz"), iters=5, extra=extra + ("--device-resident", "--fused-loss"), model_device=None)
    assert np.array_equal(np.sort(host["idx"][0]), np.sort(dev["idx"][0]))
    oh, od = np.argsort(host["idx"][0]), np.argsort(dev["idx"][0])
    # host loop: float32 pairs; device-resident + fused loss: the fp64 E_loc itself is recorded
    eh, ed = host["eloc"][0][oh], dev["eloc"][0][od]
    assert np.allclose(eh, ed, rtol=2e-5, atol=1e-6)
    assert np.allclose(host["log_eloc"][0], dev["log_eloc"][0], rtol=1e-5)
