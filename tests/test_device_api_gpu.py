"""Device-resident hand-off and fused loss statistics (§8f-2) through the reference-shaped entry points."""
import types

import numpy as np
import pytest
import torch
from conftest import load_case, load_table, load_terms_json

import naqs_b200
from naqs_b200 import _lib

pytestmark = pytest.mark.gpu


def _optimizer_stub(mol="LiH"):
    xy, yz, c, N, na, nb = load_table(mol)
    hil = naqs_b200.Hilbert.get(N, na, nb, encoding=naqs_b200.Encoding.SIGNED)
    op = types.SimpleNamespace(terms=load_terms_json(mol))
    sec = hil.get_subspace(ret_states=False, ret_idxs=True)
    ph = naqs_b200.PauliHamiltonian.get(hil, op, restricted_idxs=sec, dtype=np.float64)
    return types.SimpleNamespace(pauli_hamiltonian=ph, hilbert=hil), hil


def test_calculate_local_energy_takes_cuda_tensors_and_int8_rows():
    """CUDA state rows -> naqs_state2idx -> CUDA indices -> calculate_local_energy -> CUDA float32 [M, 2], equal to the host
    path (and to the reference fixture) without the batch ever visiting the host."""
    opt, hil = _optimizer_stub("LiH")
    case = load_case("LiH_small")
    idx = case["states"].astype(np.int64)
    psi = case["psi"].astype(np.complex64)
    host = naqs_b200.calculate_local_energy(opt, torch.from_numpy(idx.astype(np.int16)), torch.view_as_real(torch.from_numpy(psi)))
    assert host.dtype == torch.float32 and not host.is_cuda
    # int8 +-1 rows on the device, packed by the device kernel (hilbert.py:573-581)
    rows = hil.idx2state(idx).to("cuda")
    keys = torch.empty((len(idx), 1), dtype=torch.int64, device="cuda")
    _lib.check(_lib.load().naqs_state2idx(_lib.ptr(rows.contiguous()), len(idx), hil.N, 1, _lib.ptr(keys), _lib.stream_ptr(rows.device)))
    assert np.array_equal(keys.cpu().numpy().reshape(-1), idx)
    psi_dev = torch.view_as_real(torch.from_numpy(psi)).to("cuda")
    dev = naqs_b200.calculate_local_energy(opt, keys.reshape(-1).to(torch.int16), psi_dev)
    assert dev.is_cuda and dev.dtype == torch.float32 and dev.shape == (len(idx), 2)
    assert torch.equal(dev.cpu(), host)
    ref = case["eloc"]
    assert np.abs(dev.cpu().numpy()[:, 0] + 1j * dev.cpu().numpy()[:, 1] - ref).max() <= 2e-7 * np.abs(ref).max()
    dev64 = naqs_b200.calculate_local_energy(opt, keys.reshape(-1), psi_dev, ret_device=True)
    assert dev64.is_cuda and dev64.dtype == torch.float64
    assert np.abs(_lib.complex_from_pairs(dev64) - ref).max() <= 1e-12 * np.abs(ref).max()


def test_loss_terms_match_the_reference_tensor_arithmetic():
    """naqs_loss_terms against the float32 torch expressions of energy.py:316-329, 367-375, plus the gradient identity
    d exp_op / d log_psi == grad_weight."""
    opt, hil = _optimizer_stub("LiH")
    case = load_case("LiH_small")
    rng = np.random.default_rng(5)
    M = len(case["states"])
    eloc = torch.from_numpy(np.stack([case["eloc"].real, case["eloc"].imag], -1)).to("cuda")
    w = torch.from_numpy(rng.random(M)).float()
    w /= w.sum()
    terms = naqs_b200.energy.loss_terms(opt.pauli_hamiltonian.table, eloc, w)
    e32 = eloc.float().cpu()
    sw = w.unsqueeze(-1)
    e_corr = e32 - (sw * e32).sum(axis=0)                                           # energy.py:328
    log_psi = torch.from_numpy(rng.normal(size=(M, 2))).float().requires_grad_(True)
    re = log_psi[..., 0] * e_corr[..., 0] - log_psi[..., 1] * e_corr[..., 1]        # cplx.real(cplx.scalar_mult(log_psi, e_loc_corr))
    exp_op = 2 * (w * re).sum()                                                     # energy.py:329
    exp_op.backward()
    assert torch.allclose(terms["eloc"].cpu(), e32)
    assert torch.allclose(terms["eloc_corr"].cpu(), e_corr, rtol=1e-5, atol=1e-6)
    assert torch.allclose(terms["grad_weight"].cpu(), log_psi.grad, rtol=1e-4, atol=1e-7)
    wn = w / w.sum()
    energy = (wn * e32[:, 0]).sum()                                                 # energy.py:372-375
    var = ((e32[:, 0] - energy).pow(2) * wn).sum()
    ev = terms["energy_var"].cpu().numpy()
    assert abs(ev[0] - energy.item()) <= 2e-6 * abs(energy.item()) and abs(ev[2] - var.item()) <= 1e-4 * abs(var.item()) + 1e-7
    assert np.allclose(terms["sums"].cpu().numpy()[[0, 4]], [float(w.double().sum()), M])


def test_out_of_range_key_is_flagged_not_dereferenced():
    """a key >= 2^N must never become an address.  Host entry: IndexError (what the reference raises when it
    indexes its 2^N lookup table, hilbert.py:607-640); device entry: row NaN + table.check() raises."""
    xy, yz, c, N, na, nb = load_table("LiH")
    t = naqs_b200.DeviceTermTable(xy, yz, c, N, na, nb)
    case = load_case("LiH_small")
    st = case["states"].astype(np.int64).copy()
    psi = case["psi"].astype(np.complex128)
    good = _lib.complex_from_pairs(t.local_energy(st, psi))
    st_bad = st.copy()
    st_bad[3] = (1 << N) + 5
    with pytest.raises(IndexError):
        t.local_energy_host(st_bad, psi)
    out = _lib.complex_from_pairs(t.local_energy(st_bad, psi))
    with pytest.raises(IndexError):
        t.check()
    t.check()  # the flag is cleared by the report
    assert np.isnan(out[3].real)
    keep = np.ones(len(st), bool)
    keep[3] = False
    ref = _lib.complex_from_pairs(t.local_energy(st[keep], psi[keep]))  # the bad key is simply absent from the table
    assert np.allclose(out[keep], ref, rtol=1e-13, atol=0) and np.all(np.isfinite(good))


def test_small_host_batches_replay_a_cuda_graph():
    """naqs_eloc_host with a small batch: first call plain, second call captured into a CUDA graph, later calls replay it
    (table.cu).  Every call must see ITS inputs (staged through page-locked memory), whatever ran on the table in between."""
    import naqs_b200
    from oracle import c_oracle
    from oracle import eloc_oracle as eo
    xy, yz, c, N, na, nb = load_table("LiH")
    t, ct = naqs_b200.DeviceTermTable(xy, yz, c, N, na, nb), c_oracle.COracleTable(xy, yz, c, N, na, nb)
    sec = eo.sector_keys(N, na, nb)[:, 0].astype(np.uint64)
    launches = []
    for it in range(6):
        st = np.random.default_rng(100 + it).permutation(sec)[:200 if it != 3 else 150]     # call 3 has another signature
        psi = eo.synthetic_psi(len(st), seed=200 + it).astype(np.complex64)
        before = naqs_b200.launch_count()
        got = t.local_energy_host(st.astype(np.int16), psi, assume_unique=True, out_dtype=np.complex64)
        launches.append(naqs_b200.launch_count() - before)
        ref = ct.local_energy(st, psi.astype(np.complex128))
        assert np.abs(got - ref.astype(np.complex64)).max() <= 2e-6 * np.abs(ref).max(), it
        if it == 4:   # a device-API call with another batch rebuilds the lookup in between: the replay must not depend on it
            other = sec[:77]
            e = t.local_energy(other, eo.synthetic_psi(77, seed=1))
            assert np.abs(naqs_b200._lib.complex_from_pairs(e) - ct.local_energy(other, eo.synthetic_psi(77, seed=1))).max() < 1e-10
    assert launches[0] == launches[1] == launches[2] == launches[4] == launches[5] > 0   # replays are counted like plain runs
    # complex128 in/out and an out-of-range key through the replayed graph
    st = sec[:64].copy()
    for it in range(4):
        psi = eo.synthetic_psi(64, seed=300 + it)
        got = t.local_energy_host(st, psi, assume_unique=True)
        assert np.abs(got - ct.local_energy(st, psi)).max() <= 1e-12 * np.abs(got).max()
    bad = st.copy()
    bad[5] = np.uint64(1) << np.uint64(40)
    with pytest.raises(IndexError):
        t.local_energy_host(bad, eo.synthetic_psi(64, seed=9), assume_unique=True)
    got = t.local_energy_host(st, eo.synthetic_psi(64, seed=303), assume_unique=True)   # and the table is usable afterwards
    assert np.abs(got - ct.local_energy(st, eo.synthetic_psi(64, seed=303))).max() <= 1e-12 * np.abs(got).max()
