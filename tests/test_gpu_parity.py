"""Parity tests proper (run on the B200 with -m gpu): the CUDA path, called through the C ABI
(ctypes -> libnaqs_eloc.so), against the golden fixtures and the oracle on seeded inputs.

Bars (BASELINE.json north_star): coupled-state sets, matrix elements and signs BIT-EXACT;
fp64 E_loc within 1e-12 relative."""
import json
import os

import numpy as np
import pytest
import torch
from conftest import CASE_TABLE, F32_CASE_TABLE, GOLDEN, case_sector, load_case, load_table, random_sector_states

pytestmark = pytest.mark.gpu

ELOC_RTOL = 1e-12  # north_star: "fp64 E_loc within 1e-12 relative"


def _mods():
    import naqs_b200
    from oracle import c_oracle
    from oracle import eloc_oracle as eo
    return naqs_b200, c_oracle, eo


def random_keys(N, m, seed):
    """m distinct uniformly random keys below 2^N (N <= 62)."""
    rng = np.random.default_rng(seed)
    if 2 ** N <= 4 * m:
        return rng.choice(2 ** N, m, replace=False).astype(np.uint64)
    out = np.unique(rng.integers(0, 2 ** N, size=2 * m, dtype=np.int64))
    return rng.permutation(out)[:m].astype(np.uint64)


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def gpu_eloc(table, states, psi, **kw):
    import naqs_b200
    out = table.local_energy(states, psi, **kw)
    return naqs_b200._lib.complex_from_pairs(out)


def make_tables(mol, sector=True):
    nb200, c_oracle, _ = _mods()
    xy, yz, c, N, na, nb = load_table(mol)
    if not sector:
        na = nb = None
    return nb200.DeviceTermTable(xy, yz, c, N, na, nb), c_oracle.COracleTable(xy, yz, c, N, na, nb), (N, na, nb)


# ------------------------------------------------------------------------------------------- golden
@pytest.mark.parametrize("name", sorted(CASE_TABLE))
def test_eloc_matches_reference_fixture(name):
    nb200, _, _ = _mods()
    case = load_case(name)
    xy, yz, c, N, _, _ = load_table(CASE_TABLE[name])
    na, nb = case_sector(case, name)
    t = nb200.DeviceTermTable(xy, yz, c, N, na, nb)
    before = nb200.launch_count()
    e = gpu_eloc(t, case["states"], case["psi"])
    assert nb200.launch_count() > before, "no kernel of libnaqs_eloc.so was launched"
    assert rel_err(e, case["eloc"]).max() <= ELOC_RTOL
    # hash lookup and dense (direct-address) lookup agree; they may run with different launch shapes (row-order vs
    # key-order walk, table chunks), which only changes the order in which the per-group products are added up
    e_hash = gpu_eloc(t, case["states"], case["psi"], kind=nb200._lib.LOOKUP_HASH)
    assert rel_err(e_hash, case["eloc"]).max() <= ELOC_RTOL
    if N <= 26:  # a direct-address table of 2^30 entries (Li2O) would be 16 GB
        e_dense = gpu_eloc(t, case["states"], case["psi"], kind=nb200._lib.LOOKUP_DENSE)
        assert rel_err(e_hash, e_dense).max() <= 1e-13 and rel_err(e_dense, case["eloc"]).max() <= ELOC_RTOL
    # same call, same inputs -> bit-identical output (deterministic accumulation order)
    assert np.array_equal(gpu_eloc(t, case["states"], case["psi"]), e)
    # complex128 psi carrying the same values gives identical results (exact promotion, sparse_math.pyx:33-37)
    assert np.array_equal(gpu_eloc(t, case["states"], case["psi"].astype(np.complex128)), e)
    # host-buffer entry (naqs_eloc_host) == device-buffer entry
    assert np.array_equal(t.local_energy_host(case["states"], case["psi"]), e)
    # assume_unique (the reference's own call-site contract, energy.py:245) may switch to the complex64 dense table
    e_u = gpu_eloc(t, case["states"], case["psi"], assume_unique=True)
    assert rel_err(e_u, case["eloc"]).max() <= ELOC_RTOL and rel_err(e_u, e).max() <= 1e-13
    assert np.array_equal(t.local_energy_host(case["states"], case["psi"], assume_unique=True), e_u)
    # the reference's index dtype (int16 / int32, hilbert.py:405-410) and float32-pair output (complex.py:139-140)
    idt = np.int16 if N < 16 else np.int32
    e32 = t.local_energy_host(case["states"].astype(np.int64).astype(idt), case["psi"], out_dtype=np.complex64)
    assert e32.dtype == np.complex64 and np.array_equal(e32, e.astype(np.complex64))


@pytest.mark.parametrize("name", ["LiH_sector", "LiH_small", "H2O_sector", "LiH_full_600"])
def test_rows_match_reference_csr_bit_exact(name):
    nb200, _, _ = _mods()
    case = load_case(name)
    xy, yz, c, N, _, _ = load_table(CASE_TABLE[name])
    na, nb = case_sector(case, name)
    t = nb200.DeviceTermTable(xy, yz, c, N, na, nb)
    indptr, cols, ridx, vals = (x.cpu().numpy() for x in t.rows(case["states"]))
    assert np.array_equal(indptr, case["rows_indptr"])
    for m in range(len(indptr) - 1):
        lo, hi = indptr[m], indptr[m + 1]
        o = np.argsort(ridx[lo:hi], kind="stable")
        assert np.array_equal(ridx[lo:hi][o], case["rows_cols_restricted"][lo:hi])
        assert np.array_equal(cols[lo:hi, 0][o].view(np.uint64), case["rows_cols_keys"][lo:hi])
        assert np.array_equal(vals[lo:hi][o], case["rows_vals"][lo:hi])
    # coupled-state set (radix sort + unique on the device) == np.unique of the reference's columns
    uniq = t.coupled_state_set(case["states"])[:, 0].cpu().numpy().view(np.uint64)
    assert np.array_equal(uniq, np.unique(case["rows_cols_keys"]))
    assert np.array_equal(np.sort(t.restricted_index(uniq.view(np.int64)).cpu().numpy()), case["coupled_unique_restricted"])


@pytest.mark.parametrize("name", sorted(F32_CASE_TABLE))
def test_float32_hamiltonian_matches_reference_fixture(name):
    """PauliHamiltonian.get(dtype=np.float32) — the reference constructor's default (hamiltonian.py:48): matrix elements
    accumulate in float32 (__inner_int64_float) and must be bit-identical; the reference then forms E_loc in complex64
    (sparse_math.pyx:13-41), the device in complex128, so E_loc agrees to float32 rounding."""
    nb200, _, eo = _mods()
    case = load_case(name)
    xy, yz, c, N, _, _ = load_table(F32_CASE_TABLE[name])
    na, nb = case_sector(case, name)
    with pytest.raises(ValueError):  # float64 coefficients are not float32 values
        nb200.DeviceTermTable(xy, yz, c, N, na, nb).set_precision(np.float32)
    with pytest.raises(TypeError):   # np.float128 has no device type (hamiltonian_math.pyx long double kernels)
        nb200.DeviceTermTable(xy, yz, c, N, na, nb).set_precision(np.longdouble)
    c32 = c.astype(np.float32)
    t = nb200.DeviceTermTable(xy, yz, c32, N, na, nb).set_precision(np.float32)
    with pytest.raises(nb200.NaqsError):
        t.set_algo("sliced")
    if "rows_vals" in case:
        indptr, cols, ridx, vals = (x.cpu().numpy() for x in t.rows(case["states"]))
        assert np.array_equal(indptr, case["rows_indptr"])
        for m in range(len(indptr) - 1):
            lo, hi = indptr[m], indptr[m + 1]
            o = np.argsort(ridx[lo:hi], kind="stable")
            assert np.array_equal(cols[lo:hi, 0][o].view(np.uint64), case["rows_cols_keys"][lo:hi])
            assert np.array_equal(vals[lo:hi][o], case["rows_vals"][lo:hi])  # bit-exact float32 matrix elements
    # dense [M, Kxy] elements against the float32 oracle (itself pinned on the fixture by the CPU suite)
    H_o, _ = eo.hamiltonian_dense_rows(eo.TermTable(xy, yz, c, N, na, nb, dtype=np.float32), case["states"][:200])
    assert np.array_equal(t.hij_dense(case["states"][:200]).cpu().numpy(), H_o.astype(np.float64))
    scale = np.abs(case["eloc"]).max()
    for kw in ({}, {"kind": nb200._lib.LOOKUP_HASH}, {"assume_unique": True}):
        e = gpu_eloc(t, case["states"], case["psi"], **kw)
        assert np.abs(e - case["eloc"]).max() <= 2e-5 * scale
    # switching back restores float64 accumulation of the same (float32-valued) coefficients
    t.set_precision(np.float64).set_algo("sliced")
    e64 = gpu_eloc(t, case["states"], case["psi"])
    o64 = eo.local_energy(eo.TermTable(xy, yz, c32.astype(np.float64), N, na, nb), case["states"][:100], case["psi"][:100],
                          table_keys=case["states"], table_psi=case["psi"])
    assert rel_err(e64[:100], o64).max() <= ELOC_RTOL


@pytest.mark.parametrize("mol,sector,m", [("LiH", True, 225), ("N2", True, 6000), ("N2", False, 50000), ("Li2O", True, 4000), ("H2S", True, 2000)])
def test_direct_and_sliced_formulations_agree(mol, sector, m):
    """The two kernel formulations (AND/POPC walk vs nibble-sliced parity + group LUT) produce the same E_loc; both are
    checked against the oracle, at sizes that exercise every launch shape (1024/512/256-thread CTAs, table chunks)."""
    nb200, c_oracle, eo = _mods()
    t, ct, (N, na, nb) = make_tables(mol, sector)
    if sector:
        st = random_sector_states(N, na, nb, min(m, eo.comb((N + 1) // 2, na) * eo.comb(N // 2, nb)), seed=31)
    else:
        st = random_keys(N, m, seed=32)
    psi = eo.synthetic_psi(len(st), seed=33)
    ref = ct.local_energy(st, psi)
    out = {}
    for algo in ("sliced", "direct"):
        t.set_algo(algo)
        for kind in (nb200._lib.LOOKUP_HASH, nb200._lib.LOOKUP_DENSE) if N <= 22 else (nb200._lib.LOOKUP_HASH,):
            e = gpu_eloc(t, st, psi, kind=kind)
            assert rel_err(e, ref).max() <= ELOC_RTOL, (algo, kind)
            out[(algo, kind)] = e
    # the sliced formulation adds the 6-term chunk sums of groups with more than 6 terms (sliced.cuh), the direct one walks
    # every term serially: matrix elements of those groups agree to a few ulp, E_loc well inside the 1e-12 bar
    a, b = out[("sliced", nb200._lib.LOOKUP_HASH)], out[("direct", nb200._lib.LOOKUP_HASH)]
    assert rel_err(a, b).max() <= ELOC_RTOL


# ------------------------------------------------------------------------------------------- oracle, seeded
@pytest.mark.parametrize("mol,m,sector", [("LiH", 225, True), ("H2O", 441, True), ("NH3", 3136, True), ("N2", 14400, True),
                                          ("N2_2.25", 5000, True), ("C2", 8000, True), ("H2S", 3000, True), ("Li2O", 3000, True),
                                          ("N2", 20000, False), ("H2O", 16384, False), ("Li2O", 1500, False)])
def test_eloc_vs_oracle(mol, m, sector):
    nb200, c_oracle, eo = _mods()
    t, ct, (N, na, nb) = make_tables(mol, sector)
    if sector:
        sec_size = eo.comb((N + 1) // 2, na) * eo.comb(N // 2, nb)
        st = eo.sector_keys(N, na, nb)[:, 0] if sec_size <= m else random_sector_states(N, na, nb, m, seed=7)
        st = st[np.random.default_rng(3).permutation(len(st))]
    else:
        st = random_keys(N, m, seed=5)
    psi = eo.synthetic_psi(len(st), seed=11)
    e = gpu_eloc(t, st, psi)
    ref = ct.local_energy(st, psi)
    assert rel_err(e, ref).max() <= ELOC_RTOL
    assert np.all(np.isfinite(e.view(np.float64)))


@pytest.mark.parametrize("mol,m,sector", [("LiH", 225, True), ("N2", 3000, True), ("N2", 3000, False), ("Li2O", 500, True), ("H2S", 700, True)])
def test_rows_and_hij_vs_oracle_bit_exact(mol, m, sector):
    nb200, c_oracle, eo = _mods()
    t, ct, (N, na, nb) = make_tables(mol, sector)
    if sector:
        st = random_sector_states(N, na, nb, min(m, eo.comb((N + 1) // 2, na) * eo.comb(N // 2, nb)), seed=21)
    else:
        st = random_keys(N, m, seed=22)
    indptr, cols, ridx, vals = (x.cpu().numpy() for x in t.rows(st))
    i2, c2, v2 = ct.rows(st)
    assert np.array_equal(indptr, i2)
    assert np.array_equal(cols.view(np.uint64), c2)          # coupled-state sets, in the same (XY) order
    assert np.array_equal(vals, v2)                          # matrix elements incl. signs, bit-exact
    assert np.array_equal(ridx, ct.restricted_index(c2))
    h = t.hij_dense(st[:200]).cpu().numpy()
    assert np.array_equal(h.reshape(-1), ct.hij_dense(st[:200]))
    uniq = t.coupled_state_set(st)[:, 0].cpu().numpy().view(np.uint64)
    assert np.array_equal(uniq, np.unique(c2[:, 0]))


def test_rows_are_independent_of_the_table_chunking():
    """The stored-row kernels cut the term table into 16 / 8 / 4 / 2 / 1 chunks on group boundaries depending on the batch size
    (rows.cu rows_chunks_for): the CSR rows, dense H_ij and restricted column indices of the same states must not depend on it,
    and equal the oracle's serial sums."""
    nb200, c_oracle, eo = _mods()
    t, ct, (N, na, nb) = make_tables("N2", sector=False)
    rng = np.random.default_rng(77)
    head = random_keys(N, 4000, seed=78)
    i2, c2, v2 = ct.rows(head)
    for m in (4000, 30000, 60000, 120000, 250000, 320000):   # 16, 16, 8, 4, 2, 1 chunks on a 148-SM device
        st = np.concatenate([head, rng.integers(0, 2 ** N, size=m - len(head), dtype=np.int64).astype(np.uint64)])
        indptr, cols, ridx, vals = t.rows(st)
        n_head = int(indptr[len(head)].item())
        assert np.array_equal(indptr[: len(head) + 1].cpu().numpy(), i2), m
        assert np.array_equal(cols[:n_head].cpu().numpy().view(np.uint64), c2), m
        assert np.array_equal(vals[:n_head].cpu().numpy(), v2), m
        assert np.array_equal(ridx[:n_head].cpu().numpy(), ct.restricted_index(c2)), m
        del indptr, cols, ridx, vals
    ts, cts, (N, na, nb) = make_tables("N2", sector=True)
    sec = random_sector_states(N, na, nb, 2500, seed=79)
    a = ts.hij_dense(sec).cpu().numpy()
    assert np.array_equal(a.reshape(-1), cts.hij_dense(sec))


@pytest.mark.parametrize("N,K,M", [(40, 1500, 3000), (63, 4000, 2000), (64, 2000, 1500), (100, 3000, 2500), (127, 1000, 1000), (20, 30000, 500)])
def test_wide_and_large_synthetic_tables(N, K, M):
    """64-/128-bit masks and tables larger than one shared-memory tile."""
    nb200, c_oracle, eo = _mods()
    xy, yz, c = eo.synthetic_table(N, K, seed=N + K)
    st = eo.synthetic_states(N, M, seed=N)
    psi = eo.synthetic_psi(M, seed=K)
    t, ct = nb200.DeviceTermTable(xy, yz, c, N), c_oracle.COracleTable(xy, yz, c, N)
    # the lookup table = the batch plus every coupled state of its first 50 members, so that many lookups hit
    _, cols, _ = ct.rows(st[:50])
    tk = np.unique(np.concatenate([st, cols]), axis=0)
    tp = eo.synthetic_psi(len(tk), seed=5)
    e = gpu_eloc(t, st, psi, table_keys=tk, table_psi=tp)
    ref = ct.local_energy(st, psi, tk, tp)
    assert rel_err(e, ref).max() <= ELOC_RTOL
    indptr, cols_g, ridx_none, vals = t.rows(st[:300], with_restricted_index=False)
    assert ridx_none is None
    indptr, cols_g, vals = indptr.cpu().numpy(), cols_g.cpu().numpy(), vals.cpu().numpy()
    i2, c2, v2 = ct.rows(st[:300])
    assert np.array_equal(indptr, i2) and np.array_equal(cols_g.view(np.uint64), c2) and np.array_equal(vals, v2)
    uniq = t.unique_keys(c2).cpu().numpy().view(np.uint64)
    exp = np.unique(c2, axis=0)
    exp = exp[np.lexsort(tuple(exp[:, w] for w in range(exp.shape[1])))]
    assert np.array_equal(uniq, exp)


@pytest.mark.parametrize("N,K,M,R", [(40, 600, 200_000, 2000), (70, 400, 150_000, 2000), (30, 900, 300_000, 2000), (30, 900, 160_000, 200),
                                     (50, 500, 170_000, 100)])
def test_large_batches_with_global_filter(N, K, M, R):
    """Batches above 2^17 keys: the Bloom filter no longer fits shared memory and is consulted in L2 before every bucket /
    slot probe (64-bit bucketed table, 128-bit slot table, and a 32-bit-key table).  The lookup table is the batch plus the
    coupled states of some rows, so hits and misses both occur; rows are checked against the oracle on a sub-sample."""
    nb200, c_oracle, eo = _mods()
    xy, yz, c = eo.synthetic_table(N, K, seed=N + K)
    st = eo.synthetic_states(N, M, seed=N + 1)
    psi = eo.synthetic_psi(M, seed=K + 1)
    t, ct = nb200.DeviceTermTable(xy, yz, c, N), c_oracle.COracleTable(xy, yz, c, N)
    _, cols, _ = ct.rows(st[:R])
    tk = np.unique(np.concatenate([st, cols]), axis=0)
    tp = eo.synthetic_psi(len(tk), seed=7)
    assert len(tk) > (1 << 17) and (R > 500 or len(tk) <= 5 << 16)
    e = gpu_eloc(t, st, psi, table_keys=tk, table_psi=tp, kind=nb200._lib.LOOKUP_HASH)
    sub = np.concatenate([np.arange(1000), np.random.default_rng(3).choice(M, 1000, replace=False)])
    ref = ct.local_energy(st[sub], psi[sub], tk, tp)
    assert rel_err(e[sub], ref).max() <= ELOC_RTOL
    # the filter only removes probes that would miss: the same hits are added up without it (possibly in another order,
    # because the per-thread queue then holds other entries between them)
    os.environ["NAQS_ELOC_NO_FILTER"] = "1"
    try:
        e2 = gpu_eloc(t, st, psi, table_keys=tk, table_psi=tp, kind=nb200._lib.LOOKUP_HASH)
    finally:
        del os.environ["NAQS_ELOC_NO_FILTER"]
    assert rel_err(e2, e).max() <= 1e-13


def test_caller_owned_dense_table_alignment_and_sector():
    """naqs_lookup_attach_dense32: the table must be aligned to its size (XOR addressing); out-of-sector keys in a caller-owned
    table are ignored by the kernel's own sector test (the library's own tables drop them at build time)."""
    nb200, c_oracle, eo = _mods()
    from naqs_b200 import distributed as nd
    t, ct, (N, na, nb) = make_tables("LiH", True)
    st = random_sector_states(N, na, nb, 200, seed=5)
    psi = eo.synthetic_psi(len(st), seed=6)
    stray = np.array([0, 1, 2 ** N - 1], dtype=np.uint64)  # not in the (2, 2) sector
    keys = torch.from_numpy(np.concatenate([st, stray]).view(np.int64)).cuda()
    amps = torch.from_numpy(np.concatenate([psi, eo.synthetic_psi(3, seed=9)])).cuda()
    tbl = nd.aligned_dense_table(t)
    assert tbl.data_ptr() % (8 << N) == 0
    nd.allreduce_dense_table(t, keys, amps, out=tbl)   # world size 1: scatter + attach
    e = nb200._lib.complex_from_pairs(t.local_energy(st, psi, rebuild_lookup=False))
    assert rel_err(e, ct.local_energy(st, psi)).max() <= ELOC_RTOL
    raw = torch.empty(((2 << N) + 2, 2), dtype=torch.int32, device="cuda")
    view = raw[:1 << N] if raw.data_ptr() % (8 << N) else raw[1:(1 << N) + 1]   # a [2^N, 2] view that is NOT aligned to 8 * 2^N bytes
    assert view.data_ptr() % (8 << N) != 0
    with pytest.raises(ValueError):
        t.attach_dense32(view)


# ------------------------------------------------------------------------------------------- edge cases
def test_edge_cases():
    nb200, c_oracle, eo = _mods()
    t, ct, (N, na, nb) = make_tables("LiH", True)
    sec = eo.sector_keys(N, na, nb)[:, 0]
    psi = eo.synthetic_psi(len(sec), 1)
    # empty batch
    assert gpu_eloc(t, sec[:0], psi[:0]).shape == (0,)
    ip, cols, ridx, vals = t.rows(sec[:0])
    assert ip.cpu().numpy().tolist() == [0] and cols.shape[0] == 0
    # single state: only the diagonal couples
    e1 = gpu_eloc(t, sec[:1], psi[:1])
    assert rel_err(e1, ct.local_energy(sec[:1], psi[:1])).max() <= ELOC_RTOL and abs(e1[0].imag) == 0
    # ragged sizes around the CTA tile (256 threads x 4 states)
    for m in (31, 33, 255, 1023, 1024, 1025):
        st = np.resize(sec, m)  # rows may repeat (the "with replacement" throughput config of SURVEY.md §8d)
        p = eo.synthetic_psi(m, m)
        tk, tp = sec, psi
        assert rel_err(gpu_eloc(t, st, p, table_keys=tk, table_psi=tp), ct.local_energy(st, p, tk, tp)).max() <= ELOC_RTOL
    # duplicate table keys are summed like scipy's repeated column (SURVEY.md §8 input contract)
    tk = np.concatenate([sec[:100], sec[:10]])
    tp = eo.synthetic_psi(110, 9)
    e = gpu_eloc(t, sec[:100], tp[:100], table_keys=tk, table_psi=tp)
    assert rel_err(e, ct.local_energy(sec[:100], tp[:100], tk, tp)).max() <= 1e-11
    # duplicates_equal: keys may repeat with the SAME amplitude (all-gathered shards) -> one copy kept, not summed
    tk2 = np.concatenate([sec[:100], sec[:37], sec[5:60]])
    tp2 = np.concatenate([tp[:100], tp[:37], tp[5:60]])
    ref_u = ct.local_energy(sec[:100], tp[:100], sec[:100], tp[:100])
    for kind in (nb200._lib.LOOKUP_DENSE, nb200._lib.LOOKUP_HASH):
        t.build_lookup(tk2, tp2, kind=kind, duplicates_equal=True)
        e = nb200._lib.complex_from_pairs(t.local_energy(sec[:100], tp[:100], rebuild_lookup=False))
        assert rel_err(e, ref_u).max() <= ELOC_RTOL
    # a state outside the sector has no stored couplings (its couplings all fail the sector mask)
    bad = np.array([0b111], dtype=np.uint64)
    ip, _, _, _ = t.rows(bad)
    assert ip.cpu().numpy().tolist() == ct.rows(bad)[0].tolist()
    # empty term table
    t0 = nb200.DeviceTermTable(np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0), N, na, nb)
    assert np.all(gpu_eloc(t0, sec[:10], psi[:10]) == 0)
    # E_loc before any lookup table exists is a call-order error, not garbage
    t2, _, _ = make_tables("LiH", True)
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    with pytest.raises(nb200.NaqsError):
        nb200._lib.check(nb200._lib.load().naqs_eloc(t2._h, nb200._lib.ptr(buf), nb200._lib.ptr(buf), 0, 4, nb200._lib.ptr(buf), None))


def test_device_resident_inputs_and_streams():
    nb200, c_oracle, eo = _mods()
    t, ct, (N, na, nb) = make_tables("N2", True)
    st = random_sector_states(N, na, nb, 4096, seed=2)
    psi = eo.synthetic_psi(4096, 3)
    d_st = torch.from_numpy(st.view(np.int64)).cuda()
    d_psi = torch.from_numpy(psi).cuda()
    ref = ct.local_energy(st, psi)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out = t.local_energy(d_st, d_psi)
    s.synchronize()
    assert rel_err(nb200._lib.complex_from_pairs(out), ref).max() <= ELOC_RTOL
    # [M, 2] float32 torch layout, as the reference's optimizer hands psi over (complex.py)
    out2 = t.local_energy(d_st, torch.view_as_real(d_psi))
    assert torch.equal(out, out2)


# ------------------------------------------------------------------------------------------- level-0 twins
def test_level0_twins_match_reference_outputs():
    nb200, _, _ = _mods()
    from scipy.sparse import csr_matrix
    hm, sm, him = nb200.hamiltonian_math, nb200.sparse_math, nb200.hilbert_math
    d = np.load(os.path.join(GOLDEN, "level0.npz"))
    for dt in ("int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64"):
        out = hm.popcount_parity(d[f"pp_in_{dt}"])
        assert out.dtype == np.int8 and np.array_equal(out, d[f"pp_out_{dt}"])
    out = hm.popcount_parity(d["pp_in_1d"])
    assert out.shape == (19, 1) and np.array_equal(out, d["pp_out_1d"])
    with pytest.raises(TypeError):
        hm.popcount_parity(np.zeros(4, np.float32))
    for cd in (np.float64, np.float32):
        H = hm.get_Hij_cy(d["hij_states"], d["hij_uXY"], d["hij_u2aXY"], d["hij_P"], d["hij_u2aYZ"], d["hij_c"].astype(cd))
        assert H.dtype == cd and np.array_equal(H, d[f"hij_out_{np.dtype(cd).name}"])
    n = len(d["mv_indptr"]) - 1
    H = csr_matrix((d["mv_data"], d["mv_indices"], d["mv_indptr"]), shape=(n, n))
    assert np.array_equal(sm.sparse_dense_mv(H, d["mv_v128"]), d["mv_out128"])
    assert np.array_equal(sm.sparse_dense_mv(H, d["mv_v64"]), d["mv_out64"])          # c64 promoted to c128
    assert np.array_equal(sm.sparse_dense_mv(H, d["mv_vreal"]), d["mv_outreal"])      # real promoted to c128
    assert np.array_equal(sm.sparse_dense_mv(H, d["mv_v128"], par=False), d["mv_out_serial"])
    H32 = H.astype(np.float32)
    o = sm.sparse_dense_mv(H32, d["mv_v64"])
    assert o.dtype == np.complex64 and np.array_equal(o, d["mv32_out64"])
    assert np.array_equal(sm.sparse_dense_mv(H32, d["mv_v128"]), d["mv32_out128"])
    assert np.array_equal(sm.sparse_sparse_mv(H, d["ssmv_v"], d["ssmv_idx"]), d["ssmv_out"])
    assert np.array_equal(him.make_basis_idxs_cy(4), d["basis4"]) and np.array_equal(him.make_basis_idxs_cy(9), d["basis9"])


def test_level0_pipeline_equals_fused_rows():
    """The Level-0 chain AND -> popcount_parity -> get_Hij_cy (hamiltonian.py:301-337) on the device equals hij_dense."""
    nb200, c_oracle, eo = _mods()
    xy, yz, c, N, na, nb = load_table("H2O")
    t = nb200.DeviceTermTable(xy, yz, c, N, na, nb)
    st = random_sector_states(N, na, nb, 100, seed=4).astype(np.int64).astype(np.int16)
    uXY, u2aXY = np.unique(xy.astype(np.int64).astype(np.int16), return_inverse=True)
    uYZ, u2aYZ = np.unique(yz.astype(np.int64).astype(np.int16), return_inverse=True)
    P = nb200.hamiltonian_math.popcount_parity(np.bitwise_and(st[:, None], uYZ[None, :]))
    H = nb200.hamiltonian_math.get_Hij_cy(st, uXY, u2aXY, P, u2aYZ, c)
    assert np.array_equal(H, t.hij_dense(st).cpu().numpy().reshape(-1))


def test_state2idx_restricted_index_and_stats():
    nb200, c_oracle, eo = _mods()
    lib, L = nb200._lib.load(), nb200._lib
    rng = np.random.default_rng(0)
    for N in (12, 30, 40, 100):
        W = L.n_words(N)
        s = (rng.integers(0, 2, size=(777, N)).astype(np.int8) * 2 - 1)
        d_s = torch.from_numpy(s).cuda()
        d_k = torch.empty((777, W), dtype=torch.int64, device="cuda")
        L.check(lib.naqs_state2idx(L.ptr(d_s), 777, N, W, L.ptr(d_k), None))
        assert np.array_equal(d_k.cpu().numpy().view(np.uint64), eo.state2idx(s))
    t, ct, (N, na, nb) = make_tables("H2O", True)
    allk = np.arange(2 ** N, dtype=np.uint64)
    assert np.array_equal(t.restricted_index(allk).cpu().numpy(), ct.restricted_index(allk))
    t30, ct30, (N, na, nb) = make_tables("Li2O", True)
    ks = random_sector_states(N, na, nb, 5000, seed=1)
    assert np.array_equal(t30.restricted_index(ks).cpu().numpy(), ct30.restricted_index(ks))
    e = (rng.normal(size=100001) + 1j * rng.normal(size=100001))
    w = rng.random(100001)
    s5 = t.stats(e, w).cpu().numpy()
    exp = np.array([w.sum(), (w * e.real).sum(), (w * e.imag).sum(), (w * e.real ** 2).sum(), len(e)])
    assert np.allclose(s5, exp, rtol=1e-12, atol=0)
    assert np.allclose(t.stats(e).cpu().numpy()[[0, 4]], [len(e), len(e)])


# ------------------------------------------------------------------------------------------- mirror classes
def test_pauli_hamiltonian_mirror_known_answers():
    nb200, _, eo = _mods()
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        known = json.load(f)
    for mol in ("LiH", "H2O", "NH3"):
        xy, yz, c, N, na, nb = load_table(mol)
        # rebuild a terms dict from the packed table is impossible; drive the mirror through its table directly
        hil = nb200.Hilbert.get(N, na, nb, encoding=nb200.Encoding.SIGNED)
        ph = nb200.PauliHamiltonianB200.__new__(nb200.PauliHamiltonianB200)
        _init_mirror_from_table(ph, hil, xy, yz, c)
        sec = hil.get_subspace(ret_states=False, ret_idxs=True)
        H = ph.update_H(sec, check_unseen=False, assume_unique=True)
        assert H.nnz == known[mol]["nnz"] and H.shape == (known[mol]["sector"],) * 2
        assert abs(H.diagonal().sum() - known[mol]["trace"]) <= 1e-12 * abs(known[mol]["trace"])
        assert abs(np.abs(H.data).sum() - known[mol]["sum_abs"]) <= 1e-12 * known[mol]["sum_abs"]
        assert abs(H - H.T).max() == 0
        if mol != "NH3":
            assert abs(np.linalg.eigvalsh(H.toarray())[0] - known[mol]["e0"]) < 1e-9
        # cache semantics: a second update with check_unseen=True adds nothing (hamiltonian.py:294-299)
        assert ph.update_H(sec, check_unseen=True, assume_unique=True).nnz == H.nnz
        sub = sec[:50]
        Hs = ph.get_H(sub)
        assert Hs.shape == (50, 50)
        ph.freeze_H()
        assert ph.is_frozen() and ph.get_restricted_H().nnz == H.nnz


def _init_mirror_from_table(ph, hil, xy, yz, c, device=None):
    """Build a PauliHamiltonianB200 from an already packed table (fixtures hold tables, not Pauli strings)."""
    import naqs_b200
    from scipy.sparse import csr_matrix
    ph.hilbert, ph.dtype, ph.verbose, ph.n_excitations_max, ph.qubit_hamiltonian = hil, np.float64, False, None, None
    ph.restricted_idxs = hil.full2restricted_idx(hil.get_subspace(ret_states=False, ret_idxs=True))
    ph.table = naqs_b200.DeviceTermTable(xy, yz, c, hil.N, hil.N_alpha, hil.N_beta, device)
    ph.H = csr_matrix(([], ([], [])), shape=(hil.size, hil.size), dtype=np.float64)
    ph._cached_idxs = np.array([], dtype=hil.get_idx_dtype("np"))
    ph._frozen_H, ph._restricted_H = False, None


def test_pauli_hamiltonian_from_pauli_strings_and_calculate_local_energy():
    """End to end through the reference-shaped API: Pauli strings -> PauliHamiltonian.get -> calculate_local_energy."""
    import types
    from conftest import load_terms_json
    nb200, c_oracle, eo = _mods()
    case = load_case("LiH_sector")
    N, na, nb = 12, 2, 2
    hil = nb200.Hilbert.get(N, na, nb, encoding=nb200.Encoding.SIGNED)
    op = types.SimpleNamespace(terms=load_terms_json("LiH"))
    sec = hil.get_subspace(ret_states=False, ret_idxs=True)
    ph = nb200.PauliHamiltonian.get(hil, op, restricted_idxs=sec, dtype=np.float64)
    assert (ph.table.K, ph.table.Kxy, ph.table.Kyz) == (631, 84, 252)
    opt = types.SimpleNamespace(pauli_hamiltonian=ph, hilbert=hil)
    idx = torch.from_numpy(case["states"].astype(np.int64).astype(np.int16)).unsqueeze(-1)
    psi_t = torch.view_as_real(torch.from_numpy(case["psi"]))  # float32 [M, 2], as energy.py:310 passes it
    e32 = nb200.calculate_local_energy(opt, idx.squeeze(), psi=psi_t)
    assert e32.dtype == torch.float32 and e32.shape == (len(idx), 2)
    ec = nb200.calculate_local_energy(opt, idx.squeeze(), psi=psi_t, ret_complex=True)
    assert rel_err(ec, case["eloc"]).max() <= ELOC_RTOL
    assert torch.equal(e32, torch.FloatTensor(np.stack([case["eloc"].real, case["eloc"].imag], -1)))
    with pytest.raises(NotImplementedError):
        nb200.calculate_local_energy(opt, idx.squeeze(), psi=psi_t, set_unsampled_states_to_zero=False)
    # update_H rows through the mirror == the reference CSR rows of the fixture
    H = ph.update_H(idx, check_unseen=True, assume_unique=True)
    r = np.asarray(hil.full2restricted_idx(idx.squeeze().numpy())).astype(np.int64)
    got_cols = np.concatenate([H.indices[H.indptr[i]:H.indptr[i + 1]] for i in r])
    got_vals = np.concatenate([H.data[H.indptr[i]:H.indptr[i + 1]] for i in r])
    assert np.array_equal(got_cols, case["rows_cols_restricted"]) and np.array_equal(got_vals, case["rows_vals"])
    assert np.array_equal(np.asarray(ph.get_coupled_state_idxs(r, return_unique=True)), case["coupled_unique_restricted"])
    # the constructor default dtype=np.float32 (hamiltonian.py:48): float32 CSR, bit-identical to the reference's
    case32 = load_case("LiH_sector_f32")
    ph32 = nb200.PauliHamiltonian.get(hil, op, restricted_idxs=sec)
    assert ph32.dtype is np.float32 and ph32.couplings.dtype == np.float32
    idx32 = case32["states"].astype(np.int64).astype(np.int16)
    H32 = ph32.update_H(idx32, check_unseen=True, assume_unique=True)
    assert H32.dtype == np.float32
    r32 = np.asarray(hil.full2restricted_idx(idx32)).astype(np.int64)
    got32 = np.concatenate([H32.data[H32.indptr[i]:H32.indptr[i + 1]] for i in r32])
    assert np.array_equal(got32.astype(np.float64), case32["rows_vals"])
    e = ph32.local_energy(idx32, case32["psi"])
    assert np.abs(e - case32["eloc"]).max() <= 2e-5 * np.abs(case32["eloc"]).max()
    with pytest.raises(TypeError):
        nb200.PauliHamiltonian.get(hil, op, restricted_idxs=sec, dtype=np.longdouble)


# ------------------------------------------------------------------------------------------- full-size properties
def test_full_size_properties_n2_1e6():
    """BASELINE config 3 at full size (N2, M = 1e6 distinct keys of the 2^20 space): size-independent properties
    + an oracle check on a sub-sample."""
    nb200, c_oracle, eo = _mods()
    xy, yz, c, N, _, _ = load_table("N2")
    t, ct = nb200.DeviceTermTable(xy, yz, c, N), c_oracle.COracleTable(xy, yz, c, N)
    rng = np.random.default_rng(0)
    st = rng.choice(2 ** N, 1_000_000, replace=False).astype(np.uint64)
    psi = eo.synthetic_psi(len(st), 0)
    e = gpu_eloc(t, st, psi)
    assert np.all(np.isfinite(e.view(np.float64)))
    # (1) scale invariance: E_loc(a psi) = E_loc(psi) for a power-of-two scale (exact in floating point)
    assert np.array_equal(gpu_eloc(t, st, psi * np.complex64(4.0)), e)
    # (2) permutation equivariance
    perm = rng.permutation(len(st))
    assert np.array_equal(gpu_eloc(t, st[perm], psi[perm]), e[perm])
    # (3) Hermiticity: sum_s |psi_s|^2 E_loc(s)^* = <psi|H|psi> restricted to the batch is real
    p = psi.astype(np.complex128)
    tot = np.sum(np.abs(p) ** 2 * np.conj(e))
    assert abs(tot.imag) <= 1e-9 * abs(tot.real)
    # (4) oracle on a sub-sample of rows against the full table
    sub = rng.choice(len(st), 3000, replace=False)
    ref = ct.local_energy(st[sub], psi[sub], st, psi)
    assert rel_err(e[sub], ref).max() <= ELOC_RTOL
    # (4b) unique-key contract -> complex64 dense table: same numbers to rounding, still within 1e-12 of the oracle
    e_u = gpu_eloc(t, st, psi, assume_unique=True)
    assert rel_err(e_u[sub], ref).max() <= ELOC_RTOL and rel_err(e_u, e).max() <= 1e-13
    # (5) hash lookup (row-order walk) == dense lookup (key-order walk) at full size
    assert rel_err(gpu_eloc(t, st, psi, kind=nb200._lib.LOOKUP_HASH), e).max() <= 1e-13


def test_full_size_li2o_1e5():
    """BASELINE config 4 batch (Li2O, 30 qubits, K = 20 558 > one smem tile, M = 1e5 sector states)."""
    nb200, c_oracle, eo = _mods()
    t, ct, (N, na, nb) = make_tables("Li2O", True)
    st = random_sector_states(N, na, nb, 100_000, seed=0)
    psi = eo.synthetic_psi(len(st), 1)
    e = gpu_eloc(t, st, psi)
    sub = np.random.default_rng(1).choice(len(st), 1500, replace=False)
    ref = ct.local_energy(st[sub], psi[sub], st, psi)
    assert rel_err(e[sub], ref).max() <= ELOC_RTOL
    assert np.array_equal(gpu_eloc(t, st, psi * np.complex64(0.5)), e)


# ------------------------------------------------------------------------------------------- multi-GPU (needs >= 2 GPUs)
def _mgpu_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import naqs_b200
        from naqs_b200 import distributed as nd
        xy, yz, c, N, na, nb = load_table("N2")
        t = naqs_b200.DeviceTermTable(xy, yz, c, N, na, nb, device=f"cuda:{rank}")
        st = random_sector_states(N, na, nb, 5001, seed=77)  # odd size: uneven shards
        rng = np.random.default_rng(78)
        psi = (rng.normal(size=len(st)) + 1j * rng.normal(size=len(st))).astype(np.complex64)
        lo, hi = nd.shard_bounds(len(st), world, rank)
        eloc, stats = nd.sharded_local_energy_stats(t, st[lo:hi], psi[lo:hi])
        # same through the all-reduced dense table (psi is a function of the state: duplicates_equal)
        eloc2, sums2 = nd.sharded_local_energy(t, st[lo:hi], psi[lo:hi], duplicates_equal=True)
        assert nd.can_allreduce_table(t, torch.from_numpy(psi))
        diff = float((eloc2 - eloc).abs().max() / eloc.abs().max())
        assert diff < 1e-13, diff
        # uneven shards through the all-GATHER exchange with keep-one builds (a padded (key, 0) pair must never
        # replace a real amplitude): hash lookup with complex64 psi, and complex128 psi (which cannot take the all-reduced table)
        table_h = naqs_b200.DeviceTermTable(xy, yz, c, N, na, nb, device=f"cuda:{rank}")
        g_k, g_p, total = nd.gather_table(table_h._keys(st[lo:hi]), torch.from_numpy(psi[lo:hi]).to(table_h.device))
        assert total == len(st) and g_k.shape[0] == len(st)
        table_h.build_lookup(g_k, g_p, kind=naqs_b200.table.LOOKUP_HASH, duplicates_equal=True)
        eloc3 = table_h.local_energy(st[lo:hi], psi[lo:hi], rebuild_lookup=False)
        eloc4, _ = nd.sharded_local_energy(t, st[lo:hi], psi[lo:hi].astype(np.complex128), duplicates_equal=True)
        for other in (eloc3, eloc4):
            diff = float((other - eloc).abs().max() / eloc.abs().max())
            assert diff < 1e-13, diff
        # the C-ABI exchange (naqs_comm_* / naqs_table_exchange / naqs_stats_allreduce): push kernels over peer memory for this
        # 20-qubit table, then the NCCL all-gather path with uneven shards (padding must not reach the table)
        comm = nd.Comm(t.device)
        t_push = naqs_b200.DeviceTermTable(xy, yz, c, N, na, nb, device=f"cuda:{rank}")
        sizes = [nd.shard_bounds(len(st), world, r) for r in range(world)]
        max_local = max(b - a for a, b in sizes)
        modes = (0, 0x2000, 0x4000, 0x8000, 0x8000, 0x2000, 0x8000)  # auto, push, NCCL all-reduce, merge, merge, push, merge
        for rep in range(len(modes)):  # several epochs: the two peer tables alternate and are cleared in between, whatever the mode
            ps = psi * np.complex64(2.0 ** rep)  # exact in float32: E_loc is scale invariant
            e5, s5 = nd.sharded_local_energy_comm(t_push, comm, st[lo:hi], ps[lo:hi], flags=modes[rep])
            diff = float((e5 - eloc).abs().max() / eloc.abs().max())
            assert diff < 1e-13, (rep, diff)
            st5 = nd.stats_from_sums(s5.cpu().numpy())
            assert st5["n"] == len(st) and abs(st5["mean"] - stats["mean"]) <= 1e-12 * abs(stats["mean"])
        from naqs_b200 import _lib as L
        e6, s6 = nd.sharded_local_energy_comm(table_h, comm, st[lo:hi], psi[lo:hi], max_local=max_local, flags=0x1000 | naqs_b200.table.LOOKUP_HASH)
        diff = float((e6 - eloc).abs().max() / eloc.abs().max())
        assert diff < 1e-13, diff
        table_h.check()  # the out-of-range padding keys of the shorter shard are not reported
        assert nd.stats_from_sums(s6.cpu().numpy())["n"] == len(st)
        # the default for large key spaces: push-gather over peer memory (push_pairs_kernel).  First a small batch (slot capacity
        # 1024), then the full uneven shards (the region is re-created, collectively), several epochs, complex64 and complex128
        sub = st[:600]
        slo, shi = nd.shard_bounds(len(sub), world, rank)
        e_sub, s_sub = nd.sharded_local_energy_comm(table_h, comm, sub[slo:shi], psi[:600][slo:shi], max_local=300, flags=naqs_b200.table.LOOKUP_HASH)
        ref_sub = naqs_b200.DeviceTermTable(xy, yz, c, N, na, nb, device=f"cuda:{rank}").local_energy(sub, psi[:600])[slo:shi]
        assert float((e_sub - ref_sub).abs().max() / ref_sub.abs().max()) < 1e-13
        for rep in range(4):
            ps = (psi * np.complex64(2.0 ** rep)).astype(np.complex64 if rep % 2 == 0 else np.complex128)
            e7, s7 = nd.sharded_local_energy_comm(table_h, comm, st[lo:hi], ps[lo:hi], max_local=max_local, flags=naqs_b200.table.LOOKUP_HASH)
            diff = float((e7 - eloc).abs().max() / eloc.abs().max())
            assert diff < 1e-13, (rep, diff)
            table_h.check()
            assert nd.stats_from_sums(s7.cpu().numpy())["n"] == len(st)
        comm.close()
        q.put((rank, lo, hi, naqs_b200._lib.complex_from_pairs(eloc), stats))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_local_energy_two_gpus():
    import socket
    import torch.multiprocessing as mp
    nb200, c_oracle, eo = _mods()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mgpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    xy, yz, c, N, na, nb = load_table("N2")
    st = random_sector_states(N, na, nb, 5001, seed=77)
    rng = np.random.default_rng(78)
    psi = (rng.normal(size=len(st)) + 1j * rng.normal(size=len(st))).astype(np.complex64)
    ref = c_oracle.COracleTable(xy, yz, c, N, na, nb).local_energy(st, psi)
    got = np.concatenate([r[3] for r in res])
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == len(st)
    assert rel_err(got, ref).max() <= ELOC_RTOL
    for r in res:  # every rank holds the same globally reduced statistics
        assert r[4]["n"] == len(st) and abs(r[4]["mean"] - ref.mean()) <= 1e-12 * abs(ref.mean())


def test_matrix_free_apply_h_and_ground_state():
    """H.v without forming H.  (H v) equals the oracle's CSR mat-vec; Lanczos on the operator reproduces the
    FCI ground-state energies of the known-answer table."""
    import types
    from conftest import load_terms_json
    nb200, c_oracle, eo = _mods()
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        known = json.load(f)
    for mol in ("LiH", "H2O"):
        xy, yz, c, N, na, nb = load_table(mol)
        t, ct = nb200.DeviceTermTable(xy, yz, c, N, na, nb), c_oracle.COracleTable(xy, yz, c, N, na, nb)
        sec = eo.sector_keys(N, na, nb)
        v = np.random.default_rng(0).normal(size=len(sec)) + 1j * np.random.default_rng(1).normal(size=len(sec))
        got = nb200._lib.complex_from_pairs(t.apply_H(sec, v))
        indptr, cols, vals = ct.rows(sec)
        pos = {int(k): i for i, k in enumerate(sec[:, 0])}
        ref = np.array([sum(vals[e] * v[pos[int(cols[e, 0])]] for e in range(indptr[m], indptr[m + 1])) for m in range(len(sec))])
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    hil = nb200.Hilbert.get(12, 2, 2, encoding=nb200.Encoding.SIGNED)
    ph = nb200.PauliHamiltonian.get(hil, types.SimpleNamespace(terms=load_terms_json("LiH")),
                                    restricted_idxs=hil.get_subspace(ret_states=False, ret_idxs=True), dtype=np.float64)
    e0, _ = ph.solve_H(hil.get_subspace(ret_states=False, ret_idxs=True).numpy())
    assert abs(e0[0] - known["LiH"]["e0"]) < 1e-8


def test_local_energy_with_unsampled_amplitudes():
    """E_loc with psi evaluated on the deduplicated coupled set (the mode the reference leaves
    unimplemented).  For a batch that is a strict subset of the sector, it must equal (H psi)[s] / psi[s] over the FULL sector."""
    import types
    from conftest import load_terms_json
    nb200, c_oracle, eo = _mods()
    N, na, nb = 12, 2, 2
    hil = nb200.Hilbert.get(N, na, nb, encoding=nb200.Encoding.SIGNED)
    sec = hil.get_subspace(ret_states=False, ret_idxs=True).numpy().astype(np.int64)
    ph = nb200.PauliHamiltonian.get(hil, types.SimpleNamespace(terms=load_terms_json("LiH")), restricted_idxs=sec, dtype=np.float64)
    rng = np.random.default_rng(4)
    psi_all = (rng.normal(size=len(sec)) + 1j * rng.normal(size=len(sec))).astype(np.complex64)
    lut = {int(k): psi_all[i] for i, k in enumerate(sec)}
    batch = rng.permutation(len(sec))[:40]
    calls = []

    def psi_fn(keys):
        calls.append(len(keys))
        return np.array([lut[int(k)] for k in keys], dtype=np.complex64)

    e = ph.local_energy_full(sec[batch], psi_all[batch], psi_fn)
    xy, yz, c, *_ = load_table("LiH")
    ref = c_oracle.COracleTable(xy, yz, c, N, na, nb).local_energy(sec[batch].astype(np.uint64), psi_all[batch], sec.astype(np.uint64), psi_all)
    assert len(calls) == 1 and calls[0] <= len(sec)
    assert rel_err(e, ref).max() <= ELOC_RTOL


@pytest.mark.parametrize("N,K,M", [(2, 3, 4), (4, 15, 16), (5, 40, 20), (21, 800, 3000), (22, 800, 3000), (23, 800, 3000), (31, 600, 2000),
                                   (32, 600, 2000), (33, 600, 2000)])
def test_boundary_widths_and_tiny_spaces(N, K, M):
    """Key widths around every dispatch boundary (dense <-> hash at 22/23 qubits, 32-bit <-> 64-bit masks at 32/33, nibble
    counts 5/8/16) and spaces so small that the whole space is the batch (key-order walk with a handful of keys)."""
    nb200, c_oracle, eo = _mods()
    xy, yz, c = eo.synthetic_table(N, K, seed=1000 + N)
    M = min(M, 2 ** N)
    if 2 ** N <= 4 * M:
        st = np.random.default_rng(N).permutation(2 ** N)[:M].astype(np.uint64)
    else:
        st = eo.synthetic_states(N, M, seed=N)[:, 0]
    psi = eo.synthetic_psi(len(st), seed=N)
    t, ct = nb200.DeviceTermTable(xy, yz, c, N), c_oracle.COracleTable(xy, yz, c, N)
    ref = ct.local_energy(st, psi)
    for kw in ({}, {"assume_unique": True}, {"kind": nb200._lib.LOOKUP_HASH}):
        assert rel_err(gpu_eloc(t, st, psi, **kw), ref).max() <= ELOC_RTOL, kw
    t.set_algo("direct")
    assert rel_err(gpu_eloc(t, st, psi), ref).max() <= ELOC_RTOL
    indptr, cols, _, vals = t.rows(st[:64], with_restricted_index=False)
    i2, c2, v2 = ct.rows(st[:64])
    assert np.array_equal(indptr.cpu().numpy(), i2) and np.array_equal(cols.cpu().numpy().view(np.uint64), c2) and np.array_equal(vals.cpu().numpy(), v2)
