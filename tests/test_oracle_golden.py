"""Pin the oracle (numpy + C restatements, and the compiled-reference path) against the golden fixtures
the reference's own code produced (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
from conftest import CASE_TABLE, F32_CASE_TABLE, GOLDEN, case_sector, load_case, load_table, load_terms_json

from oracle import c_oracle
from oracle import eloc_oracle as eo
from oracle import ref_harness

ALL_CASES = sorted(CASE_TABLE)
SMALL_CASES = ["LiH_sector", "LiH_small", "LiH_full_600"]


@pytest.mark.parametrize("mol", ["H2", "LiH"])
def test_pack_terms_matches_reference_table(mol):
    xy, yz, c, N, na, nb = load_table(mol)
    oxy, oyz, oc = eo.pack_terms(load_terms_json(mol), N)
    assert np.array_equal(oxy[:, 0], xy) and np.array_equal(oyz[:, 0], yz)
    assert np.array_equal(oc, c)  # bit-exact coefficients


@pytest.mark.parametrize("mol", ["H2", "LiH", "H2O", "NH3", "N2", "N2_1.5", "N2_2.25", "C2", "H2S", "Li2O"])
def test_group_counts(mol):
    xy, yz, c, N, na, nb = load_table(mol)
    n_unique = np.load(os.path.join(GOLDEN, "tables", f"{mol}.npz"))["n_unique"]
    t = eo.TermTable(xy, yz, c, N, na, nb)
    assert (t.Kxy, len(t.unique_yz)) == tuple(int(x) for x in n_unique)
    ct = c_oracle.COracleTable(xy, yz, c, N, na, nb)
    assert ct.G == t.Kxy


@pytest.mark.parametrize("name", SMALL_CASES)
def test_numpy_oracle_eloc_bit_exact(name):
    case = load_case(name)
    xy, yz, c, N, _, _ = load_table(CASE_TABLE[name])
    na, nb = case_sector(case, name)
    e = eo.local_energy(eo.TermTable(xy, yz, c, N, na, nb), case["states"], case["psi"])
    assert np.array_equal(e, case["eloc"])


@pytest.mark.parametrize("name", ALL_CASES)
def test_c_oracle_eloc_bit_exact(name):
    case = load_case(name)
    xy, yz, c, N, _, _ = load_table(CASE_TABLE[name])
    na, nb = case_sector(case, name)
    e = c_oracle.COracleTable(xy, yz, c, N, na, nb).local_energy(case["states"], case["psi"])
    assert np.array_equal(e, case["eloc"])


@pytest.mark.parametrize("name", ["LiH_sector", "LiH_small", "H2O_sector", "LiH_full_600"])
def test_rows_match_reference_csr(name):
    case = load_case(name)
    xy, yz, c, N, _, _ = load_table(CASE_TABLE[name])
    na, nb = case_sector(case, name)
    ct = c_oracle.COracleTable(xy, yz, c, N, na, nb)
    indptr, cols, vals = ct.rows(case["states"])
    assert np.array_equal(indptr, case["rows_indptr"])
    ridx = ct.restricted_index(cols)
    # the reference CSR is canonical (columns ascending in restricted index); ours is in XY order
    for m in range(len(indptr) - 1):
        lo, hi = indptr[m], indptr[m + 1]
        o = np.argsort(ridx[lo:hi], kind="stable")
        assert np.array_equal(ridx[lo:hi][o], case["rows_cols_restricted"][lo:hi])
        assert np.array_equal(cols[lo:hi, 0][o], case["rows_cols_keys"][lo:hi])
        assert np.array_equal(vals[lo:hi][o], case["rows_vals"][lo:hi])  # bit-exact matrix elements
    if name in SMALL_CASES:
        i2, c2, v2 = eo.hamiltonian_rows(eo.TermTable(xy, yz, c, N, na, nb), case["states"])
        assert np.array_equal(i2, indptr) and np.array_equal(c2, cols) and np.array_equal(v2, vals)
    uniq = np.unique(ridx)
    assert np.array_equal(uniq, case["coupled_unique_restricted"])


@pytest.mark.parametrize("name", sorted(F32_CASE_TABLE))
def test_float32_hamiltonian_oracle(name):
    """dtype=np.float32 (hamiltonian.py:48 default): float32 couplings, float32 accumulation of H_ij
    (__inner_int64_float), complex64 mat-vec (sparse_math.pyx:13-41)."""
    case = load_case(name)
    assert int(case["h_bits"]) == 32
    xy, yz, c, N, _, _ = load_table(F32_CASE_TABLE[name])
    na, nb = case_sector(case, name)
    t = eo.TermTable(xy, yz, c, N, na, nb, dtype=np.float32)
    if "rows_vals" in case:
        indptr, cols, vals = eo.hamiltonian_rows(t, case["states"])
        assert vals.dtype == np.float32 and np.array_equal(indptr, case["rows_indptr"])
        ridx = eo.restricted_index(cols, N, na, nb)
        for m in range(len(indptr) - 1):
            lo, hi = indptr[m], indptr[m + 1]
            o = np.argsort(ridx[lo:hi], kind="stable")
            assert np.array_equal(cols[lo:hi, 0][o], case["rows_cols_keys"][lo:hi])
            assert np.array_equal(vals[lo:hi][o].astype(np.float64), case["rows_vals"][lo:hi])  # bit-exact float32 elements
    e = eo.local_energy(t, case["states"], case["psi"])
    assert e.dtype == np.complex64
    scale = np.abs(case["eloc"]).max()
    assert np.abs(e - case["eloc"]).max() <= 2e-5 * scale  # complex64 sums; the order of the adds is the reference's


@pytest.mark.parametrize("mol", ["H2", "LiH", "H2O", "NH3", "N2", "C2"])
def test_known_answers(mol):
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        known = json.load(f)[mol]
    xy, yz, c, N, na, nb = load_table(mol)
    sec = eo.sector_keys(N, na, nb)
    assert len(sec) == known["sector"]
    ct = c_oracle.COracleTable(xy, yz, c, N, na, nb)
    indptr, cols, vals = ct.rows(sec)
    assert int(indptr[-1]) == known["nnz"]
    diag = vals[(cols[:, 0] == np.repeat(sec[:, 0], np.diff(indptr)))]
    assert abs(diag.sum() - known["trace"]) <= 1e-12 * abs(known["trace"])
    assert abs(np.abs(vals).sum() - known["sum_abs"]) <= 1e-12 * known["sum_abs"]
    if known["sector"] <= 15000:
        from scipy.sparse import csr_matrix
        from scipy.sparse.linalg import eigsh
        H = csr_matrix((vals, (np.repeat(np.arange(len(sec)), np.diff(indptr)), ct.restricted_index(cols))), shape=(len(sec),) * 2)
        assert abs(H - H.T).max() == 0
        e0 = np.linalg.eigvalsh(H.toarray())[0] if len(sec) <= 500 else eigsh(H, k=1, which="SA", tol=1e-12)[0][0]
        assert abs(e0 - known["e0"]) < 1e-8  # FCI ground-state energy (Hartree) of the paper


def test_sector_order_and_rank():
    for (N, na, nb) in [(4, 1, 1), (12, 2, 2), (14, 5, 5), (9, 2, 3)]:
        sec = eo.sector_keys(N, na, nb)
        assert np.array_equal(eo.restricted_index(sec, N, na, nb), np.arange(len(sec)))
        ct = c_oracle.COracleTable(np.zeros(1, np.uint64), np.zeros(1, np.uint64), np.ones(1), N, na, nb)
        assert np.array_equal(ct.restricted_index(sec), np.arange(len(sec)))
        allk = np.arange(2 ** N, dtype=np.uint64)
        r = ct.restricted_index(allk)
        assert (r >= 0).sum() == len(sec) and np.array_equal(r, eo.restricted_index(eo.as_keys(allk), N, na, nb))


def test_level0_oracle():
    d = np.load(os.path.join(GOLDEN, "level0.npz"))
    for dt in ("int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64"):
        assert np.array_equal(eo.popcount_parity(d[f"pp_in_{dt}"]), d[f"pp_out_{dt}"])
    assert np.array_equal(eo.popcount_parity(d["pp_in_1d"]), d["pp_out_1d"])
    with pytest.raises(TypeError):
        eo.popcount_parity(np.zeros(3, np.float32))
    H = eo.get_hij(len(d["hij_states"]), len(d["hij_uXY"]), d["hij_u2aXY"], d["hij_P"], d["hij_u2aYZ"], d["hij_c"])
    assert np.array_equal(H, d["hij_out_float64"])
    H32 = eo.get_hij(len(d["hij_states"]), len(d["hij_uXY"]), d["hij_u2aXY"], d["hij_P"], d["hij_u2aYZ"], d["hij_c"].astype(np.float32))
    assert np.array_equal(H32, d["hij_out_float32"])
    out = eo.sparse_dense_mv(d["mv_data"], d["mv_indices"], d["mv_indptr"], d["mv_v128"])
    assert np.array_equal(out, d["mv_out128"]) and np.array_equal(out, d["mv_out_serial"])
    out = eo.sparse_sparse_mv(d["mv_data"], d["mv_indices"], d["mv_indptr"], d["ssmv_v"], d["ssmv_idx"])
    assert np.array_equal(out, d["ssmv_out"])
    assert np.array_equal(eo.make_basis_idxs(4), d["basis4"]) and np.array_equal(eo.make_basis_idxs(9), d["basis9"])


def test_state2idx_roundtrip():
    rng = np.random.default_rng(0)
    for N in (12, 40, 100):
        s = rng.integers(0, 2, size=(50, N)).astype(np.int8) * 2 - 1
        keys = eo.state2idx(s)
        ints = eo.keys_to_int(keys)
        assert ints == [sum(1 << q for q in range(N) if row[q] > 0) for row in s]


def test_wide_masks_numpy_vs_c():
    """64/128-bit masks, where the reference cannot go: the two restatements must agree bit for bit."""
    for N in (40, 63, 100, 127):
        xy, yz, c = eo.synthetic_table(N, 300, seed=N)
        st = eo.synthetic_states(N, 64, seed=N + 1)
        psi = eo.synthetic_psi(64, seed=N + 2)
        t, ct = eo.TermTable(xy, yz, c, N), c_oracle.COracleTable(xy, yz, c, N)
        assert np.array_equal(eo.local_energy(t, st, psi), ct.local_energy(st, psi))
        i1, c1, v1 = eo.hamiltonian_rows(t, st)
        i2, c2, v2 = ct.rows(st)
        assert np.array_equal(i1, i2) and np.array_equal(c1, c2) and np.array_equal(v1, v2)


@pytest.mark.skipif(not os.path.isdir(os.path.join(os.path.dirname(ref_harness.__file__), "_ref", "src")), reason="oracle/_ref not built")
def test_reference_path_with_compiled_reference_kernels():
    """oracle/ref_path.py (reference Cython kernels + restated numpy/scipy orchestration) reproduces the fixtures."""
    from oracle import ref_path
    for name in ("LiH_sector", "H2O_sector", "N2_2000", "LiH_full_600"):
        case = load_case(name)
        xy, yz, c, N, _, _ = load_table(CASE_TABLE[name])
        na, nb = case_sector(case, name)
        r = ref_path.ReferencePath(xy, yz, c, N, na, nb)
        e = r.local_energy(case["states"].astype(np.int64), case["psi"])
        assert np.array_equal(e, case["eloc"])


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present (GPU box)")
def test_live_reference_agrees_with_oracle():
    """Where /root/reference exists: run the reference's own classes on a fresh seeded batch."""
    hil, ph = ref_harness.make_reference("H2O")
    xy, yz, c, N, na, nb = load_table("H2O")
    assert np.array_equal(ph.XY_sites_idx.astype(np.int64).astype(np.uint64), xy) and np.array_equal(ph.couplings.squeeze(), c)
    sec = hil.get_subspace(ret_states=False, ret_idxs=True).numpy()
    rng = np.random.default_rng(99)
    st = sec[rng.permutation(len(sec))[:300]]
    psi = eo.synthetic_psi(300, 98)
    e_ref = ref_harness.reference_local_energy(ph, st, psi)
    e_c = c_oracle.COracleTable(xy, yz, c, N, na, nb).local_energy(st, psi)
    assert np.array_equal(e_ref, e_c)
