"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Everything written here comes out of the unmodified reference classes/functions
(src/optimizer/hamiltonian.py, src/utils/hilbert.py, src_cpp/*.pyx compiled by
oracle/build_ref.py), driven through oracle/ref_harness.py.  Inputs are seeded.

Files:
  tables/<mol>.npz        packed Pauli sum in reference term order: xy, yz (uint64), coeff (float64),
                          meta = (n_qubits, n_alpha, n_beta); from __calc_coupling_info
                          (hamiltonian.py:373-430) + the np.unique group counts (:248-249)
  known_answers.json      full-sector H fingerprints: nnz, trace, sum|H|, E0
  eloc_<case>.npz         seeded batches: states, psi (complex64), eloc (complex128) from
                          energy.py:245-248; for small cases also the stored CSR rows of H.
                          *_f32 cases: the same with PauliHamiltonian.get(dtype=np.float32), the constructor
                          default (float32 H_ij, E_loc computed by the reference in complex64); h_bits = 32
  level0.npz              input/output pairs of the five Cython entry points
  terms_<mol>.json        the raw Pauli strings of H2 / LiH (inputs for the pack_terms tests; the pickles cannot travel)
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TABLE_MOLS = ["H2", "LiH", "H2O", "NH3", "N2", "N2_1.5", "N2_2.25", "C2", "H2S", "Li2O"]
KNOWN_MOLS = ["H2", "LiH", "H2O", "NH3", "N2", "C2"]


def reference_table(mol):
    """Call the reference's __calc_coupling_info on a minimal stand-in for `self` (it only reads
    hilbert.N / N_occ / _idx_basis_vec / to_idx_array, qubit_hamiltonian.terms, n_excitations_max,
    dtype) so that even Li2O is packed by the reference code without building its 2^30 LUT."""
    import torch
    mods = rh.reference_modules()
    N, na, nb = rh.MOLECULES[mol]
    hil = types.SimpleNamespace(N=N, N_occ=0, _idx_basis_vec=torch.tensor([2 ** n for n in range(N)], dtype=torch.int64),
                                to_idx_array=lambda t: np.asarray(t).astype(np.int64))
    me = types.SimpleNamespace(hilbert=hil, qubit_hamiltonian=rh.load_operator(mol), n_excitations_max=None, dtype=np.float64)
    cls = mods.hamiltonian._PauliHamiltonianDynamic
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        xy, yz, c = cls._PauliHamiltonianDynamic__calc_coupling_info(me)
    return xy.astype(np.uint64), yz.astype(np.uint64), c.squeeze().astype(np.float64)


def make_tables():
    os.makedirs(os.path.join(OUT, "tables"), exist_ok=True)
    for mol in TABLE_MOLS:
        xy, yz, c = reference_table(mol)
        N, na, nb = rh.MOLECULES[mol]
        np.savez_compressed(os.path.join(OUT, "tables", f"{mol}.npz"), xy=xy, yz=yz, coeff=c,
                            meta=np.array([N, na, nb], np.int64),
                            n_unique=np.array([len(np.unique(xy)), len(np.unique(yz))], np.int64))
        print(f"table {mol}: K={len(c)} Kxy={len(np.unique(xy))} Kyz={len(np.unique(yz))}")


def make_known_answers():
    from scipy.sparse.linalg import eigsh
    out = {}
    for mol in KNOWN_MOLS:
        hil, ph = rh.make_reference(mol)
        sec = hil.get_subspace(ret_states=False, ret_idxs=True).numpy()
        ph.update_H(sec, check_unseen=False, assume_unique=True)
        H = ph.get_H()
        e0 = eigsh(H.astype(np.float64), k=1, which="SA", tol=1e-12)[0][0] if H.shape[0] > 8 else np.linalg.eigvalsh(H.toarray())[0]
        out[mol] = {"n_qubits": hil.N, "sector": int(H.shape[0]), "nnz": int(H.nnz), "trace": float(H.diagonal().sum()),
                    "sum_abs": float(np.abs(H.data).sum()), "e0": float(e0),
                    "symmetric": bool(abs(H - H.T).max() == 0)}
        print("known", mol, out[mol])
    with open(os.path.join(OUT, "known_answers.json"), "w") as f:
        json.dump(out, f, indent=1)


def _psi(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)


# Li2O (30 qubits): the reference's own _HilbertRestricted — 41 409 225 sector states, a 2^30-entry full2restricted LUT, about 35 GB
# and ten minutes of host time; paid once (VERDICT r1 "weak" 4): `python tests/golden/make_golden.py li2o`
LI2O_CASES = [("Li2O_500", "Li2O", True, 500, 19, False)]

F32_CASES = [  # the constructor-default dtype=np.float32 (hamiltonian.py:48): float32 H_ij, complex64 E_loc
    ("LiH_sector_f32", "LiH", True, None, 21, True, np.float32),
    ("H2O_300_f32", "H2O", True, 300, 22, True, np.float32),
    ("N2_full_1000_f32", "N2", False, 1000, 23, False, np.float32),
]


def make_eloc(cases=None):
    mods = rh.reference_modules()
    cases = cases or [  # name, molecule, restricted, batch size (None = sector minus one, avoids quirk q1), seed, with_rows
        ("LiH_sector", "LiH", True, None, 11, True),
        ("LiH_small", "LiH", True, 50, 12, True),
        ("H2O_sector", "H2O", True, None, 13, True),
        ("NH3_1000", "NH3", True, 1000, 14, False),
        ("N2_2000", "N2", True, 2000, 15, False),
        ("N2_1.5_500", "N2_1.5", True, 500, 16, False),
        ("N2_full_3000", "N2", False, 3000, 17, False),
        ("LiH_full_600", "LiH", False, 600, 18, True),
    ]
    for name, mol, restricted, m, seed, with_rows, *rest in cases:
        h_dtype = rest[0] if rest else np.float64
        hil, ph = rh.make_reference(mol, restricted=restricted, dtype=h_dtype)
        rng = np.random.default_rng(seed)
        if restricted:
            sec = hil.get_subspace(ret_states=False, ret_idxs=True).numpy()
            m_ = len(sec) - 1 if m is None else m
            st = sec[rng.permutation(len(sec))[:m_]]
        else:
            st = rng.choice(2 ** hil.N, m, replace=False).astype(hil._idx_np_dtype)
        psi = _psi(len(st), seed + 100)
        eloc = rh.reference_local_energy(ph, st, psi)
        assert eloc.dtype == (np.complex64 if h_dtype is np.float32 else np.complex128), eloc.dtype
        d = dict(states=st.astype(np.int64).astype(np.uint64), psi=psi, eloc=eloc.astype(np.complex128),
                 h_bits=np.array(np.dtype(h_dtype).itemsize * 8, np.int64),
                 meta=np.array([hil.N, rh.MOLECULES[mol][1] if restricted else -1, rh.MOLECULES[mol][2] if restricted else -1], np.int64))
        if with_rows:
            H = ph.get_H()
            ridx = np.asarray(hil.full2restricted_idx(st)).astype(np.int64)
            rows = [H.indices[H.indptr[r]:H.indptr[r + 1]] for r in ridx]
            vals = [H.data[H.indptr[r]:H.indptr[r + 1]] for r in ridx]
            d["rows_indptr"] = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
            cols = np.concatenate(rows).astype(np.int64)
            d["rows_cols_restricted"] = cols
            d["rows_cols_keys"] = (np.asarray(hil.restricted2full_idx(cols)).astype(np.int64).astype(np.uint64)
                                   if restricted else cols.astype(np.uint64))
            assert H.dtype == np.dtype(h_dtype), H.dtype
            d["rows_vals"] = np.concatenate(vals).astype(np.float64)   # float32 values are exact in float64
            d["coupled_unique_restricted"] = np.asarray(ph.get_coupled_state_idxs(ridx, return_unique=True)).astype(np.int64)
        np.savez_compressed(os.path.join(OUT, f"eloc_{name}.npz"), **d)
        print("eloc", name, len(st), "E[0]=", eloc[0])


def make_level0():
    mods = rh.reference_modules()
    hm, sm, him = mods.hamiltonian_math, mods.sparse_math, mods.hilbert_math
    rng = np.random.default_rng(5)
    d = {}
    for dt in ("int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64"):
        info = np.iinfo(dt)
        x = rng.integers(info.min, info.max, size=(37, 5), dtype=dt, endpoint=True)
        d[f"pp_in_{dt}"] = x
        d[f"pp_out_{dt}"] = hm.popcount_parity(x)
    x1 = rng.integers(0, 2 ** 31 - 1, size=19, dtype=np.int32)
    d["pp_in_1d"], d["pp_out_1d"] = x1, hm.popcount_parity(x1)
    # get_Hij_cy on LiH with the reference's own groupings
    hil, ph = rh.make_reference("LiH")
    sec = hil.get_subspace(ret_states=False, ret_idxs=True).numpy()
    st = sec[rng.permutation(len(sec))[:40]]
    P = hm.popcount_parity(np.bitwise_and(st[:, None], ph._unique_YZ_sites_idx[None, :]))
    for cd in (np.float64, np.float32):
        Hij = hm.get_Hij_cy(st, ph._unique_XY_sites_idx, ph._unique2all_XY_sites_idx, P, ph._unique2all_YZ_sites_idx,
                            ph.couplings.squeeze().astype(cd))
        d[f"hij_out_{np.dtype(cd).name}"] = Hij
    d["hij_states"], d["hij_uXY"], d["hij_u2aXY"] = st, ph._unique_XY_sites_idx, ph._unique2all_XY_sites_idx
    d["hij_P"], d["hij_u2aYZ"], d["hij_c"] = P, ph._unique2all_YZ_sites_idx, ph.couplings.squeeze()
    # sparse mat-vecs on the LiH sector Hamiltonian
    ph.update_H(sec, check_unseen=False, assume_unique=True)
    H = ph.get_H().tocsr()
    d["mv_data"], d["mv_indices"], d["mv_indptr"] = H.data, H.indices, H.indptr
    v128 = (rng.normal(size=H.shape[0]) + 1j * rng.normal(size=H.shape[0])).astype(np.complex128)
    d["mv_v128"], d["mv_out128"] = v128, sm.sparse_dense_mv(H, v128)
    d["mv_v64"], d["mv_out64"] = v128.astype(np.complex64), sm.sparse_dense_mv(H, v128.astype(np.complex64))
    d["mv_vreal"], d["mv_outreal"] = v128.real.copy(), sm.sparse_dense_mv(H, v128.real.copy())
    d["mv_out_serial"] = sm.sparse_dense_mv(H, v128, par=False)
    H32 = H.astype(np.float32)
    d["mv32_out64"] = sm.sparse_dense_mv(H32, v128.astype(np.complex64))
    d["mv32_out128"] = sm.sparse_dense_mv(H32, v128)
    vi = rng.permutation(H.shape[0])[:60].astype(np.int32)
    vv = (rng.normal(size=60) + 1j * rng.normal(size=60)).astype(np.complex128)
    d["ssmv_idx"], d["ssmv_v"], d["ssmv_out"] = vi, vv, sm.sparse_sparse_mv(H, vv, vi)
    d["basis4"] = him.make_basis_idxs_cy(4)
    d["basis9"] = him.make_basis_idxs_cy(9)
    np.savez_compressed(os.path.join(OUT, "level0.npz"), **d)
    print("level0 done")


def make_terms_json():
    for mol in ("H2", "LiH"):
        terms = rh.load_terms(mol)
        out = [[[[int(q), p] for q, p in t], float(c.real), float(c.imag)] for t, c in terms.items()]
        with open(os.path.join(OUT, f"terms_{mol}.json"), "w") as f:
            json.dump(out, f)


if __name__ == "__main__":
    which = sys.argv[1:] or ["tables", "known", "eloc", "level0", "terms"]
    if "terms" in which:
        make_terms_json()
    if "tables" in which:
        make_tables()
    if "known" in which:
        make_known_answers()
    if "eloc" in which:
        make_eloc()
    if "eloc" in which or "eloc_f32" in which:
        make_eloc(F32_CASES)
    if "level0" in which:
        make_level0()
    if "li2o" in which:
        make_eloc(LI2O_CASES)
