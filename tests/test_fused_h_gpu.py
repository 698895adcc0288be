"""Matrix elements of the FUSED kernels, extracted column by column with one-hot amplitude vectors through naqs_apply_h,
against the serially summed stored rows (naqs_rows_*, which are bit-equal to the reference CSR, hamiltonian.py:301-363).

The fused kernels sum groups of more than 6 terms in 6-term chunks: this test bounds that re-association
(in units of the last place of the group's largest partial sum) and asserts that the EXACT-ZERO SET — which decides what the
hash walk queues (hamiltonian.py:363 drops H == 0.0) — is the one the serial sum defines."""
import numpy as np
import pytest
import torch
from conftest import load_table, random_sector_states

pytestmark = pytest.mark.gpu


def _fused_columns(nb200, t, keys, col_pos, kind, c64):
    """H[:, col] for every col in col_pos through the fused walk: lookup table = one-hot vector on keys[col]."""
    lib = nb200._lib.load()
    k = t._keys(keys)
    M = k.shape[0]
    out = torch.empty((M, 2), dtype=torch.float64, device=t.device)
    cols = np.empty((M, len(col_pos)))
    for j, cp in enumerate(col_pos):
        v = torch.zeros(M, dtype=torch.complex64 if c64 else torch.complex128, device=t.device)
        v[cp] = 1.0
        t.build_lookup(k, v, kind, assume_unique=True)
        with torch.cuda.device(t.device):
            nb200._lib.check(lib.naqs_apply_h(t._h, nb200._lib.ptr(k), M, nb200._lib.ptr(out), t._stream()), "naqs_apply_h")
        o = out.cpu().numpy()
        assert not o[:, 1].any()
        cols[:, j] = o[:, 0]
    return cols


def _serial_columns(t, keys, col_pos):
    indptr, ck, _, vals = t.rows(keys, with_restricted_index=False)
    indptr, ck, vals = indptr.cpu().numpy(), ck.cpu().numpy()[:, 0].astype(np.uint64), vals.cpu().numpy()
    want = {int(keys[cp]): j for j, cp in enumerate(col_pos)}
    out = np.zeros((len(keys), len(col_pos)))
    stored = np.zeros((len(keys), len(col_pos)), dtype=bool)
    row = np.repeat(np.arange(len(keys)), np.diff(indptr))
    for e in np.nonzero(np.isin(ck, np.array(list(want), dtype=np.uint64)))[0]:
        out[row[e], want[int(ck[e])]] = vals[e]
        stored[row[e], want[int(ck[e])]] = True
    return out, stored


def _group_scale(xy, c, keys, col_pos):
    """(sum |c_k|, number of terms) of the group that couples row i to column j: the magnitude the partial sums of H[i, j] can
    reach and the number of additions behind it."""
    xy0 = xy.reshape(len(xy), -1)[:, 0].astype(np.uint64)
    u, inv, counts = np.unique(xy0, return_inverse=True, return_counts=True)
    tot = np.zeros(len(u))
    np.add.at(tot, inv, np.abs(c))
    flips = keys[:, None] ^ keys[np.asarray(col_pos)][None, :]
    idx = np.searchsorted(u, flips)
    idx[idx >= len(u)] = 0
    hit = u[idx] == flips
    return np.where(hit, tot[idx], 0.0), np.where(hit, counts[idx], 0)


CASES = [("LiH", True, None, 225, "all", "dense", False),      # dense table, per-row walk
         ("H2O", True, None, 441, "all", "hash", False),       # bucketed hash walk, shared-memory filter
         ("H2O", False, None, 1 << 14, 40, "dense", True),     # key-order kernel (complex64 table, full key space)
         ("H2O", False, None, 1 << 14, 24, "dense", False),    # key-order walk of the generic kernel (complex128 table)
         ("N2", True, None, 14400, 40, "dense", False),
         ("N2", True, None, 14400, 24, "hash", False),
         ("Li2O", True, None, 6000, 32, "hash", False)]


@pytest.mark.parametrize("mol,sector,_,m,ncol,kind,c64", CASES)
def test_fused_matrix_elements_zero_set_and_ulp_drift(mol, sector, _, m, ncol, kind, c64):
    import naqs_b200 as nb200
    xy, yz, c, N, na, nb = load_table(mol)
    t = nb200.DeviceTermTable(xy, yz, c, N, na if sector else None, nb if sector else None)
    if sector:
        keys = random_sector_states(N, na, nb, m, seed=5) if mol == "Li2O" else None
        if keys is None:
            from oracle import eloc_oracle as eo
            keys = eo.sector_keys(N, na, nb)[:, 0].astype(np.uint64)
            assert len(keys) == m
    else:
        keys = np.arange(1 << N, dtype=np.uint64)
    rng = np.random.default_rng(11)
    col_pos = list(range(m)) if ncol == "all" else sorted(rng.choice(m, ncol, replace=False).tolist())
    k = nb200._lib.LOOKUP_DENSE if kind == "dense" else nb200._lib.LOOKUP_HASH
    fused = _fused_columns(nb200, t, keys, col_pos, k, c64)
    serial, stored = _serial_columns(t, keys, col_pos)
    # (1) the exact-zero set of the fused path IS the serial one (stored <=> H != 0.0 exactly, hamiltonian.py:363)
    assert np.array_equal(fused != 0.0, stored), f"zero sets differ at {int((( fused != 0.0) != stored).sum())} entries"
    # (2) drift of the chunked sums.  Serial order: n additions, each rounded to <= 0.5 ulp of the running sum (<= sum |c_k|);
    # chunked order: the same number of additions inside the chunks (done on the host, reference order) plus one per chunk.  Two
    # correctly rounded evaluations of the same sum therefore differ by at most 0.5 * (2 n + ceil(n / 6)) ulp of sum |c_k|; what
    # the molecules actually show is far below it (<= 10 ulp for the 100+-term diagonal groups), asserted as a regression bar.
    scale, n_terms = _group_scale(xy, c, keys, col_pos)
    ulps = np.abs(fused - serial)[stored] / np.spacing(scale[stored])
    n_st = n_terms[stored]
    assert np.all(ulps <= 0.5 * (2 * n_st + np.ceil(n_st / 6.0))), "drift beyond the rounding-error bound of the two summation orders"
    assert ulps.max(initial=0.0) <= 16.0, f"max drift {ulps.max():.2f} ulp of sum|c_k|"
    # groups of <= 6 terms come from LUT entries produced by the reference's own serial additions: bit-identical
    xy0 = xy.reshape(len(xy), -1)[:, 0].astype(np.uint64)
    u, counts = np.unique(xy0, return_counts=True)
    small = set(u[counts <= 6].tolist())
    flips = keys[:, None] ^ keys[np.asarray(col_pos)][None, :]
    is_small = np.isin(flips, np.array(sorted(small), dtype=np.uint64))
    assert np.array_equal(fused[is_small & stored], serial[is_small & stored])
