"""CPU oracle for the NAQS local-energy hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

A dependency-free (numpy only) restatement of the reference algorithm.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; the product package never does.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
against fixtures under tests/golden/ that were produced by running the
reference's own code (tests/golden/make_golden.py, reference imported from
/root/reference with its Cython kernels compiled by oracle/build_ref.py), and
against the known-answer ground-state energies of SURVEY.md §8c.

Reference lines restated (paths relative to the reference root):
  pack_terms            src/optimizer/hamiltonian.py:373-430   (__calc_coupling_info)
  group_terms           src/optimizer/hamiltonian.py:248-252   (np.unique(return_inverse))
  popcount_parity       src_cpp/hamiltonian_math.pyx:295-484
  get_hij               src_cpp/hamiltonian_math.pyx:21-34,200-288 (get_Hij_cy)
  idx_dtype             src/utils/hilbert.py:405-410
  state2idx             src/utils/hilbert.py:573-581
  sector_keys           src/utils/hilbert.py:446-469           (__prepare_basis order)
  in_sector/restricted  src/utils/hilbert.py:429-434,607-640   (full2restricted_idx)
  hamiltonian_rows      src/optimizer/hamiltonian.py:301-363   (update_H)
  local_energy          src/optimizer/energy.py:238-261 + src_cpp/sparse_math.pyx:13-41,87-100
  sparse_dense_mv       src_cpp/sparse_math.pyx:49-100
  sparse_sparse_mv      src_cpp/sparse_math.pyx:251-342
  make_basis_idxs       src_cpp/hilbert_math.pyx:12-44

Keys are arrays of shape [n, W] of uint64 words (word 0 = qubits 0..63), W = 1 or 2,
bit q of the key set iff qubit q is occupied (hilbert.py:425,576-577).
"""
from itertools import combinations
from math import comb

import numpy as np

U64 = np.uint64


# --------------------------------------------------------------------------- keys
def n_words(n_qubits):
    return 1 if n_qubits <= 63 else 2


def as_keys(x, W=1):
    """Coerce ints / 1-D integer arrays / [n,W] arrays into [n, W] uint64."""
    if isinstance(x, np.ndarray) and x.ndim == 2 and x.dtype == U64:
        assert x.shape[1] == W
        return x
    if isinstance(x, np.ndarray) and x.dtype != object:
        a = np.asarray(x).astype(np.int64).astype(U64).reshape(-1)
        out = np.zeros((a.size, W), U64)
        out[:, 0] = a
        return out
    vals = [int(v) for v in np.asarray(x, dtype=object).reshape(-1)]
    out = np.zeros((len(vals), W), U64)
    for w in range(W):
        out[:, w] = np.array([(v >> (64 * w)) & 0xFFFFFFFFFFFFFFFF for v in vals], dtype=U64)
    return out


def keys_to_int(keys):
    """[n,W] uint64 -> list of python ints."""
    keys = np.asarray(keys, U64)
    return [sum(int(keys[i, w]) << (64 * w) for w in range(keys.shape[1])) for i in range(keys.shape[0])]


def idx_dtype(n_qubits):
    """hilbert.py:405-410."""
    if n_qubits < 16:
        return np.int16
    if n_qubits < 30:
        return np.int32
    return np.int64


def state2idx(states, n_qubits=None):
    """±1 (SIGNED) or 0/1 occupation rows [n, N] -> [n, W] keys; occupied = value > 0
    (hilbert.py:573-581: clamp_min(0) then dot with 2**q)."""
    s = np.asarray(states)
    N = s.shape[-1] if n_qubits is None else n_qubits
    W = n_words(N)
    out = np.zeros((s.shape[0], W), U64)
    for q in range(N):
        out[:, q // 64] |= (s[:, q] > 0).astype(U64) << U64(q % 64)
    return out


def make_basis_idxs(N):
    """hilbert_math.pyx:12-22: out[i, j] = i & (1 << j), int32 [2^N, N]."""
    i = np.arange(2 ** N, dtype=np.int32)[:, None]
    return (i & (1 << np.arange(N, dtype=np.int32))[None, :]).astype(np.int32)


# --------------------------------------------------------------------------- term table
def pack_terms(terms, n_qubits, n_occ=0, n_excitations_max=None):
    """Pauli strings -> (xy[K,W], yz[K,W], coeff[K] float64), reference term order.

    hamiltonian.py:383-416: XY bit for X|Y, YZ bit for Y|Z, count Y;
    a term that flips a frozen qubit (q < n_occ) or exceeds n_excitations_max is
    dropped; coefficient = (1j**nY).real * coeff, then cast to float64 (the
    imaginary part is discarded by astype, hamiltonian.py:424).
    """
    W = n_words(n_qubits)
    xy, yz, cs = [], [], []
    for term, coeff in terms.items():
        valid, num_exc, num_y = True, 0, 0
        mxy = myz = 0
        for q, p in term:
            if p in ("X", "Y"):
                mxy |= 1 << q
                if p == "Y":
                    num_y += 1
                    myz |= 1 << q
                if q < n_occ:
                    valid = False
                    break
                elif n_excitations_max is not None:
                    num_exc += 1
                    if num_exc > n_excitations_max:
                        valid = False
                        break
            elif p == "Z":
                myz |= 1 << q
        if valid:
            xy.append(mxy)
            yz.append(myz)
            cs.append(complex((1j ** num_y).real * coeff).real)
    return as_keys(np.array(xy, dtype=object), W), as_keys(np.array(yz, dtype=object), W), np.array(cs, np.float64)


def _key_order(keys):
    """Sort order of [n,W] keys as multi-word unsigned integers (most significant word last)."""
    return np.lexsort(tuple(keys[:, w] for w in range(keys.shape[1])))


def group_terms(masks):
    """np.unique(masks, return_inverse=True) for [K,W] keys (hamiltonian.py:248-249):
    returns (unique[Ku,W] ascending, inverse[K])."""
    order = _key_order(masks)
    sm = masks[order]
    new = np.ones(len(sm), bool)
    new[1:] = np.any(sm[1:] != sm[:-1], axis=1)
    gid_sorted = np.cumsum(new) - 1
    inverse = np.empty(len(sm), np.int64)
    inverse[order] = gid_sorted
    return sm[new], inverse


# --------------------------------------------------------------------------- bit helpers
def _parity_words(x):
    """parity (0/1) of popcount of the multi-word value x [..., W]."""
    folded = np.bitwise_xor.reduce(x, axis=-1)
    return (np.bitwise_count(folded) & 1).astype(np.int8)


def popcount_parity(arr):
    """hamiltonian_math.pyx:455-484: 1 - 2*(popcount(x) & 1) as int8, 2-D output
    (1-D input is reshaped to (-1, 1)); unsupported dtypes raise TypeError."""
    arr = np.asarray(arr)
    if arr.ndim == 1:
        arr = arr.reshape(-1, 1)
    if arr.dtype not in (np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64):
        raise TypeError(f"Unsupported array dtype for popcount_parity(...): {arr.dtype}.")
    u = arr.view(np.dtype(f"u{arr.dtype.itemsize}"))
    return (1 - 2 * (np.bitwise_count(u) & 1)).astype(np.int8)


def get_hij(n_states, n_xy, unique2all_xy, parity_by_unique_yz, unique2all_yz, couplings):
    """hamiltonian_math.pyx:31-34: for k ascending:
    H_ij[m*Kxy + g(k)] += parity[m, y(k)] * c_k.  Serial in k => bit-exact sums."""
    c = np.asarray(couplings).reshape(-1)
    H = np.zeros((n_states, n_xy), c.dtype)
    P = np.asarray(parity_by_unique_yz)
    for k in range(len(c)):
        H[:, unique2all_xy[k]] += P[:, unique2all_yz[k]].astype(c.dtype) * c[k]
    return H.reshape(-1)


# --------------------------------------------------------------------------- sector
def _even_odd_masks(n_qubits, W):
    ev = [0] * W
    od = [0] * W
    for q in range(n_qubits):
        if q % 2 == 0:
            ev[q // 64] |= 1 << (q % 64)
        else:
            od[q // 64] |= 1 << (q % 64)
    return np.array(ev, dtype=object).astype(U64), np.array(od, dtype=object).astype(U64)


def in_sector(keys, n_qubits, n_alpha, n_beta):
    """full2restricted_idx(j) >= 0 (hamiltonian.py:321-328): n_alpha bits on even
    qubits and n_beta bits on odd qubits (hilbert.py:448-449).  n_alpha=None => full space."""
    keys = np.asarray(keys, U64)
    if n_alpha is None:
        return np.ones(len(keys), bool)
    ev, od = _even_odd_masks(n_qubits, keys.shape[1])
    na = np.bitwise_count(keys & ev[None, :]).sum(axis=1)
    nb = np.bitwise_count(keys & od[None, :]).sum(axis=1)
    return (na == n_alpha) & (nb == n_beta)


def sector_keys(n_qubits, n_alpha, n_beta):
    """Sector keys in the reference's restricted order (hilbert.py:446-469):
    product(alpha combinations over even qubits, beta combinations over odd qubits),
    alpha outer / beta inner, each in itertools.combinations (lexicographic) order."""
    a = [sum(1 << q for q in c) for c in combinations(range(0, n_qubits, 2), n_alpha)]
    b = [sum(1 << q for q in c) for c in combinations(range(1, n_qubits, 2), n_beta)]
    vals = np.array([x | y for x in a for y in b], dtype=object)
    return as_keys(vals, n_words(n_qubits))


def _lex_rank(positions, n):
    """Rank of the sorted combination `positions` (values in 0..n-1) in
    itertools.combinations(range(n), k) order."""
    k = len(positions)
    r, prev = 0, -1
    for i, p in enumerate(positions):
        for v in range(prev + 1, p):
            r += comb(n - 1 - v, k - 1 - i)
        prev = p
    return r


def restricted_index(keys, n_qubits, n_alpha, n_beta):
    """full2restricted_idx (hilbert.py:429-434,607-640): rank in sector_keys order, -1 outside.
    n_alpha=None => identity on the low word (the _HilbertFull case, hilbert.py:377-378)."""
    keys = np.asarray(keys, U64)
    if n_alpha is None:
        return keys[:, 0].astype(np.int64)
    ok = in_sector(keys, n_qubits, n_alpha, n_beta)
    n_even, n_odd = (n_qubits + 1) // 2, n_qubits // 2
    nb_comb = comb(n_odd, n_beta)
    out = np.full(len(keys), -1, np.int64)
    ints = keys_to_int(keys)
    for i, v in enumerate(ints):
        if ok[i]:
            ea = [q // 2 for q in range(0, n_qubits, 2) if (v >> q) & 1]
            ob = [q // 2 for q in range(1, n_qubits, 2) if (v >> q) & 1]
            out[i] = _lex_rank(ea, n_even) * nb_comb + _lex_rank(ob, n_odd)
    return out


# --------------------------------------------------------------------------- H rows
class TermTable:
    """Packed Pauli sum + the np.unique groupings the reference keeps (hamiltonian.py:241-252)."""

    def __init__(self, xy, yz, coeff, n_qubits, n_alpha=None, n_beta=None, dtype=np.float64):
        """dtype: the reference's `dtype` argument (hamiltonian.py:48,424): couplings are cast to it and H_ij is
        accumulated in it (float64 in the experiments, _base.py:234; float32 is the constructor default)."""
        self.n_qubits, self.n_alpha, self.n_beta = n_qubits, n_alpha, n_beta
        self.W = n_words(n_qubits)
        self.xy, self.yz = as_keys(xy, self.W), as_keys(yz, self.W)
        self.coeff = np.asarray(coeff).reshape(-1).astype(dtype)
        self.unique_xy, self.unique2all_xy = group_terms(self.xy)
        self.unique_yz, self.unique2all_yz = group_terms(self.yz)

    @classmethod
    def from_terms(cls, terms, n_qubits, n_alpha=None, n_beta=None):
        xy, yz, c = pack_terms(terms, n_qubits)
        return cls(xy, yz, c, n_qubits, n_alpha, n_beta)

    @property
    def K(self):
        return len(self.coeff)

    @property
    def Kxy(self):
        return len(self.unique_xy)


def hamiltonian_dense_rows(table, states):
    """update_H up to (and including) get_Hij_cy (hamiltonian.py:301-337).
    Returns (H_ij [M, Kxy] float64, coupled keys [M, Kxy, W]).  Exactly the
    reference's arithmetic: parity per unique YZ mask, then serial accumulation in k."""
    s = as_keys(states, table.W)
    P_bits = s[:, None, :] & table.unique_yz[None, :, :]              # :301
    P = (1 - 2 * _parity_words(P_bits)).astype(np.int8)               # :305
    j = s[:, None, :] ^ table.unique_xy[None, :, :]                   # :313
    H = get_hij(len(s), table.Kxy, table.unique2all_xy, P, table.unique2all_yz, table.coeff)  # :335
    return H.reshape(len(s), table.Kxy), j


def hamiltonian_rows(table, states):
    """The reference's stored rows: sector filter (hamiltonian.py:328) + exact-zero drop of
    the sparse add (hamiltonian.py:363).  Returns CSR-like (indptr[M+1], col_keys[nnz, W],
    vals[nnz]) with columns of each row in ascending unique-XY order."""
    H, j = hamiltonian_dense_rows(table, states)
    M, Kxy = H.shape
    keep = (H != 0.0) & in_sector(j.reshape(M * Kxy, -1), table.n_qubits, table.n_alpha, table.n_beta).reshape(M, Kxy)
    indptr = np.zeros(M + 1, np.int64)
    indptr[1:] = np.cumsum(keep.sum(axis=1))
    return indptr, j[keep], H[keep]


def coupled_state_set(table, states):
    """get_coupled_state_idxs(return_unique=True) (hamiltonian.py:122-132): sorted unique union."""
    _, cols, _ = hamiltonian_rows(table, states)
    if len(cols) == 0:
        return cols
    u, _ = group_terms(cols)
    return u


# --------------------------------------------------------------------------- E_loc
def local_energy(table, states, psi, table_keys=None, table_psi=None):
    """E_loc(s) = conj( sum_{s' in batch} H[s,s'] psi(s') / psi(s) )  (energy.py:247-248).

    psi (complex64 at the reference boundary) is promoted to complex128 because H is
    float64 (sparse_math.pyx:33-37).  Couplings to states absent from the lookup table
    contribute nothing (energy.py:238-249, set_unsampled_states_to_zero=True).
    Duplicate table keys are summed (scipy fancy indexing repeats the column).
    Summation order over the coupled states of a row: ascending restricted index
    (canonical CSR of hamiltonian.py:350,363 + row-order-preserving sub-matrix :94)."""
    s = as_keys(states, table.W)
    # __type_mv (sparse_math.pyx:13-41): float32 matrix x complex64 vector stays complex64, everything else is complex128
    cdt = np.complex64 if (table.coeff.dtype == np.float32 and np.asarray(psi).dtype == np.complex64) else np.complex128
    psi = np.asarray(psi).astype(cdt)
    tk = s if table_keys is None else as_keys(table_keys, table.W)
    tp = psi if table_psi is None else np.asarray(table_psi).astype(cdt)
    lut = {}
    for i, v in enumerate(keys_to_int(tk)):
        lut.setdefault(v, []).append(i)
    indptr, cols, vals = hamiltonian_rows(table, s)
    cols_int = keys_to_int(cols) if len(cols) else []
    ridx = restricted_index(cols, table.n_qubits, table.n_alpha, table.n_beta) if len(cols) else np.zeros(0, np.int64)
    out = np.zeros(len(s), cdt)
    for m in range(len(s)):
        lo, hi = indptr[m], indptr[m + 1]
        order = lo + np.argsort(ridx[lo:hi], kind="stable")
        acc = cdt(0)
        for e in order:
            for t in lut.get(cols_int[e], ()):
                acc = acc + vals[e] * tp[t]
        out[m] = acc
    with np.errstate(divide="ignore", invalid="ignore"):
        return (out / psi).conj()


# --------------------------------------------------------------------------- level-0 mat-vecs
def sparse_dense_mv(data, indices, indptr, v):
    """sparse_math.pyx:87-100: out[r] = sum_i data[indptr[r]+i] * v[indices[indptr[r]+i]]
    (real CSR x complex vector, accumulated in storage order)."""
    v = np.asarray(v)
    out = np.zeros(len(indptr) - 1, v.dtype)
    for r in range(len(indptr) - 1):
        acc = out.dtype.type(0)
        for e in range(indptr[r], indptr[r + 1]):
            acc = acc + data[e] * v[indices[e]]
        out[r] = acc
    return out


def sparse_sparse_mv(data, indices, indptr, v, v_idxs):
    """sparse_math.pyx:251-342 (assume_sorted=False path): for each k, merge-join row
    v_idxs[k] of the CSR with the sorted (v_idxs, v) list; result returned in caller order."""
    v = np.asarray(v)
    v_idxs = np.asarray(v_idxs)
    sort_args = np.argsort(v_idxs)
    vi, vv = v_idxs[sort_args], v[sort_args]
    out = np.zeros(len(vi), v.dtype)
    pos = {int(x): n for n, x in enumerate(vi)}
    for k in range(len(vi)):
        r = int(vi[k])
        acc = out.dtype.type(0)
        for e in range(indptr[r], indptr[r + 1]):
            n = pos.get(int(indices[e]))
            if n is not None:
                acc = acc + data[e] * vv[n]
        out[k] = acc
    unsort = np.zeros_like(sort_args)
    unsort[sort_args] = np.arange(len(sort_args))
    return out[unsort]


# --------------------------------------------------------------------------- synthetic inputs
def synthetic_table(n_qubits, n_terms, seed=0, group_ratio=6.0):
    """random 2-/4-bit flip masks (Kxy ~ K/6, plus the diagonal group),
    YZ = JW-like contiguous run between the flipped qubits + random weight-~N/4 mask, normal fp64
    coefficients.  Y count (|xy & yz|) is forced even so every coefficient is 'real'."""
    rng = np.random.default_rng(seed)
    W = n_words(n_qubits)
    n_groups = max(2, int(n_terms / group_ratio))
    flips = [0]
    for _ in range(n_groups - 1):
        nb = min(2 if rng.random() < 0.3 else 4, n_qubits)
        qs = rng.choice(n_qubits, nb, replace=False)
        flips.append(sum(1 << int(q) for q in qs))
    xy, yz = [], []
    for k in range(n_terms):
        f = flips[k % n_groups] if k >= n_groups else flips[k]
        lo, hi = sorted(int(q) for q in rng.choice(n_qubits, 2, replace=False))
        run = ((1 << hi) - 1) ^ ((1 << lo) - 1)
        rnd = sum(1 << int(q) for q in np.nonzero(rng.random(n_qubits) < 0.25)[0])
        m = (run ^ rnd) & ~f
        ybits = [q for q in range(n_qubits) if (f >> q) & 1]
        ny = 2 * int(rng.integers(0, len(ybits) // 2 + 1))
        for q in ybits[:ny]:
            m |= 1 << q
        xy.append(f)
        yz.append(m)
    coeff = rng.normal(size=n_terms)
    return as_keys(np.array(xy, dtype=object), W), as_keys(np.array(yz, dtype=object), W), coeff


def synthetic_states(n_qubits, n_states, seed=0, weight=None):
    """Random distinct keys of Hamming weight `weight` (default N/2), seeded."""
    rng = np.random.default_rng(seed)
    W = n_words(n_qubits)
    weight = n_qubits // 2 if weight is None else weight
    seen, out = set(), []
    while len(out) < n_states:
        qs = rng.choice(n_qubits, weight, replace=False)
        v = sum(1 << int(q) for q in qs)
        if v not in seen:
            seen.add(v)
            out.append(v)
    return as_keys(np.array(out, dtype=object), W)


def synthetic_psi(n, seed=0):
    """normal + 1j*normal cast to complex64 (the dtype the reference hands over, complex.py:142)."""
    rng = np.random.default_rng(seed)
    return (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)
