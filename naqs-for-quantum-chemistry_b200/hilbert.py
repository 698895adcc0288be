"""Hilbert-space encodings of the hot path (mirror of the parts of src/utils/hilbert.py the E_loc
path touches): idx dtype rule, state <-> key packing, sector enumeration in the reference's restricted
order and full->restricted ranking WITHOUT the 2^N-entry LUT (hilbert.py:429-434 allocates 16 GB at N=30).
"""
from enum import Enum
from itertools import combinations
from math import comb

import numpy as np
import torch


class Encoding(Enum):
    BINARY = 0
    SIGNED = 1


def idx_dtypes(N):
    """(torch dtype, numpy dtype) of state indices (hilbert.py:405-410)."""
    if N < 16:
        return torch.int16, np.int16
    if N < 30:
        return torch.int32, np.int32
    return torch.int64, np.int64


def _lex_ranks(bits, n, k):
    """Vectorised rank of k-combinations in itertools.combinations(range(n), k) order.
    bits: bool [m, n] occupation of the n positions."""
    m = bits.shape[0]
    rank = np.zeros(m, np.int64)
    seen = np.zeros(m, np.int64)
    for p in range(n):
        occ = bits[:, p]
        rem = k - 1 - seen
        add = np.array([comb(n - 1 - p, int(r)) if r >= 0 else 0 for r in range(k)], dtype=np.int64)
        contrib = np.where((~occ) & (seen < k), add[np.clip(rem, 0, k - 1)], 0) if k > 0 else 0
        rank += contrib
        seen += occ
    return rank


class Hilbert:
    """Hilbert.get(N, N_alpha, N_beta, encoding) -> restricted sector; Hilbert.get(N, encoding=...) -> full space
    (hilbert.py:28-37)."""

    @staticmethod
    def get(N, N_alpha=None, N_beta=None, encoding=Encoding.SIGNED, **_ignored):
        return HilbertSpace(N, N_alpha, N_beta, encoding)


class HilbertSpace:
    def __init__(self, N, N_alpha=None, N_beta=None, encoding=Encoding.SIGNED):
        if N > 62:
            raise ValueError("integer state indices need N <= 62 (int64 keys, hilbert.py:405-410); use raw uint64 key arrays beyond")
        self.N, self.N_alpha, self.N_beta = N, N_alpha, N_beta
        self.N_occ = 0
        self.encoding = encoding
        self.restricted = N_alpha is not None
        self.N_up = (N_alpha + N_beta) if self.restricted else None
        self._idx_torch_dtype, self._idx_np_dtype = idx_dtypes(N)
        self._state_torch_dtype, self._state_np_dtype = torch.int8, np.int8
        self.size = comb((N + 1) // 2, N_alpha) * comb(N // 2, N_beta) if self.restricted else 2 ** N
        self._sector_keys = None

    # -- dtype helpers (hilbert.py to_idx_array / to_idx_tensor)
    def to_idx_array(self, idx):
        if torch.is_tensor(idx):
            idx = idx.detach().cpu().numpy()
        return np.asarray(idx).astype(self._idx_np_dtype)

    def to_idx_tensor(self, idx):
        if torch.is_tensor(idx):
            return idx.to(self._idx_torch_dtype)
        return torch.as_tensor(np.asarray(idx).astype(self._idx_np_dtype))

    def get_idx_dtype(self, kind="np"):
        return self._idx_np_dtype if kind == "np" else self._idx_torch_dtype

    # -- sector
    def sector_keys(self):
        """Sector keys (full indices) in restricted order: alpha combinations over even qubits (outer) x beta
        combinations over odd qubits (inner), each in itertools.combinations order (hilbert.py:446-469)."""
        if not self.restricted:
            return np.arange(2 ** self.N, dtype=np.int64)
        if self._sector_keys is None:
            a = np.array([sum(1 << q for q in c) for c in combinations(range(0, self.N, 2), self.N_alpha)], np.int64)
            b = np.array([sum(1 << q for q in c) for c in combinations(range(1, self.N, 2), self.N_beta)], np.int64)
            self._sector_keys = (a[:, None] | b[None, :]).reshape(-1)
        return self._sector_keys

    def get_subspace(self, ret_states=True, ret_idxs=False, use_restricted_idxs=False, **_ignored):
        """(states int8 [size, N] in the encoding, idxs) like hilbert.py:509-571 for the configured sector."""
        keys = self.sector_keys()
        idxs = np.arange(len(keys)) if use_restricted_idxs else keys
        out = []
        if ret_states:
            out.append(self.idx2state(keys))
        if ret_idxs:
            out.append(self.to_idx_tensor(idxs))
        return out[0] if len(out) == 1 else tuple(out)

    # -- encodings
    def state2idx(self, state, use_restricted_idxs=False):
        """+-1 / 0-1 rows -> keys, occupied = value > 0 (hilbert.py:573-581); shape [n, 1] like the reference."""
        s = state.detach().cpu().numpy() if torch.is_tensor(state) else np.asarray(state)
        s = s.reshape(-1, self.N)
        keys = ((s > 0).astype(np.int64) << np.arange(self.N, dtype=np.int64)[None, :]).sum(axis=1)
        if use_restricted_idxs:
            keys = self.full2restricted_idx(keys)
        return self.to_idx_tensor(keys.reshape(-1, 1))

    def idx2state(self, idx, use_restricted_idxs=False):
        k = self.to_idx_array(idx).astype(np.int64).reshape(-1)
        if use_restricted_idxs:
            k = self.sector_keys()[k]
        bits = ((k[:, None] >> np.arange(self.N, dtype=np.int64)[None, :]) & 1).astype(np.int8)
        if self.encoding == Encoding.SIGNED:
            bits = 2 * bits - 1
        return torch.from_numpy(bits)

    def restricted2full_idx(self, idx):
        np_out = not torch.is_tensor(idx)
        k = self.to_idx_array(idx).astype(np.int64)
        full = self.sector_keys()[k] if self.restricted else k
        return self.to_idx_array(full) if np_out else self.to_idx_tensor(full)

    def full2restricted_idx(self, idx):
        """Rank in the restricted order, -1 outside the sector (hilbert.py:607-640); identity on the full space."""
        if idx is None:
            return None
        np_out = not torch.is_tensor(idx)
        k = self.to_idx_array(idx).astype(np.int64)
        if not self.restricted:
            return self.to_idx_array(k) if np_out else self.to_idx_tensor(k)
        shape = k.shape
        k = k.reshape(-1)
        q = np.arange(self.N, dtype=np.int64)
        bits = ((k[:, None] >> q[None, :]) & 1).astype(bool)
        ev, od = bits[:, 0::2], bits[:, 1::2]
        ok = (ev.sum(1) == self.N_alpha) & (od.sum(1) == self.N_beta) & (k >= 0) & (k < 2 ** self.N)
        r = _lex_ranks(ev, ev.shape[1], self.N_alpha) * comb(self.N // 2, self.N_beta) + _lex_ranks(od, od.shape[1], self.N_beta)
        r = np.where(ok, r, -1).reshape(shape)
        return self.to_idx_array(r) if np_out else self.to_idx_tensor(r)
