"""Drop-in for the reference's compiled module `src.utils.hilbert_math` (src_cpp/hilbert_math.pyx)."""
import numpy as np
import torch

from . import _lib


def make_basis_idxs_cy(N, dtype=np.int32):
    """out[i, j] = i & (1 << j), int32 [2^N, N] (src_cpp/hilbert_math.pyx:12-44).
    The reference's int64 variant raises (buffer dtype mismatch, quirk q4 of SURVEY.md §8a); here
    dtype=np.int64 returns the int32 result widened, which is what that variant meant to produce."""
    dev = _lib.require_cuda()
    out = torch.empty((2 ** N, N), dtype=torch.int32, device=dev)
    _lib.check(_lib.load().naqs_make_basis_idxs(int(N), _lib.ptr(out), _lib.stream_ptr(dev)), "naqs_make_basis_idxs")
    res = out.cpu().numpy()
    return res.astype(np.int64) if dtype is np.int64 else res
