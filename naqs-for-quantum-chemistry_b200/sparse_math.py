"""Drop-in for the reference's compiled module `src.utils.sparse_math` (src_cpp/sparse_math.pyx):
sparse_dense_mv and sparse_sparse_mv with the reference's dtype promotion rules, on the device.

`sparse_dense_exp_op` is dead and broken in the reference (sparse_math.pyx:413 unpacks 3 of 4
values, :436 reads an uninitialised index) and is deliberately not reproduced.
"""
import numpy as np
import torch

from . import _lib


def __type_mv(m, v):
    """src_cpp/sparse_math.pyx:13-41: (float32 m, complex64 v) or (float64 m, complex128 v)."""
    if m.dtype is np.dtype(np.float64):
        n_bit = 64
    elif m.dtype is np.dtype(np.float32):
        n_bit = 32
    else:
        raise Exception("m must have dtype of np.float32 or np.float64.")
    if not np.iscomplexobj(v):
        v = v.astype(np.complex64 if n_bit == 32 else np.complex128)
    else:
        if (v.dtype is np.dtype(np.complex128)) and n_bit == 32:
            n_bit = 64
            m = m.astype(np.float64)
        elif (v.dtype is np.dtype(np.complex64)) and n_bit == 64:
            v = v.astype(np.complex128)
    return m, v, n_bit, m.indices.dtype


_type_mv = __type_mv


def _csr_to_device(m, dev):
    idx_dtype = np.int32 if m.indices.dtype == np.int32 else np.int64
    data = torch.from_numpy(np.ascontiguousarray(m.data)).to(dev)
    indices = torch.from_numpy(np.ascontiguousarray(m.indices.astype(idx_dtype, copy=False))).to(dev)
    indptr = torch.from_numpy(np.ascontiguousarray(m.indptr.astype(idx_dtype, copy=False))).to(dev)
    return data, indices, indptr, np.dtype(idx_dtype).itemsize


def sparse_dense_mv(m, v, par=None):
    """out[r] = sum_e data[e] * v[indices[e]] over CSR row r (src_cpp/sparse_math.pyx:49-243).
    `par` (the OpenMP switch, :52-54) is accepted and ignored: every row has its own thread."""
    m, v, n_bit, _ = _type_mv(m, np.asarray(v))
    dev = _lib.require_cuda()
    data, indices, indptr, isz = _csr_to_device(m, dev)
    d_v = torch.view_as_real(torch.from_numpy(np.ascontiguousarray(v)).to(dev))
    n_rows = m.shape[1]  # the reference sizes `out` by shape[1] and walks that many rows (:58, :96)
    if n_rows > m.shape[0]:
        raise ValueError("sparse_dense_mv walks m.shape[1] rows; the matrix must have at least that many")
    d_out = torch.zeros((n_rows, 2), dtype=d_v.dtype, device=dev)
    _lib.check(_lib.load().naqs_sparse_dense_mv(_lib.ptr(data), n_bit // 8, _lib.ptr(indices), _lib.ptr(indptr), isz, n_rows,
                                                _lib.ptr(d_v), _lib.ptr(d_out), _lib.stream_ptr(dev)), "naqs_sparse_dense_mv")
    return torch.view_as_complex(d_out).cpu().numpy()


def sparse_sparse_mv(m, v, v_idxs, assume_sorted=False):
    """Row v_idxs[k] of m dotted with the sparse vector (v_idxs, v) (src_cpp/sparse_math.pyx:251-402)."""
    m, v, n_bit, idx_type = _type_mv(m, np.asarray(v))
    v_idxs = np.asarray(v_idxs).astype(idx_type)
    if not assume_sorted:
        sort_args = np.argsort(v_idxs)
        v_idxs = v_idxs[sort_args]
        v = v[sort_args]
        unsort_args = np.zeros_like(sort_args)
        unsort_args[sort_args] = np.arange(len(sort_args))
    dev = _lib.require_cuda()
    data, indices, indptr, isz = _csr_to_device(m, dev)
    d_v = torch.view_as_real(torch.from_numpy(np.ascontiguousarray(v)).to(dev))
    d_vi = torch.from_numpy(np.ascontiguousarray(v_idxs.astype(np.int32 if isz == 4 else np.int64))).to(dev)
    d_out = torch.zeros((len(v_idxs), 2), dtype=d_v.dtype, device=dev)
    _lib.check(_lib.load().naqs_sparse_sparse_mv(_lib.ptr(data), n_bit // 8, _lib.ptr(indices), _lib.ptr(indptr), isz, _lib.ptr(d_v),
                                                 _lib.ptr(d_vi), len(v_idxs), _lib.ptr(d_out), _lib.stream_ptr(dev)),
               "naqs_sparse_sparse_mv")
    out = torch.view_as_complex(d_out).cpu().numpy()
    return out if assume_sorted else out[unsort_args]
