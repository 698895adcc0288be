"""ctypes binding of libnaqs_eloc.so — the C ABI declared in include/naqs_eloc.h.

There is no CPU fallback: if the shared library cannot be loaded (and cannot be built because
nvcc is absent) importing a compute entry raises, and every compute call on a machine without a
CUDA device fails with NaqsError (NAQS_ERR_CUDA).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _build

NAQS_C128, NAQS_C64 = 0, 1
LOOKUP_AUTO, LOOKUP_DENSE, LOOKUP_HASH = 0, 1, 2
LOOKUP_ASSUME_UNIQUE = 0x100
LOOKUP_DUPLICATES_EQUAL = 0x200
_OK, _ERR_ARG, _ERR_DTYPE, _ERR_CUDA, _ERR_ALLOC, _ERR_STATE, _ERR_INDEX = range(7)

# every symbol include/naqs_eloc.h declares: (restype, argtypes)
_p, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
SIGNATURES = {
    "naqs_last_error": (C.c_char_p, []),
    "naqs_abi_version": (_i, []),
    "naqs_device_count": (_i, []),
    "naqs_launch_count": (_i64, []),
    "naqs_table_create": (_i, [C.POINTER(_p), _p, _p, _p, _i64, _i, _i, _i, _i, _i]),
    "naqs_table_destroy": (_i, [_p]),
    "naqs_table_info": (_i, [_p, _p]),
    "naqs_lookup_build": (_i, [_p, _p, _p, _i, _i64, _i, _p]),
    "naqs_eloc": (_i, [_p, _p, _p, _i, _i64, _p, _p]),
    "naqs_table_set_algo": (_i, [_p, _i]),
    "naqs_table_set_precision": (_i, [_p, _i]),
    "naqs_apply_h": (_i, [_p, _p, _i64, _p, _p]),
    "naqs_dense32_scatter": (_i, [_p, _i64, _p, _p, _i64, _p]),
    "naqs_lookup_attach_dense32": (_i, [_p, _p, _i64]),
    "naqs_eloc_host": (_i, [_p, _p, _i, _p, _i, _i64, _p, _p, _i64, _i, _p, _i]),
    "naqs_eloc_host_begin": (_i, [_p, _p, _i, _p, _i, _i64, _p, _p, _i64, _i, _p, _i]),
    "naqs_eloc_host_end": (_i, [_p]),
    "naqs_table_check": (_i, [_p, _p]),
    "naqs_comm_unique_id": (_i, [_p]),
    "naqs_comm_init": (_i, [C.POINTER(_p), _p, _i, _i, _i]),
    "naqs_comm_from_nccl": (_i, [C.POINTER(_p), _p, _i, _i, _i]),
    "naqs_comm_destroy": (_i, [_p]),
    "naqs_comm_info": (_i, [_p, C.POINTER(_i), C.POINTER(_i)]),
    "naqs_table_exchange": (_i, [_p, _p, _p, _p, _i, _i64, _i64, _i, _p]),
    "naqs_stats_allreduce": (_i, [_p, _p, _p]),
    "naqs_rows_count": (_i, [_p, _p, _i64, _p, _p]),
    "naqs_exclusive_scan": (_i, [_p, _p, _i64, _p, _p]),
    "naqs_rows_fill": (_i, [_p, _p, _i64, _p, _p, _p, _p, _p]),
    "naqs_hij_dense": (_i, [_p, _p, _i64, _p, _p]),
    "naqs_unique_keys": (_i, [_p, _p, _i64, _p, _p, _p]),
    "naqs_popcount_parity": (_i, [_p, _i, _i64, _p, _p]),
    "naqs_get_hij": (_i, [_i64, _i64, _i64, _i64, _p, _p, _p, _p, _i, _p, _p]),
    "naqs_sparse_dense_mv": (_i, [_p, _i, _p, _p, _i, _i64, _p, _p, _p]),
    "naqs_sparse_sparse_mv": (_i, [_p, _i, _p, _p, _i, _p, _p, _i64, _p, _p]),
    "naqs_make_basis_idxs": (_i, [_i, _p, _p]),
    "naqs_state2idx": (_i, [_p, _i64, _i, _i, _p, _p]),
    "naqs_restricted_index": (_i, [_p, _p, _i64, _p, _p]),
    "naqs_eloc_stats": (_i, [_p, _p, _p, _i64, _p, _p]),
    "naqs_loss_terms": (_i, [_p, _p, _i64, _p, _p, _p, _p, _p, _p]),
}


class NaqsError(RuntimeError):
    """A libnaqs_eloc call failed (CUDA error, no device, out of memory, call-order violation)."""


_lib = None


def library_path():
    return _build.LIB


def load():
    """Load (building first if the sources are newer and nvcc exists) libnaqs_eloc.so."""
    global _lib
    if _lib is not None:
        return _lib
    if _build.needs_build():
        try:
            _build.build_library()
        except Exception as e:  # noqa: BLE001
            if not os.path.exists(_build.LIB):
                raise ImportError(
                    "libnaqs_eloc.so is missing and could not be built; the B200 E_loc path has no CPU "
                    f"fallback. Run `python -c 'import __graft_entry__ as g; g.build()'`. ({e})") from e
    lib = C.CDLL(_build.LIB)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => the header and the library disagree
        fn.restype, fn.argtypes = res, args
    if lib.naqs_abi_version() != 1:
        raise ImportError("libnaqs_eloc.so has an unexpected ABI version")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc == _OK:
        return
    msg = load().naqs_last_error().decode("utf-8", "replace")
    if rc == _ERR_DTYPE:
        raise TypeError(msg)
    if rc == _ERR_ARG:
        raise ValueError(msg)
    if rc == _ERR_INDEX:
        raise IndexError(msg)
    raise NaqsError(f"{what or 'libnaqs_eloc'}: {msg} (status {rc})")


def require_cuda(device=None):
    """-> torch.device; raises NaqsError when no CUDA device is usable (no CPU fallback)."""
    if load().naqs_device_count() <= 0 or not torch.cuda.is_available():
        raise NaqsError("no CUDA device available: the B200 E_loc path has no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise NaqsError(f"device {device} is not a CUDA device: the B200 E_loc path has no CPU fallback")
    return torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())


def ptr(t):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    return C.c_void_p(t.data_ptr())


def stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count():
    return int(load().naqs_launch_count())


# ---------------------------------------------------------------------------------- conversions
def n_words(n_qubits):
    return 1 if n_qubits <= 63 else 2


def keys_to_numpy(x, words=1):
    """ints / integer arrays / torch int tensors / [n, words] uint64 -> contiguous uint64 [n, words] (host)."""
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()
    a = np.asarray(x)
    if a.dtype == object:
        vals = [int(v) for v in a.reshape(-1)]
        out = np.zeros((len(vals), words), np.uint64)
        for w in range(words):
            out[:, w] = np.array([(v >> (64 * w)) & 0xFFFFFFFFFFFFFFFF for v in vals], dtype=np.uint64)
        return out
    if a.dtype == np.uint64:
        if a.ndim == 2 and a.shape[1] == words:
            return np.ascontiguousarray(a)
        flat = a.reshape(-1)
    else:
        if not np.issubdtype(a.dtype, np.integer):
            raise TypeError(f"state indices must be integers, got {a.dtype}")
        flat = a.reshape(-1).astype(np.int64).view(np.uint64)
    if words == 1:
        return np.ascontiguousarray(flat.reshape(-1, 1))
    out = np.zeros((flat.size, words), np.uint64)
    out[:, 0] = flat
    return out


def keys_to_device(x, words, device):
    """-> int64 tensor [n, words] on `device` holding the uint64 bit patterns."""
    if torch.is_tensor(x) and x.is_cuda:
        if x.dtype == torch.int64 and x.dim() == 2 and x.shape[1] == words:
            return x.contiguous()
        if words == 1 and x.dtype in (torch.int16, torch.int32, torch.int64, torch.uint8, torch.int8):
            return x.reshape(-1, 1).to(torch.int64).contiguous()
        x = x.cpu()
    k = keys_to_numpy(x, words)
    return torch.from_numpy(k.view(np.int64)).to(device, non_blocking=False)


def psi_to_device(psi, device):
    """numpy complex / torch complex / torch real [n, 2] -> (tensor viewed as real [n, 2], dtype code).
    float32 / complex64 stays single precision on the wire (promoted exactly in the kernel,
    src_cpp/sparse_math.pyx:33-37); everything else is complex128."""
    if isinstance(psi, np.ndarray):
        if np.iscomplexobj(psi):
            psi = psi.astype(np.complex64 if psi.dtype == np.complex64 else np.complex128, copy=False)
        else:
            psi = psi.astype(np.complex128)
        t = torch.from_numpy(np.ascontiguousarray(psi).reshape(-1))
    else:
        t = psi.detach()
        if not t.is_complex():
            if t.dim() >= 1 and t.shape[-1] == 2 and t.dtype in (torch.float32, torch.float64):
                t = torch.view_as_complex(t.contiguous().reshape(-1, 2))
            else:
                t = t.reshape(-1).to(torch.float64).to(torch.complex128)
        t = t.reshape(-1)
        if t.dtype not in (torch.complex64, torch.complex128):
            t = t.to(torch.complex128)
    t = t.to(device).contiguous()
    code = NAQS_C64 if t.dtype == torch.complex64 else NAQS_C128
    return torch.view_as_real(t), code


def complex_from_pairs(t):
    """[n, 2] float64 tensor -> complex128 numpy array (host)."""
    return torch.view_as_complex(t.contiguous()).cpu().numpy()
