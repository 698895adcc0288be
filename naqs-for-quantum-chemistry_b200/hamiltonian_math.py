"""Drop-in for the reference's compiled module `src.utils.hamiltonian_math`
(src_cpp/hamiltonian_math.pyx): same function names, arguments, return types and errors,
computed by the sm_100a kernels behind the C ABI (naqs_get_hij / naqs_popcount_parity).

numpy in -> numpy out, like the Cython originals; the arrays cross to the device and back.
These Level-0 twins exist for function-for-function parity; the fast path is the fused
`DeviceTermTable.local_energy` (the Level-0 seam forces M x Kyz / M x Kxy host arrays).
"""
import numpy as np
import torch

from . import _lib

_PP_DTYPES = (np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64)


def popcount_parity(arr):
    """1 - 2*(popcount(x) & 1) elementwise -> int8, always 2-D (src_cpp/hamiltonian_math.pyx:455-484)."""
    arr = np.asarray(arr)
    if len(arr.shape) == 1:
        arr = arr.reshape(-1, 1)
    if arr.dtype not in _PP_DTYPES:
        raise TypeError(f"Unsupported array dtype for popcount_parity(...): {arr.dtype}.")
    dev = _lib.require_cuda()
    a = np.ascontiguousarray(arr)
    # torch has no uint16/32/64 arithmetic, but only the bit pattern matters: ship raw bytes
    d_in = torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
    d_out = torch.empty(a.size, dtype=torch.int8, device=dev)
    _lib.check(_lib.load().naqs_popcount_parity(_lib.ptr(d_in), a.dtype.itemsize, a.size, _lib.ptr(d_out), _lib.stream_ptr(dev)),
               "naqs_popcount_parity")
    return d_out.cpu().numpy().reshape(arr.shape)


def get_Hij_cy(state_i_idx, _unique_XY_sites_idx, _unique2all_XY_sites_idx,
               P_k_by_unique_YZ_sites, _unique2all_YZ_sites_idx,
               couplings):
    """H_ij[m*Kxy + unique2all_XY[k]] += P[m, unique2all_YZ[k]] * couplings[k], k ascending
    (src_cpp/hamiltonian_math.pyx:200-288).  Returns a 1-D array of couplings.dtype.
    float32 / float64 couplings; long double has no device type -> TypeError."""
    M = len(state_i_idx)
    Kxy = len(_unique_XY_sites_idx)
    K = len(_unique2all_XY_sites_idx)
    couplings = np.asarray(couplings).squeeze()
    if couplings.dtype not in (np.float32, np.float64):
        raise TypeError(f"get_Hij_cy on the device supports float32/float64 couplings, got {couplings.dtype}.")
    dev = _lib.require_cuda()
    P = np.ascontiguousarray(np.asarray(P_k_by_unique_YZ_sites).astype(np.int8, copy=False))
    Kyz = P.shape[1] if P.ndim == 2 else 1
    d_u2a_xy = torch.from_numpy(np.ascontiguousarray(np.asarray(_unique2all_XY_sites_idx).astype(np.int64))).to(dev)
    d_u2a_yz = torch.from_numpy(np.ascontiguousarray(np.asarray(_unique2all_YZ_sites_idx).astype(np.int64))).to(dev)
    d_P = torch.from_numpy(P.reshape(-1)).to(dev)
    d_c = torch.from_numpy(np.ascontiguousarray(couplings.reshape(-1))).to(dev)
    d_H = torch.zeros(M * Kxy, dtype=d_c.dtype, device=dev)
    _lib.check(_lib.load().naqs_get_hij(M, Kxy, Kyz, K, _lib.ptr(d_u2a_xy), _lib.ptr(d_P), _lib.ptr(d_u2a_yz), _lib.ptr(d_c),
                                        couplings.dtype.itemsize, _lib.ptr(d_H), _lib.stream_ptr(dev)), "naqs_get_hij")
    return d_H.cpu().numpy()
