// level0.cu — function-for-function device twins of the reference's Cython entry points
// (src_cpp/hamiltonian_math.pyx, sparse_math.pyx, hilbert_math.pyx), the state2idx packing of
// src/utils/hilbert.py:573-581 and the E_loc statistics of src/optimizer/energy.py:328,372-375.
// C ABI documented in include/naqs_eloc.h.
#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace naqs;

namespace {

// ---------------------------------------------------------------- popcount_parity
template <typename T>
__global__ void popcount_parity_kernel(const T* __restrict__ in, int64_t n, int8_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p;
    if constexpr (sizeof(T) == 8) p = __popcll((unsigned long long)in[i]);
    else p = __popc((unsigned)in[i]);
    out[i] = (int8_t)(1 - 2 * (p & 1));
}

// ---------------------------------------------------------------- get_Hij_cy
// thread per (state m, XY group g): walks the group's terms in ascending k (host-built CSR of terms)
template <typename F>
__global__ void get_hij_kernel(int64_t M, int64_t Kxy, int64_t Kyz, const int64_t* __restrict__ grp_ptr,
                               const int64_t* __restrict__ grp_terms, const int8_t* __restrict__ parity,
                               const int64_t* __restrict__ u2a_yz, const F* __restrict__ coeff, F* __restrict__ hij) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * Kxy) return;
    const int64_t m = i / Kxy, g = i - m * Kxy;
    F acc = (F)0;
    for (int64_t e = grp_ptr[g]; e < grp_ptr[g + 1]; ++e) {
        const int64_t k = grp_terms[e];
        const F term = (F)parity[m * Kyz + u2a_yz[k]] * coeff[k];  // exact for +-1
        if constexpr (sizeof(F) == 8) acc = __dadd_rn(acc, term);
        else acc = __fadd_rn(acc, term);
    }
    hij[i] = acc;
}

// ---------------------------------------------------------------- sparse_dense_mv / sparse_sparse_mv
template <typename F> struct cplx_of;
template <> struct cplx_of<double> { using type = double2; };
template <> struct cplx_of<float> { using type = float2; };

template <typename F>
__device__ __forceinline__ F mul_rn(F a, F b) {
    if constexpr (sizeof(F) == 8) return __dmul_rn(a, b);
    else return __fmul_rn(a, b);
}
template <typename F>
__device__ __forceinline__ F add_rn(F a, F b) {
    if constexpr (sizeof(F) == 8) return __dadd_rn(a, b);
    else return __fadd_rn(a, b);
}

template <typename F, typename I>
__global__ void spmv_kernel(const F* __restrict__ data, const I* __restrict__ indices, const I* __restrict__ indptr,
                            int64_t n_rows, const typename cplx_of<F>::type* __restrict__ v,
                            typename cplx_of<F>::type* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    F re = (F)0, im = (F)0;
    for (int64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
        const auto x = v[indices[e]];
        const F a = data[e];
        re = add_rn<F>(re, mul_rn<F>(a, x.x));
        im = add_rn<F>(im, mul_rn<F>(a, x.y));
    }
    out[r].x = re; out[r].y = im;
}

template <typename F, typename I>
__global__ void spsmv_kernel(const F* __restrict__ data, const I* __restrict__ indices, const I* __restrict__ indptr,
                             const typename cplx_of<F>::type* __restrict__ v, const I* __restrict__ v_idxs, int64_t n_v,
                             typename cplx_of<F>::type* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_v) return;
    const int64_t row = (int64_t)v_idxs[k];
    F re = (F)0, im = (F)0;
    for (int64_t e = indptr[row]; e < indptr[row + 1]; ++e) {
        const I col = indices[e];
        int64_t lo = 0, hi = n_v;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (v_idxs[mid] < col) lo = mid + 1; else hi = mid;
        }
        if (lo < n_v && v_idxs[lo] == col) {
            const auto x = v[lo];
            const F a = data[e];
            re = add_rn<F>(re, mul_rn<F>(a, x.x));
            im = add_rn<F>(im, mul_rn<F>(a, x.y));
        }
    }
    out[k].x = re; out[k].y = im;
}

// ---------------------------------------------------------------- make_basis_idxs_cy
__global__ void make_basis_kernel(int n_qubits, int64_t total, int32_t* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int64_t i = e / n_qubits;
    const int j = (int)(e - i * n_qubits);
    out[e] = (int32_t)(i & (1ll << j));
}

// ---------------------------------------------------------------- state2idx
// one warp packs 32-qubit slices of a state row with a ballot: lane q reads byte [m][base+q] (coalesced)
__global__ void state2idx_kernel(const int8_t* __restrict__ states, int64_t M, int n_qubits, int words, uint64_t* __restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp; m < M; m += n_warps) {
        for (int w = 0; w < words; ++w) {
            unsigned lo = 0, hi = 0;
            for (int half = 0; half < 2; ++half) {
                const int q = w * 64 + half * 32 + lane;
                const bool occ = q < n_qubits && states[m * n_qubits + q] > 0;
                const unsigned b = __ballot_sync(0xffffffffu, occ);
                if (half == 0) lo = b; else hi = b;
            }
            if (lane == 0) keys[m * words + w] = (unsigned long long)lo | ((unsigned long long)hi << 32);
        }
    }
}

// ---------------------------------------------------------------- E_loc statistics (deterministic two-stage)
constexpr int kStatThreads = 256;

__device__ __forceinline__ void block_sum5(double (&v)[5], double* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_down_sync(0xffffffffu, v[c], o);
        if (lane == 0) s_red[warp * 5 + c] = v[c];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0.0;
        for (int w = 0; w < kStatThreads / 32; ++w) s += s_red[w * 5 + threadIdx.x];
        s_red[threadIdx.x] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kStatThreads) stats_partial_kernel(const double2* __restrict__ eloc, const double* __restrict__ w,
                                                                     int64_t n, double* __restrict__ partial) {
    __shared__ double s_red[(kStatThreads / 32) * 5];
    double v[5] = {0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * kStatThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kStatThreads) {
        const double2 e = eloc[i];
        const double wi = w ? w[i] : 1.0;
        v[0] += wi; v[1] += wi * e.x; v[2] += wi * e.y; v[3] += wi * e.x * e.x; v[4] += 1.0;
    }
    block_sum5(v, s_red);
    if (threadIdx.x < 5) partial[blockIdx.x * 5 + threadIdx.x] = s_red[threadIdx.x];
}

__global__ void __launch_bounds__(kStatThreads) stats_final_kernel(const double* __restrict__ partial, int n_part, double* __restrict__ out5) {
    __shared__ double s_red[(kStatThreads / 32) * 5];
    double v[5] = {0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < n_part; i += kStatThreads)
#pragma unroll
        for (int c = 0; c < 5; ++c) v[c] += partial[i * 5 + c];
    block_sum5(v, s_red);
    if (threadIdx.x < 5) out5[threadIdx.x] = s_red[threadIdx.x];
}

inline unsigned blocks_for(int64_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

template <typename F>
int run_spmv(const void* data, const void* indices, const void* indptr, int idx_itemsize, int64_t n_rows, const void* v,
             void* out, cudaStream_t st) {
    using C = typename cplx_of<F>::type;
    if (idx_itemsize == 4)
        spmv_kernel<F, int32_t><<<blocks_for(n_rows), 256, 0, st>>>((const F*)data, (const int32_t*)indices, (const int32_t*)indptr, n_rows, (const C*)v, (C*)out);
    else
        spmv_kernel<F, int64_t><<<blocks_for(n_rows), 256, 0, st>>>((const F*)data, (const int64_t*)indices, (const int64_t*)indptr, n_rows, (const C*)v, (C*)out);
    NAQS_LAUNCHED();
    return NAQS_OK;
}

template <typename F>
int run_spsmv(const void* data, const void* indices, const void* indptr, int idx_itemsize, const void* v, const void* v_idxs,
              int64_t n_v, void* out, cudaStream_t st) {
    using C = typename cplx_of<F>::type;
    if (idx_itemsize == 4)
        spsmv_kernel<F, int32_t><<<blocks_for(n_v), 256, 0, st>>>((const F*)data, (const int32_t*)indices, (const int32_t*)indptr, (const C*)v, (const int32_t*)v_idxs, n_v, (C*)out);
    else
        spsmv_kernel<F, int64_t><<<blocks_for(n_v), 256, 0, st>>>((const F*)data, (const int64_t*)indices, (const int64_t*)indptr, (const C*)v, (const int64_t*)v_idxs, n_v, (C*)out);
    NAQS_LAUNCHED();
    return NAQS_OK;
}

}  // namespace

// Loss statistics of one VMC step from E_loc and the (reduced) five sums (src/optimizer/energy.py:316-329, 367-375), fp64:
//   mean = (sum w E) / (sum w)                         the weighted complex mean subtracted at energy.py:328
//   e_loc_corr_i = E_i - mean
//   grad_w_i = 2 w_i / (sum w) * conj(e_loc_corr_i)     so that  exp_op = sum_i Re(log_psi_i * 2 w_i e_loc_corr_i)
//                                                                   = sum_i (Re log_psi_i * grad_w_i.x + Im log_psi_i * grad_w_i.y):
//                                                       autograd only needs this detached weight vector (energy.py:329)
//   energy = Re mean,  variance = sum w (Re E - energy)^2 / sum w = (sum w (Re E)^2) / (sum w) - energy^2   (energy.py:372-375)
// Outputs are float32 pairs like the tensors the reference's loss sees (complex.py:139-140); any of them may be NULL.
__global__ void loss_terms_kernel(const double2* __restrict__ eloc, const double* __restrict__ w, int64_t n, const double* __restrict__ sums5,
                                  float2* __restrict__ eloc32, float2* __restrict__ corr32, float2* __restrict__ gradw32, double* __restrict__ energy_var) {
    const double sw = sums5[0];
    const double mean_re = sums5[1] / sw, mean_im = sums5[2] / sw;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && energy_var) {
        energy_var[0] = mean_re;
        energy_var[1] = mean_im;
        energy_var[2] = sums5[3] / sw - mean_re * mean_re;
    }
    if (i >= n) return;
    const double2 e = eloc[i];
    const double wi = (w ? w[i] : 1.0) / sw;
    const double cr = e.x - mean_re, ci = e.y - mean_im;
    if (eloc32) eloc32[i] = make_float2((float)e.x, (float)e.y);
    if (corr32) corr32[i] = make_float2((float)cr, (float)ci);
    if (gradw32) gradw32[i] = make_float2((float)(2.0 * wi * cr), (float)(-2.0 * wi * ci));
}

extern "C" {

int naqs_popcount_parity(const void* d_in, int itemsize, int64_t n, int8_t* d_out, void* stream) {
    NAQS_REQUIRE(itemsize == 1 || itemsize == 2 || itemsize == 4 || itemsize == 8, NAQS_ERR_DTYPE,
                 "Unsupported array dtype for popcount_parity(...): itemsize must be 1, 2, 4 or 8.");
    NAQS_REQUIRE(n >= 0 && (n == 0 || (d_in && d_out)), NAQS_ERR_ARG, "naqs_popcount_parity: NULL buffers");
    if (n == 0) return NAQS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (itemsize) {
        case 1: popcount_parity_kernel<uint8_t><<<blocks_for(n), 256, 0, st>>>((const uint8_t*)d_in, n, d_out); break;
        case 2: popcount_parity_kernel<uint16_t><<<blocks_for(n), 256, 0, st>>>((const uint16_t*)d_in, n, d_out); break;
        case 4: popcount_parity_kernel<uint32_t><<<blocks_for(n), 256, 0, st>>>((const uint32_t*)d_in, n, d_out); break;
        default: popcount_parity_kernel<uint64_t><<<blocks_for(n), 256, 0, st>>>((const uint64_t*)d_in, n, d_out); break;
    }
    NAQS_LAUNCHED();
    return NAQS_OK;
}

int naqs_get_hij(int64_t M, int64_t Kxy, int64_t Kyz, int64_t K, const int64_t* d_u2a_xy, const int8_t* d_parity,
                 const int64_t* d_u2a_yz, const void* d_coeff, int coeff_itemsize, void* d_hij, void* stream) {
    NAQS_REQUIRE(coeff_itemsize == 8 || coeff_itemsize == 4, NAQS_ERR_DTYPE,
                 "naqs_get_hij: couplings must be float64 or float32 (long double has no device type)");
    NAQS_REQUIRE(M >= 0 && Kxy >= 0 && K >= 0, NAQS_ERR_ARG, "naqs_get_hij: negative size");
    if (M * Kxy == 0) return NAQS_OK;
    NAQS_REQUIRE(d_hij && (K == 0 || (d_u2a_xy && d_parity && d_u2a_yz && d_coeff)), NAQS_ERR_ARG, "naqs_get_hij: NULL buffers");
    cudaStream_t st = (cudaStream_t)stream;
    // CSR of terms per XY group, k ascending (stable counting sort on the host: K is small, this is a parity shim)
    std::vector<int64_t> u2a(K), ptr(Kxy + 1, 0), terms(K);
    if (K) NAQS_CUDA(cudaMemcpyAsync(u2a.data(), d_u2a_xy, K * 8, cudaMemcpyDeviceToHost, st));
    NAQS_CUDA(cudaStreamSynchronize(st));
    for (int64_t k = 0; k < K; ++k) {
        NAQS_REQUIRE(u2a[k] >= 0 && u2a[k] < Kxy, NAQS_ERR_ARG, "naqs_get_hij: unique2all_XY index out of range");
        ptr[u2a[k] + 1]++;
    }
    for (int64_t g = 0; g < Kxy; ++g) ptr[g + 1] += ptr[g];
    {
        std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
        for (int64_t k = 0; k < K; ++k) terms[fill[u2a[k]]++] = k;
    }
    int64_t *d_ptr = nullptr, *d_terms = nullptr;
    NAQS_CUDA(cudaMallocAsync((void**)&d_ptr, (Kxy + 1) * 8, st));
    NAQS_CUDA(cudaMallocAsync((void**)&d_terms, std::max<int64_t>(K, 1) * 8, st));
    NAQS_CUDA(cudaMemcpyAsync(d_ptr, ptr.data(), (Kxy + 1) * 8, cudaMemcpyHostToDevice, st));
    if (K) NAQS_CUDA(cudaMemcpyAsync(d_terms, terms.data(), K * 8, cudaMemcpyHostToDevice, st));
    if (coeff_itemsize == 8)
        get_hij_kernel<double><<<blocks_for(M * Kxy), 256, 0, st>>>(M, Kxy, Kyz, d_ptr, d_terms, d_parity, d_u2a_yz, (const double*)d_coeff, (double*)d_hij);
    else
        get_hij_kernel<float><<<blocks_for(M * Kxy), 256, 0, st>>>(M, Kxy, Kyz, d_ptr, d_terms, d_parity, d_u2a_yz, (const float*)d_coeff, (float*)d_hij);
    NAQS_LAUNCHED();
    NAQS_CUDA(cudaStreamSynchronize(st));  // host vectors are about to go away
    cudaFreeAsync(d_ptr, st);
    cudaFreeAsync(d_terms, st);
    return NAQS_OK;
}

int naqs_sparse_dense_mv(const void* d_data, int data_itemsize, const void* d_indices, const void* d_indptr, int idx_itemsize,
                         int64_t n_rows, const void* d_v, void* d_out, void* stream) {
    NAQS_REQUIRE(data_itemsize == 8 || data_itemsize == 4, NAQS_ERR_DTYPE, "m must have dtype of np.float32 or np.float64.");
    NAQS_REQUIRE(idx_itemsize == 4 || idx_itemsize == 8, NAQS_ERR_DTYPE, "naqs_sparse_dense_mv: indices must be int32 or int64");
    NAQS_REQUIRE(n_rows >= 0 && (n_rows == 0 || (d_indptr && d_out)), NAQS_ERR_ARG, "naqs_sparse_dense_mv: NULL buffers");
    if (n_rows == 0) return NAQS_OK;
    return data_itemsize == 8 ? run_spmv<double>(d_data, d_indices, d_indptr, idx_itemsize, n_rows, d_v, d_out, (cudaStream_t)stream)
                              : run_spmv<float>(d_data, d_indices, d_indptr, idx_itemsize, n_rows, d_v, d_out, (cudaStream_t)stream);
}

int naqs_sparse_sparse_mv(const void* d_data, int data_itemsize, const void* d_indices, const void* d_indptr, int idx_itemsize,
                          const void* d_v, const void* d_v_idxs_sorted, int64_t n_v, void* d_out, void* stream) {
    NAQS_REQUIRE(data_itemsize == 8 || data_itemsize == 4, NAQS_ERR_DTYPE, "m must have dtype of np.float32 or np.float64.");
    NAQS_REQUIRE(idx_itemsize == 4 || idx_itemsize == 8, NAQS_ERR_DTYPE, "naqs_sparse_sparse_mv: indices must be int32 or int64");
    NAQS_REQUIRE(n_v >= 0 && (n_v == 0 || (d_indptr && d_v && d_v_idxs_sorted && d_out)), NAQS_ERR_ARG, "naqs_sparse_sparse_mv: NULL buffers");
    if (n_v == 0) return NAQS_OK;
    return data_itemsize == 8 ? run_spsmv<double>(d_data, d_indices, d_indptr, idx_itemsize, d_v, d_v_idxs_sorted, n_v, d_out, (cudaStream_t)stream)
                              : run_spsmv<float>(d_data, d_indices, d_indptr, idx_itemsize, d_v, d_v_idxs_sorted, n_v, d_out, (cudaStream_t)stream);
}

int naqs_make_basis_idxs(int n_qubits, int32_t* d_out, void* stream) {
    NAQS_REQUIRE(n_qubits >= 0 && n_qubits <= 31, NAQS_ERR_ARG, "naqs_make_basis_idxs: n_qubits must be in [0, 31] (int32 output)");
    NAQS_REQUIRE(d_out || n_qubits == 0, NAQS_ERR_ARG, "naqs_make_basis_idxs: NULL output");
    const int64_t total = (1ll << n_qubits) * n_qubits;
    if (total == 0) return NAQS_OK;
    make_basis_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(n_qubits, total, d_out);
    NAQS_LAUNCHED();
    return NAQS_OK;
}

int naqs_state2idx(const int8_t* d_states, int64_t M, int n_qubits, int words, uint64_t* d_keys, void* stream) {
    NAQS_REQUIRE(words == 1 || words == 2, NAQS_ERR_ARG, "naqs_state2idx: words must be 1 or 2");
    NAQS_REQUIRE(n_qubits >= 1 && n_qubits <= 64 * words, NAQS_ERR_ARG, "naqs_state2idx: n_qubits does not fit the key width");
    NAQS_REQUIRE(M >= 0 && (M == 0 || (d_states && d_keys)), NAQS_ERR_ARG, "naqs_state2idx: NULL buffers");
    if (M == 0) return NAQS_OK;
    const int64_t warps = std::min<int64_t>(M, 148 * 64);
    state2idx_kernel<<<blocks_for(warps * 32), 256, 0, (cudaStream_t)stream>>>(d_states, M, n_qubits, words, d_keys);
    NAQS_LAUNCHED();
    return NAQS_OK;
}

int naqs_loss_terms(const double* d_eloc, const double* d_w, int64_t n, const double* d_sums5, float* d_eloc_f32, float* d_eloc_corr_f32,
                    float* d_grad_w_f32, double* d_energy_var, void* stream) {
    NAQS_REQUIRE(n >= 0 && d_sums5 && (n == 0 || d_eloc), NAQS_ERR_ARG, "naqs_loss_terms: NULL buffers");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, (n + 255) / 256);
    loss_terms_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const double2*>(d_eloc), d_w, n, d_sums5, reinterpret_cast<float2*>(d_eloc_f32),
                                              reinterpret_cast<float2*>(d_eloc_corr_f32), reinterpret_cast<float2*>(d_grad_w_f32), d_energy_var);
    NAQS_LAUNCHED();
    return NAQS_OK;
}

int naqs_eloc_stats(naqs_table_t* t, const double* d_eloc, const double* d_w, int64_t n, double* d_out5, void* stream) {
    NAQS_REQUIRE(t && d_out5, NAQS_ERR_ARG, "naqs_eloc_stats: NULL argument");
    NAQS_REQUIRE(n >= 0 && (n == 0 || d_eloc), NAQS_ERR_ARG, "naqs_eloc_stats: NULL buffers");
    DeviceGuard guard(t->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int n_part = (int)std::max<int64_t>(1, std::min<int64_t>((n + kStatThreads - 1) / kStatThreads, 148 * 4));
    int rc = ensure_ws(t, (size_t)n_part * 5 * 8);
    if (rc) return rc;
    double* partial = static_cast<double*>(t->d_ws);
    stats_partial_kernel<<<n_part, kStatThreads, 0, st>>>(reinterpret_cast<const double2*>(d_eloc), d_w, n, partial);
    NAQS_LAUNCHED();
    stats_final_kernel<<<1, kStatThreads, 0, st>>>(partial, n_part, d_out5);
    NAQS_LAUNCHED();
    return NAQS_OK;
}

}  // extern "C"
