// sort_scan.cuh — hand-written device primitives for the coupled-set (K4) path:
//   * exclusive scan of int64 (reduce / scan-of-partials / down-sweep, 2048 items per CTA)
//   * stable LSD radix sort of 64/128-bit keys, 8-bit digits, only the bits a key can have
//     (ceil(n_qubits / 8) passes): per-CTA digit histogram -> global scan -> stable scatter with
//     warp-level __match_any_sync ranking
//   * adjacent-unique compaction
// Together they reproduce np.unique(np.concatenate(coupled_idxs)) of
// src/optimizer/hamiltonian.py:131 on the device.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace naqs {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* s_warp, long long& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = lane < (kScanThreads / 32) ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += n;
        }
        if (lane < (kScanThreads / 32)) s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const long long warp_off = warp == 0 ? 0 : s_warp[warp - 1];
    total = s_warp[kScanThreads / 32 - 1];
    const long long r = warp_off + incl - v;
    __syncthreads();
    return r;
}

// pass 1: per-tile sums
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ partial) {
    __shared__ long long s_warp[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    long long v = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int64_t idx = base + (int64_t)i * kScanThreads + threadIdx.x;
        if (idx < n) v += in[idx];
    }
    long long total;
    block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// pass 2: one CTA scans the partials in place (exclusive) and writes the grand total to *total_out
__global__ void __launch_bounds__(kScanThreads) scan_partials_kernel(int64_t* partial, int64_t n_part, int64_t* total_out) {
    __shared__ long long s_warp[kScanThreads / 32];
    long long carry = 0;
    for (int64_t base = 0; base < n_part; base += kScanThreads) {
        const int64_t idx = base + threadIdx.x;
        const long long v = idx < n_part ? partial[idx] : 0;
        long long total;
        const long long ex = block_exclusive_scan(v, s_warp, total);
        if (idx < n_part) partial[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// pass 3: down-sweep.  Each thread owns kScanItems CONSECUTIVE items so the scan is in element order.
__global__ void __launch_bounds__(kScanThreads) scan_downsweep_kernel(const int64_t* __restrict__ in, int64_t n,
                                                                      const int64_t* __restrict__ partial, int64_t* __restrict__ out) {
    __shared__ long long s_warp[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    long long item[kScanItems], sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        item[i] = (base + i < n) ? in[base + i] : 0;
        sum += item[i];
    }
    long long total;
    long long run = block_exclusive_scan(sum, s_warp, total) + partial[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = run;
        run += item[i];
    }
}

// out[0..n) = exclusive scan, out[n] = total.  `in` and `out` may not alias.
static int exclusive_scan_i64(naqs_table* t, const int64_t* d_in, int64_t n, int64_t* d_out, cudaStream_t stream,
                              size_t ws_offset = 0) {
    if (n == 0) {
        NAQS_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int64_t), stream));
        return NAQS_OK;
    }
    const int64_t n_part = (n + kScanTile - 1) / kScanTile;
    int rc = ensure_ws(t, ws_offset + (size_t)n_part * 8 + 256);
    if (rc) return rc;
    int64_t* partial = reinterpret_cast<int64_t*>(static_cast<char*>(t->d_ws) + ws_offset);
    scan_reduce_kernel<<<(unsigned)n_part, kScanThreads, 0, stream>>>(d_in, n, partial);
    NAQS_LAUNCHED();
    scan_partials_kernel<<<1, kScanThreads, 0, stream>>>(partial, n_part, d_out + n);
    NAQS_LAUNCHED();
    scan_downsweep_kernel<<<(unsigned)n_part, kScanThreads, 0, stream>>>(d_in, n, partial, d_out);
    NAQS_LAUNCHED();
    return NAQS_OK;
}

// ------------------------------------------------------------------------------------------ radix sort
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortChunks = 16;                               // 32-key chunks per warp
constexpr int kSortTile = kSortWarps * kSortChunks * 32;      // 4096 keys per CTA

template <int W>
__device__ __forceinline__ unsigned digit_of(const uint64_t* __restrict__ keys, int64_t i, int bit) {
    const uint64_t w = keys[i * W + (bit >> 6)];
    return (unsigned)(w >> (bit & 63)) & 0xffu;
}

// histogram of the current digit, digit-major output hist[d * n_blocks + block]
template <int W>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int bit,
                                                                  int64_t* __restrict__ hist, int n_blocks) {
    __shared__ unsigned s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += kSortThreads) {
        const int64_t idx = base + i;
        if (idx < n) atomicAdd(&s_hist[digit_of<W>(keys, idx, bit)], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * n_blocks + blockIdx.x] = s_hist[threadIdx.x];
}

// stable scatter: warp w of the CTA owns the contiguous sub-tile [w*512, (w+1)*512) of the tile
template <int W>
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint64_t* __restrict__ keys, int64_t n, int bit,
                                                                     const int64_t* __restrict__ offsets, int n_blocks,
                                                                     uint64_t* __restrict__ out) {
    __shared__ unsigned s_cnt[kSortWarps][256];   // per-warp digit counts, then running positions
    __shared__ long long s_base[256];             // global offset of (digit, this CTA)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
    s_base[threadIdx.x] = offsets[(int64_t)threadIdx.x * n_blocks + blockIdx.x];
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (kSortChunks * 32);
    // phase 1: per-warp counts
    for (int c = 0; c < kSortChunks; ++c) {
        const int64_t idx = wbase + c * 32 + lane;
        const bool ok = idx < n;
        const unsigned d = ok ? digit_of<W>(keys, idx, bit) : 0x100u + lane;  // inactive lanes never match
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (ok && (peers & ((1u << lane) - 1)) == 0) s_cnt[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over warps for each digit (thread d handles digit d)
    {
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const unsigned c = s_cnt[w][threadIdx.x];
            s_cnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase 2: ranked scatter in sub-tile order
    for (int c = 0; c < kSortChunks; ++c) {
        const int64_t idx = wbase + c * 32 + lane;
        const bool ok = idx < n;
        const unsigned d = ok ? digit_of<W>(keys, idx, bit) : 0x100u + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned rank = __popc(peers & ((1u << lane) - 1));
        unsigned pos = 0;
        if (ok) pos = s_cnt[warp][d] + rank;
        __syncwarp();
        if (ok && rank == 0) s_cnt[warp][d] += __popc(peers);
        __syncwarp();
        if (ok) {
            const int64_t dst = s_base[d] + pos;
#pragma unroll
            for (int w = 0; w < W; ++w) out[dst * W + w] = keys[idx * W + w];
        }
    }
}

template <int W>
__global__ void unique_flag_kernel(const uint64_t* __restrict__ keys, int64_t n, int64_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool first = i == 0;
    if (!first) {
#pragma unroll
        for (int w = 0; w < W; ++w) first |= keys[i * W + w] != keys[(i - 1) * W + w];
    }
    flags[i] = first ? 1 : 0;
}

template <int W>
__global__ void unique_compact_kernel(const uint64_t* __restrict__ keys, int64_t n, const int64_t* __restrict__ flags,
                                      const int64_t* __restrict__ pos, uint64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
#pragma unroll
    for (int w = 0; w < W; ++w) out[pos[i] * W + w] = keys[i * W + w];
}

template <int W>
static int sort_unique_impl(naqs_table* t, const uint64_t* d_keys, int64_t n, uint64_t* d_out, int64_t* h_n_unique,
                            cudaStream_t stream) {
    const int n_blocks = (int)((n + kSortTile - 1) / kSortTile);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t key_bytes = al((size_t)n * W * 8);
    const size_t hist_bytes = al((size_t)256 * n_blocks * 8 + 8);
    const size_t flag_bytes = al((size_t)n * 8);
    // layout: bufA | bufB | hist | hist_scanned | flags | pos | scan partials
    const size_t o_a = 0, o_b = key_bytes, o_h = 2 * key_bytes, o_hs = o_h + hist_bytes, o_f = o_hs + hist_bytes,
                 o_p = o_f + flag_bytes, o_sc = o_p + flag_bytes + 256;
    const size_t scan_part = al((size_t)(std::max<int64_t>(n, (int64_t)256 * n_blocks) / kScanTile + 2) * 8);
    int rc = ensure_ws(t, o_sc + scan_part + 256);
    if (rc) return rc;
    char* ws = static_cast<char*>(t->d_ws);
    uint64_t* buf[2] = {reinterpret_cast<uint64_t*>(ws + o_a), reinterpret_cast<uint64_t*>(ws + o_b)};
    int64_t* hist = reinterpret_cast<int64_t*>(ws + o_h);
    int64_t* hist_s = reinterpret_cast<int64_t*>(ws + o_hs);
    int64_t* flags = reinterpret_cast<int64_t*>(ws + o_f);
    int64_t* pos = reinterpret_cast<int64_t*>(ws + o_p);

    const uint64_t* src = d_keys;
    int which = 0;
    const int n_bits = t->n_qubits;
    for (int bit = 0; bit < n_bits; bit += 8) {
        // a digit must not straddle the 64-bit word boundary: 64 % 8 == 0, so it never does
        radix_hist_kernel<W><<<n_blocks, kSortThreads, 0, stream>>>(src, n, bit, hist, n_blocks);
        NAQS_LAUNCHED();
        rc = exclusive_scan_i64(t, hist, (int64_t)256 * n_blocks, hist_s, stream, o_sc);
        if (rc) return rc;
        radix_scatter_kernel<W><<<n_blocks, kSortThreads, 0, stream>>>(src, n, bit, hist_s, n_blocks, buf[which]);
        NAQS_LAUNCHED();
        src = buf[which];
        which ^= 1;
    }
    const unsigned blocks = (unsigned)((n + 255) / 256);
    unique_flag_kernel<W><<<blocks, 256, 0, stream>>>(src, n, flags);
    NAQS_LAUNCHED();
    rc = exclusive_scan_i64(t, flags, n, pos, stream, o_sc);  // pos has room for n+1 (flag_bytes + 256)
    if (rc) return rc;
    unique_compact_kernel<W><<<blocks, 256, 0, stream>>>(src, n, flags, pos, d_out);
    NAQS_LAUNCHED();
    NAQS_CUDA(cudaMemcpyAsync(h_n_unique, pos + n, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    NAQS_CUDA(cudaStreamSynchronize(stream));
    return NAQS_OK;
}

static int sort_unique_keys(naqs_table* t, const uint64_t* d_keys, int64_t n, uint64_t* d_out, int64_t* h_n_unique,
                            cudaStream_t stream) {
    return t->words == 1 ? sort_unique_impl<1>(t, d_keys, n, d_out, h_n_unique, stream)
                         : sort_unique_impl<2>(t, d_keys, n, d_out, h_n_unique, stream);
}

}  // namespace naqs
