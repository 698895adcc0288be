// rows.cu — stored Hamiltonian rows (CSR / coupled-set mode), dense H_ij, restricted-index ranker,
// exclusive scan and sorted-unique of keys.  C ABI documented in include/naqs_eloc.h.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "eloc_kernels.cuh"
#include "sort_scan.cuh"

using namespace naqs;

namespace {

constexpr int kRowThreads = 128;

__global__ void rows_sum_chunk_counts_kernel(const int32_t* __restrict__ chunk_counts, int n_chunks, int64_t M, int64_t* __restrict__ counts) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    int64_t n = 0;
    for (int c = 0; c < n_chunks; ++c) n += chunk_counts[(int64_t)c * M + m];
    counts[m] = n;
}

int ensure_binom(naqs_table* t) {
    if (t->d_binom) return NAQS_OK;
    std::vector<long long> b(65 * 66, 0);
    for (int n = 0; n <= 64; ++n)
        for (int k = 0; k < 66; ++k) {
            long long v;
            if (k == 0) v = 1;
            else if (n == 0) v = 0;
            else {
                const long long a = b[(n - 1) * 66 + k - 1], c = b[(n - 1) * 66 + k];
                v = (a > (1ll << 62) || c > (1ll << 62)) ? (1ll << 62) : a + c;  // saturate (unreachable for int64 sectors)
            }
            b[n * 66 + k] = v;
        }
    NAQS_CUDA(cudaMalloc((void**)&t->d_binom, b.size() * sizeof(long long)));
    NAQS_CUDA(cudaMemcpy(t->d_binom, b.data(), b.size() * sizeof(long long), cudaMemcpyHostToDevice));
    // per-spin rank tables: rank of an occupation pattern among the C(n, k) combinations in itertools order (hilbert.py:446-469)
    if (t->sector.enabled && t->nw32 == 1) {
        const int n_even = (t->n_qubits + 1) / 2, n_odd = t->n_qubits / 2;
        std::vector<int32_t> tab(((size_t)1 << n_even) + ((size_t)1 << n_odd), -1);
        auto fill = [&](int32_t* dst, int n, int k) {
            for (uint32_t v = 0; v < (1u << n); ++v) {
                if (__builtin_popcount(v) != k) continue;
                long long r = 0;
                int seen = 0;
                for (int p = 0; p < n && seen < k; ++p) {
                    if ((v >> p) & 1u) ++seen;
                    else r += b[(size_t)(n - 1 - p) * 66 + (size_t)(k - 1 - seen)];
                }
                dst[v] = (int32_t)r;
            }
        };
        if (t->n_alpha <= n_even && t->n_beta <= n_odd && b[(size_t)n_even * 66 + t->n_alpha] < (1ll << 31) && b[(size_t)n_odd * 66 + t->n_beta] < (1ll << 31)) {
            fill(tab.data(), n_even, t->n_alpha);
            fill(tab.data() + ((size_t)1 << n_even), n_odd, t->n_beta);
            NAQS_CUDA(cudaMalloc((void**)&t->d_rank_tab, tab.size() * sizeof(int32_t)));
            NAQS_CUDA(cudaMemcpy(t->d_rank_tab, tab.data(), tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            t->rank_n_b = b[(size_t)n_odd * 66 + t->n_beta];
        }
    }
    return NAQS_OK;
}

RankView rank_view(const naqs_table* t) {
    const int n_even = (t->n_qubits + 1) / 2;
    return RankView{t->d_binom, t->d_rank_tab, t->d_rank_tab ? t->d_rank_tab + ((size_t)1 << n_even) : nullptr, t->rank_n_b};
}

// Number of table chunks (grid.y) for a batch of M rows: enough threads for about one full wave of resident threads, a power of
// two so that base chunks merge evenly; 1 (the plain tile list) for large batches.
int rows_chunks_for(const naqs_table* t, int64_t M) {
    if (t->n_row_chunks == 0) return 1;
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, t->device);
    const int64_t want_threads = (int64_t)sm_count * 2048;
    int c = 1;
    while (c < t->n_row_chunks && (int64_t)c * M < want_threads) c *= 2;
    return c;
}

template <int NW, int MODE>
int launch_rows(naqs_table* t, const uint64_t* d_states, int64_t M, int n_chunks, int64_t* d_counts, int32_t* d_chunk_counts,
                const int64_t* d_indptr, uint64_t* d_col_keys, int64_t* d_col_ridx, double* d_vals, cudaStream_t stream) {
    ChunkBounds cb;
    const Tile* tiles = t->d_tiles;
    int cap = t->tile_cap;
    if (n_chunks > 1) {
        const int merge = t->n_row_chunks / n_chunks;
        for (int c = 0; c <= kMaxChunks; ++c) cb.lo[c] = t->row_chunk_lo[std::min(c * merge, t->n_row_chunks)];
        tiles = t->d_row_tiles;
        cap = t->row_tile_cap;
    } else {
        cb.lo[0] = 0;
        for (int c = 1; c <= kMaxChunks; ++c) cb.lo[c] = t->n_tiles;
    }
    const size_t smem = tile_smem_bytes<NW>(cap);
    auto kern = rows_kernel<NW, MODE, kRowThreads>;
    NAQS_SMEM_ATTR(kern, smem, t->device);
    const int64_t blocks = (M + kRowThreads - 1) / kRowThreads;
    kern<<<dim3((unsigned)blocks, (unsigned)n_chunks), kRowThreads, smem, stream>>>(t->view(), tiles, cb, n_chunks, cap, t->sector, d_states,
                                                                                  M, t->words, rank_view(t), d_counts, d_chunk_counts, d_indptr,
                                                                                  d_col_keys, d_col_ridx, d_vals);
    NAQS_LAUNCHED();
    return NAQS_OK;
}

template <int MODE>
int dispatch_rows(naqs_table* t, const uint64_t* d_states, int64_t M, int n_chunks, int64_t* d_counts, int32_t* d_chunk_counts,
                  const int64_t* d_indptr, uint64_t* d_col_keys, int64_t* d_col_ridx, double* d_vals, cudaStream_t stream) {
    switch (t->nw32) {
        case 1: return launch_rows<1, MODE>(t, d_states, M, n_chunks, d_counts, d_chunk_counts, d_indptr, d_col_keys, d_col_ridx, d_vals, stream);
        case 2: return launch_rows<2, MODE>(t, d_states, M, n_chunks, d_counts, d_chunk_counts, d_indptr, d_col_keys, d_col_ridx, d_vals, stream);
        default: return launch_rows<4, MODE>(t, d_states, M, n_chunks, d_counts, d_chunk_counts, d_indptr, d_col_keys, d_col_ridx, d_vals, stream);
    }
}

// per-(chunk, row) counts of a chunked launch: [n_chunks][M] int32 in a buffer of their own (the caller's scan between
// naqs_rows_count and naqs_rows_fill uses the generic workspace)
int chunk_counts_buffer(naqs_table* t, int n_chunks, int64_t M, int32_t** out) {
    *out = nullptr;
    if (n_chunks <= 1) return NAQS_OK;
    const size_t bytes = (size_t)n_chunks * (size_t)M * sizeof(int32_t);
    if (t->row_cc_bytes < bytes) {
        cudaFree(t->d_row_cc); t->d_row_cc = nullptr; t->row_cc_bytes = 0;
        NAQS_CUDA(cudaMalloc((void**)&t->d_row_cc, bytes));
        t->row_cc_bytes = bytes;
    }
    *out = t->d_row_cc;
    return NAQS_OK;
}

template <int NW>
__global__ void restricted_index_kernel(Sector sec, const uint64_t* __restrict__ keys, int64_t n,
                                        RankView rank, int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t j[NW];
    load_key<NW>(keys, i, j);
    out[i] = restricted_index<NW>(j, sec, rank);
}

}  // namespace

extern "C" {

int naqs_rows_count(naqs_table_t* t, const uint64_t* d_states, int64_t M, int64_t* d_counts, void* stream) {
    NAQS_REQUIRE(t, NAQS_ERR_ARG, "naqs_rows_count: NULL table");
    NAQS_REQUIRE(M >= 0 && (M == 0 || (d_states && d_counts)), NAQS_ERR_ARG, "naqs_rows_count: NULL buffers");
    if (M == 0) return NAQS_OK;
    DeviceGuard guard(t->device);
    const int n_chunks = rows_chunks_for(t, M);
    int32_t* cc = nullptr;
    t->row_cc_valid = false;
    if (int rc = chunk_counts_buffer(t, n_chunks, M, &cc)) return rc;
    if (int rc = dispatch_rows<kRowsCount>(t, d_states, M, n_chunks, d_counts, cc, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream)) return rc;
    if (n_chunks > 1) {
        rows_sum_chunk_counts_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cc, n_chunks, M, d_counts);
        NAQS_LAUNCHED();
        t->row_cc_states = d_states; t->row_cc_M = M; t->row_cc_chunks = n_chunks; t->row_cc_stream = (cudaStream_t)stream;
        t->row_cc_valid = true;
    }
    return NAQS_OK;
}

int naqs_rows_fill(naqs_table_t* t, const uint64_t* d_states, int64_t M, const int64_t* d_indptr, uint64_t* d_col_keys,
                   int64_t* d_col_ridx, double* d_vals, void* stream) {
    NAQS_REQUIRE(t, NAQS_ERR_ARG, "naqs_rows_fill: NULL table");
    NAQS_REQUIRE(M >= 0 && (M == 0 || (d_states && d_indptr)), NAQS_ERR_ARG, "naqs_rows_fill: NULL buffers");  // col/val buffers may be NULL when nnz == 0
    if (M == 0) return NAQS_OK;
    DeviceGuard guard(t->device);
    if (d_col_ridx) { int rc = ensure_binom(t); if (rc) return rc; }
    // a chunked fill needs the per-(chunk, row) counts: those of the naqs_rows_count that produced this indptr when it was the last
    // rows call on this table for the same (d_states, M, stream) — used once — otherwise they are recomputed here
    const int n_chunks = rows_chunks_for(t, M);
    int32_t* cc = nullptr;
    const bool reuse = t->row_cc_valid && t->row_cc_states == d_states && t->row_cc_M == M && t->row_cc_chunks == n_chunks &&
                       t->row_cc_stream == (cudaStream_t)stream;
    t->row_cc_valid = false;
    if (int rc = chunk_counts_buffer(t, n_chunks, M, &cc)) return rc;
    if (n_chunks > 1 && !reuse)
        if (int rc = dispatch_rows<kRowsCount>(t, d_states, M, n_chunks, nullptr, cc, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream)) return rc;
    return dispatch_rows<kRowsFill>(t, d_states, M, n_chunks, nullptr, cc, d_indptr, d_col_keys, d_col_ridx, d_vals, (cudaStream_t)stream);
}

int naqs_hij_dense(naqs_table_t* t, const uint64_t* d_states, int64_t M, double* d_hij, void* stream) {
    NAQS_REQUIRE(t, NAQS_ERR_ARG, "naqs_hij_dense: NULL table");
    NAQS_REQUIRE(M >= 0 && (M == 0 || (d_states && d_hij)), NAQS_ERR_ARG, "naqs_hij_dense: NULL buffers");
    if (M == 0 || t->G == 0) return NAQS_OK;
    DeviceGuard guard(t->device);
    return dispatch_rows<kRowsDense>(t, d_states, M, rows_chunks_for(t, M), nullptr, nullptr, nullptr, nullptr, nullptr, d_hij, (cudaStream_t)stream);
}

int naqs_restricted_index(naqs_table_t* t, const uint64_t* d_keys, int64_t n, int64_t* d_out, void* stream) {
    NAQS_REQUIRE(t, NAQS_ERR_ARG, "naqs_restricted_index: NULL table");
    NAQS_REQUIRE(n >= 0 && (n == 0 || (d_keys && d_out)), NAQS_ERR_ARG, "naqs_restricted_index: NULL buffers");
    if (n == 0) return NAQS_OK;
    DeviceGuard guard(t->device);
    int rc = ensure_binom(t);
    if (rc) return rc;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    switch (t->nw32) {
        case 1: restricted_index_kernel<1><<<blocks, 256, 0, st>>>(t->sector, d_keys, n, rank_view(t), d_out); break;
        case 2: restricted_index_kernel<2><<<blocks, 256, 0, st>>>(t->sector, d_keys, n, rank_view(t), d_out); break;
        default: restricted_index_kernel<4><<<blocks, 256, 0, st>>>(t->sector, d_keys, n, rank_view(t), d_out); break;
    }
    NAQS_LAUNCHED();
    return NAQS_OK;
}

int naqs_exclusive_scan(naqs_table_t* t, const int64_t* d_counts, int64_t n, int64_t* d_indptr, void* stream) {
    NAQS_REQUIRE(t, NAQS_ERR_ARG, "naqs_exclusive_scan: NULL table");
    NAQS_REQUIRE(n >= 0 && d_indptr && (n == 0 || d_counts), NAQS_ERR_ARG, "naqs_exclusive_scan: NULL buffers");
    DeviceGuard guard(t->device);
    return exclusive_scan_i64(t, d_counts, n, d_indptr, (cudaStream_t)stream);
}

int naqs_unique_keys(naqs_table_t* t, const uint64_t* d_keys, int64_t n, uint64_t* d_out, int64_t* h_n_unique, void* stream) {
    NAQS_REQUIRE(t && h_n_unique, NAQS_ERR_ARG, "naqs_unique_keys: NULL argument");
    NAQS_REQUIRE(n >= 0 && (n == 0 || (d_keys && d_out)), NAQS_ERR_ARG, "naqs_unique_keys: NULL buffers");
    *h_n_unique = 0;
    if (n == 0) return NAQS_OK;
    DeviceGuard guard(t->device);
    return sort_unique_keys(t, d_keys, n, d_out, h_n_unique, (cudaStream_t)stream);
}

}  // extern "C"
