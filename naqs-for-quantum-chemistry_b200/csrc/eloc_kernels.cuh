// eloc_kernels.cuh — fused local-energy kernels (K1+K2+K3 of SURVEY.md §2) for sm_100a.
//
// eloc_direct_kernel: one thread owns R register-resident states; the Pauli table is staged tile by
// tile into shared memory and every term is read by the whole warp with ONE broadcast LDS per word.
// Per (state, term): AND(+XOR-fold across words) -> POPC -> sign bit XORed into the fp64 high word
// -> DADD, accumulated serially in reference term order inside each XY group
// (src_cpp/hamiltonian_math.pyx:31-34).  At a group end the coupled key s^u is formed, filtered by
// the sector (src/optimizer/hamiltonian.py:321-328) and by H != 0.0 (:363), looked up in the
// amplitude table (dense direct-address or 32 B-slot hash) and H*psi(s') is accumulated in complex128
// (src_cpp/sparse_math.pyx:87-100); the finalisation divides by psi(s) and conjugates
// (src/optimizer/energy.py:248).
#pragma once
#include "common.cuh"

namespace naqs {

template <int NW>
__device__ __forceinline__ void load_key(const uint64_t* __restrict__ keys, int64_t m, uint32_t (&s)[NW]) {
    if constexpr (NW == 1) {
        s[0] = (uint32_t)keys[m];
    } else if constexpr (NW == 2) {
        unsigned long long v = keys[m];
        s[0] = (uint32_t)v; s[1] = (uint32_t)(v >> 32);
    } else {
        const ulonglong2 v = reinterpret_cast<const ulonglong2*>(keys)[m];
        s[0] = (uint32_t)v.x; s[1] = (uint32_t)(v.x >> 32); s[2] = (uint32_t)v.y; s[3] = (uint32_t)(v.y >> 32);
    }
}

template <int NW>
__device__ __forceinline__ void key_words64(const uint32_t (&s)[NW], unsigned long long& k0, unsigned long long& k1) {
    k0 = s[0]; k1 = 0;
    if constexpr (NW >= 2) k0 |= (unsigned long long)s[1] << 32;
    if constexpr (NW == 4) k1 = (unsigned long long)s[2] | ((unsigned long long)s[3] << 32);
}

template <int NW>
__device__ __forceinline__ bool in_sector(const uint32_t (&j)[NW], const Sector& sec) {
    int na = 0, nb = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        na += __popc(j[w] & sec.even[w]);
        nb += __popc(j[w] & sec.odd[w]);
    }
    return na == sec.n_alpha && nb == sec.n_beta;
}

__device__ __forceinline__ double2 load_psi(const void* __restrict__ psi, int dtype, int64_t m) {
    if (dtype == NAQS_C64) {
        const float2 v = reinterpret_cast<const float2*>(psi)[m];
        return make_double2((double)v.x, (double)v.y);
    }
    return reinterpret_cast<const double2*>(psi)[m];
}

// the 4 keys of a bucket with ONE 256-bit load (sm_100a LDG.E.256): one request, one 32-byte sector
struct BucketKeys {
    unsigned long long k[4];
};
__device__ __forceinline__ BucketKeys load_bucket_keys(const HashBucket* __restrict__ bk) {
    BucketKeys r;
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(r.k[0]), "=l"(r.k[1]), "=l"(r.k[2]), "=l"(r.k[3]) : "l"(bk->key));
    return r;
}
__device__ __forceinline__ int bucket_match(const BucketKeys& b, unsigned long long k0) {
    int hs = -1;
    if ((b.k[0] & kKeyMask63) == k0) hs = 0;
    if (b.k[1] == k0) hs = 1;
    if (b.k[2] == k0) hs = 2;
    if (b.k[3] == k0) hs = 3;
    return hs;
}
// one probe of the bucketed hash: -> true when the search is over (hit or definite miss)
__device__ __forceinline__ bool bucket_probe(const HashBucket* __restrict__ bk, unsigned long long k0, int& hit_slot) {
    const BucketKeys b = load_bucket_keys(bk);
    hit_slot = bucket_match(b, k0);
    return hit_slot >= 0 || !(b.k[0] & kOverflowFlag);
}

// psi(s') from the amplitude table; (0,0) when s' was not sampled.
template <int NW>
__device__ __forceinline__ double2 lookup_psi(const LookupView& lv, const uint32_t (&j)[NW]) {
    unsigned long long k0, k1;
    key_words64<NW>(j, k0, k1);
    if (lv.kind == NAQS_LOOKUP_DENSE) {
        return __ldg(lv.dense + k0);
    }
    if constexpr (NW <= 2) {
        unsigned b = hash32(k0, 0ull) >> lv.bshift;
        while (true) {
            int hs;
            const bool done = bucket_probe(lv.buckets + b, k0, hs);
            if (hs >= 0) return __ldg(&lv.buckets[b].psi[hs]);
            if (done) return make_double2(0.0, 0.0);
            b = (b + 1) & lv.bmask;
        }
    } else {
        unsigned long long h = hash_slot(k0, k1, lv.shift);
        while (true) {
            const HashSlot* sl = lv.slots + h;
            const ulonglong2 kk = __ldg(reinterpret_cast<const ulonglong2*>(sl));
            if (kk.x == k0 && kk.y == k1) return __ldg(reinterpret_cast<const double2*>(sl) + 1);
            if (kk.x == kEmptyKey && kk.y == kEmptyKey) return make_double2(0.0, 0.0);
            h = (h + 1) & lv.mask;
        }
    }
}

// numpy's complex128 division (Smith's algorithm), every operation individually rounded so the
// result matches the host reference bit for bit; followed by the conjugation of energy.py:248.
__device__ __forceinline__ double2 div_conj(double2 a, double2 b) {
    double qr, qi;
    const double br = fabs(b.x), bi = fabs(b.y);
    if (br >= bi) {
        if (br == 0.0 && bi == 0.0) {
            qr = __ddiv_rn(a.x, br); qi = __ddiv_rn(a.y, bi);
        } else {
            const double rat = __ddiv_rn(b.y, b.x);
            const double scl = __ddiv_rn(1.0, __dadd_rn(b.x, __dmul_rn(b.y, rat)));
            qr = __dmul_rn(__dadd_rn(a.x, __dmul_rn(a.y, rat)), scl);
            qi = __dmul_rn(__dadd_rn(a.y, -__dmul_rn(a.x, rat)), scl);
        }
    } else {
        const double rat = __ddiv_rn(b.x, b.y);
        const double scl = __ddiv_rn(1.0, __dadd_rn(b.y, __dmul_rn(b.x, rat)));
        qr = __dmul_rn(__dadd_rn(__dmul_rn(a.x, rat), a.y), scl);
        qi = __dmul_rn(__dadd_rn(__dmul_rn(a.y, rat), -a.x), scl);
    }
    return make_double2(qr, -qi);
}

// finalisation of one row: E_loc = conj(S / psi) (energy.py:248), or the raw sum S = (H psi)[row] when psi == nullptr
// (matrix-free H.v for solve_H / calculate_energy)
__device__ __forceinline__ double2 finalize_row(double2 sum, const void* __restrict__ psi, int psi_dtype, int64_t m) {
    return psi ? div_conj(sum, load_psi(psi, psi_dtype, m)) : sum;
}

// +c or -c according to the parity of popcount(f): the sign bit goes straight into the high word.
__device__ __forceinline__ double signed_coeff(int c_hi, int c_lo, uint32_t f) {
    return __hiloint2double(c_hi ^ (int)(__popc(f) << 31), c_lo);
}

// Shared-memory layout of one staged tile (dynamic smem):
//   double c[tile_cap] | u32 yz[NW][tile_cap] | u32 gstart[tile_cap + 2] | u32 gxy[NW][tile_cap + 1]
template <int NW>
__host__ __device__ inline size_t tile_smem_bytes(int tile_cap) {
    return (size_t)tile_cap * 8 + (size_t)NW * tile_cap * 4 + (size_t)(tile_cap + 2) * 4 + (size_t)NW * (tile_cap + 1) * 4;
}

// Walk the whole term table for R register-resident states per thread.  `emit(g, u, h)` is called
// once per XY group g (global index) with the flip mask u and the R accumulated matrix elements
// h[r] = H[s_r, s_r ^ u], summed serially in reference term order.  Must be called by all threads of
// the CTA (it synchronises on the staged tiles).
template <int NW, int R, int THREADS, class Emit>
__device__ __forceinline__ void walk_table(const TableView& tv, const Tile* __restrict__ tiles, int n_tiles, int tile_cap,
                                           const uint32_t (&s)[R][NW], Emit&& emit) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_c = reinterpret_cast<double*>(smem_raw);
    uint32_t* s_yz = reinterpret_cast<uint32_t*>(s_c + tile_cap);
    uint32_t* s_gstart = s_yz + (size_t)NW * tile_cap;
    uint32_t* s_gxy = s_gstart + tile_cap + 2;

    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;

    for (int ti = 0; ti < n_tiles; ++ti) {
        const Tile tile = tiles[ti];
        const int nt = tile.t1 - tile.t0, ng = tile.g1 - tile.g0;
        __syncthreads();  // previous tile fully consumed
        for (int i = threadIdx.x; i < nt; i += THREADS) {
            s_c[i] = tv.coeff[tile.t0 + i];
#pragma unroll
            for (int w = 0; w < NW; ++w) s_yz[w * tile_cap + i] = tv.yz[(size_t)w * tv.K + tile.t0 + i];
        }
        for (int i = threadIdx.x; i <= ng; i += THREADS) s_gstart[i] = tv.gstart[tile.g0 + i];
        for (int i = threadIdx.x; i < ng; i += THREADS) {
#pragma unroll
            for (int w = 0; w < NW; ++w) s_gxy[w * (tile_cap + 1) + i] = tv.gxy[(size_t)w * tv.G + tile.g0 + i];
        }
        __syncthreads();

        for (int g = 0; g < ng; ++g) {
            const uint32_t gs = s_gstart[g], ge = s_gstart[g + 1];
            const int beg = (int)(max(gs, tile.t0) - tile.t0);
            const int end = (int)(min(ge, tile.t1) - tile.t0);
#pragma unroll 2
            for (int k = beg; k < end; ++k) {
                const double c = s_c[k];
                const int c_hi = __double2hiint(c), c_lo = __double2loint(c);
                uint32_t yz[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) yz[w] = s_yz[w * tile_cap + k];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    uint32_t f = s[r][0] & yz[0];
#pragma unroll
                    for (int w = 1; w < NW; ++w) f ^= s[r][w] & yz[w];
                    acc[r] = __dadd_rn(acc[r], signed_coeff(c_hi, c_lo, f));
                    // float32 table (hamiltonian_math.pyx __inner_int64_float): both operands are float32 values, so
                    // rounding their double sum to float32 equals the float32 addition (53 >= 2*24 + 2 bits)
                    if (tv.f32) acc[r] = (double)__double2float_rn(acc[r]);
                }
            }
            if (ge <= tile.t1) {  // group complete (a straddling group continues in the next tile)
                uint32_t u[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) u[w] = s_gxy[w * (tile_cap + 1) + g];
                emit((int)(tile.g0 + g), u, acc);
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = 0.0;
            }
        }
    }
}

template <int NW, int R, int THREADS>
__device__ __forceinline__ void load_states(const uint64_t* __restrict__ states, int64_t M, uint32_t (&s)[R][NW],
                                            bool (&valid)[R]) {
    const int64_t base = (int64_t)blockIdx.x * (THREADS * R) + threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t m = base + (int64_t)r * THREADS;
        valid[r] = m < M;
        if (valid[r]) {
            load_key<NW>(states, m, s[r]);
        } else {
#pragma unroll
            for (int w = 0; w < NW; ++w) s[r][w] = 0;
        }
    }
}

// grid = (state blocks, table chunks): with more than one chunk (small batches: the chunk list of the stored-row kernels, cut on
// group boundaries) the raw row sums of chunk c go to partial[c * M + m] and eloc_finalize_kernel adds them in chunk order.
template <int NW, int R, int THREADS>
__global__ void __launch_bounds__(THREADS)
eloc_direct_kernel(TableView tv, const Tile* __restrict__ tiles_all, const __grid_constant__ ChunkBounds chunks, int n_chunks, int tile_cap, Sector sec,
                   LookupView lv, const uint64_t* __restrict__ states, const void* __restrict__ psi, int psi_dtype, int64_t M,
                   double2* __restrict__ out, double2* __restrict__ partial) {
    const Tile* tiles = tiles_all + chunks.lo[blockIdx.y];
    const int n_tiles = chunks.lo[blockIdx.y + 1] - chunks.lo[blockIdx.y];
    uint32_t s[R][NW];
    bool valid[R];
    load_states<NW, R, THREADS>(states, M, s, valid);
    double e_re[R], e_im[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { e_re[r] = 0.0; e_im[r] = 0.0; }

    walk_table<NW, R, THREADS>(tv, tiles, n_tiles, tile_cap, s, [&](int, const uint32_t (&u)[NW], const double (&h)[R]) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (h[r] != 0.0 && valid[r]) {
                uint32_t j[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) j[w] = s[r][w] ^ u[w];
                if (!sec.enabled || in_sector<NW>(j, sec)) {
                    const double2 p = lookup_psi<NW>(lv, j);
                    // real x complex product, then add: two roundings each, as the host reference's
                    // `out[j] + data * v` (sparse_math.pyx:98) without FMA contraction
                    e_re[r] = __dadd_rn(e_re[r], __dmul_rn(h[r], p.x));
                    e_im[r] = __dadd_rn(e_im[r], __dmul_rn(h[r], p.y));
                }
            }
        }
    });

    const int64_t base = (int64_t)blockIdx.x * (THREADS * R) + threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t m = base + (int64_t)r * THREADS;
        if (!valid[r]) continue;
        if (n_chunks > 1) partial[(int64_t)blockIdx.y * M + m] = make_double2(e_re[r], e_im[r]);
        else out[m] = finalize_row(make_double2(e_re[r], e_im[r]), psi, psi_dtype, m);
    }
}

// ---------------------------------------------------------------------------------------------
// Stored-row kernels (CSR / coupled-set mode, K2 of SURVEY.md §2).  One state per thread.
// A stored entry = coupled state inside the sector with H != 0.0 (hamiltonian.py:328,363).
constexpr int kRowsCount = 0, kRowsFill = 1, kRowsDense = 2;

// Combinatorial rank of a sector key in the reference's restricted order (hilbert.py:446-469):
// alpha combination (even qubits) outer, beta (odd qubits) inner, each in itertools.combinations
// (lexicographic) order.  binom[n*66 + k] = C(n, k) as int64 (n <= 64, k <= 65).
template <int NW>
__device__ __forceinline__ long long lex_rank(const uint32_t (&j)[NW], int start, int n, int k, const long long* __restrict__ binom) {
    long long r = 0;
    int seen = 0;
    for (int p = 0; p < n && seen < k; ++p) {
        const int q = start + 2 * p;
        if ((j[q >> 5] >> (q & 31)) & 1u) ++seen;
        else r += binom[(n - 1 - p) * 66 + (k - 1 - seen)];
    }
    return r;
}

// Ranker inputs: the binomial table, and — for keys of <= 32 qubits — the rank of every alpha / beta occupation pattern
// (tab_a[pattern of the even qubits], tab_b[pattern of the odd qubits], 2^ceil(N/2) / 2^floor(N/2) int32 entries): the analogue
// of the reference's 2^N lookup table (hilbert.py:607-640) factorised per spin, two loads instead of a loop over the qubits.
struct RankView {
    const long long* binom;
    const int32_t* tab_a;
    const int32_t* tab_b;
    long long n_b;  // C(n_odd, n_beta)
};

__device__ __forceinline__ uint32_t compress_even_bits(uint32_t x) {  // bits 0, 2, 4, ... -> bits 0, 1, 2, ...
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}

template <int NW>
__device__ __forceinline__ long long restricted_index(const uint32_t (&j)[NW], const Sector& sec, const RankView& rv) {
    if (!sec.enabled) {
        unsigned long long k0, k1;
        key_words64<NW>(j, k0, k1);
        return (long long)k0;
    }
    if (!in_sector<NW>(j, sec)) return -1;
    if (NW == 1 && rv.tab_a)
        return (long long)rv.tab_a[compress_even_bits(j[0])] * rv.n_b + (long long)rv.tab_b[compress_even_bits(j[0] >> 1)];
    const int n_even = (sec.n_qubits + 1) / 2, n_odd = sec.n_qubits / 2;
    return lex_rank<NW>(j, 0, n_even, sec.n_alpha, rv.binom) * rv.binom[n_odd * 66 + sec.n_beta] +
           lex_rank<NW>(j, 1, n_odd, sec.n_beta, rv.binom);
}

// grid = (row blocks, table chunks).  A chunk is a contiguous range of tiles that starts and ends on a group boundary
// (naqs_table_create builds the list), so every group is summed by ONE thread in reference term order whatever the number of
// chunks.  With more than one chunk the count pass writes per-(chunk, row) counts (chunk_counts[c * M + m], int32) and the fill
// pass starts chunk c of row m at indptr[m] + sum_{c' < c} chunk_counts[c' * M + m]: a row's entries stay in ascending group order.
template <int NW, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS)
rows_kernel(TableView tv, const Tile* __restrict__ tiles, const __grid_constant__ ChunkBounds chunks, int n_chunks, int tile_cap, Sector sec,
            const uint64_t* __restrict__ states, int64_t M, int words, RankView rank,
            int64_t* __restrict__ counts, int32_t* __restrict__ chunk_counts, const int64_t* __restrict__ indptr,
            uint64_t* __restrict__ col_keys, int64_t* __restrict__ col_ridx, double* __restrict__ vals) {
    uint32_t s[1][NW];
    bool valid[1];
    load_states<NW, 1, THREADS>(states, M, s, valid);
    const int64_t m = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    const int chunk = blockIdx.y;
    int64_t n = 0;
    int64_t e = 0;
    if (MODE == kRowsFill && valid[0]) {
        e = indptr[m];
        for (int c = 0; c < chunk; ++c) e += chunk_counts[(int64_t)c * M + m];
    }
    walk_table<NW, 1, THREADS>(tv, tiles + chunks.lo[chunk], chunks.lo[chunk + 1] - chunks.lo[chunk], tile_cap, s,
                               [&](int g, const uint32_t (&u)[NW], const double (&h)[1]) {
        if (!valid[0]) return;
        if (MODE == kRowsDense) {
            vals[m * tv.G + g] = h[0];
            return;
        }
        if (h[0] == 0.0) return;
        uint32_t j[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) j[w] = s[0][w] ^ u[w];
        if (sec.enabled && !in_sector<NW>(j, sec)) return;
        if (MODE == kRowsCount) {
            ++n;
        } else {
            unsigned long long k0, k1;
            key_words64<NW>(j, k0, k1);
            col_keys[e * words] = k0;
            if (words > 1) col_keys[e * words + 1] = k1;
            if (col_ridx) col_ridx[e] = restricted_index<NW>(j, sec, rank);
            vals[e] = h[0];
            ++e;
        }
    });
    if (MODE == kRowsCount && valid[0]) {
        if (n_chunks > 1) chunk_counts[(int64_t)chunk * M + m] = (int32_t)n;
        else counts[m] = n;
    }
}

}  // namespace naqs
