// table.cu — table handle, amplitude lookup construction and the fused E_loc entry points.
// C ABI documented in include/naqs_eloc.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#include "common.cuh"
#include "eloc_kernels.cuh"
#include "sliced.cuh"
#include "keyorder.cuh"

namespace naqs {

static thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};
void set_error(const std::string& msg) { g_error = msg; }

int ensure_ws(naqs_table* t, size_t bytes) {
    if (bytes <= t->ws_bytes) return NAQS_OK;
    if (t->d_ws) cudaFree(t->d_ws);
    t->d_ws = nullptr; t->ws_bytes = 0;
    size_t want = std::max(bytes, (size_t)1 << 20);
    NAQS_CUDA(cudaMalloc(&t->d_ws, want));
    t->ws_bytes = want;
    return NAQS_OK;
}

// The per-table buffers (lookup structures, partial sums, staging) are reused from call to call.  Calls on ONE stream are
// ordered by the stream; when the caller switches streams (e.g. the host-buffer entry runs on the table's own stream, the
// device-tensor entries on the caller's) the previous stream is drained first, so a rebuild can never overtake a kernel
// that still reads the old table.
static int stream_handover(naqs_table* t, cudaStream_t stream) {
    if (t->last_stream_valid && t->last_stream != stream) NAQS_CUDA(cudaStreamSynchronize(t->last_stream));
    t->last_stream = stream;
    t->last_stream_valid = true;
    return NAQS_OK;
}

// ------------------------------------------------------------------------------------------ lookup build
// Only keys INSIDE the sector enter a lookup structure: a coupled state s ^ u outside the sector can then never be found,
// which IS the reference's sector filter on coupled states (hamiltonian.py:321-328) — applied once per table key here
// instead of once per (state, group) pair in the fused kernel.  (Sampled states are in-sector by contract; a stray one
// is ignored exactly as the reference ignores it.)
// A key with bits at or above n_qubits is not a state index at all (the reference raises IndexError, hilbert.py:607-640 indexes
// a 2^N LUT with it): it is never used as an address, bit 0 of the table's flag word is raised and naqs_table_check reports it.
__device__ __forceinline__ bool key_in_range(const uint64_t* __restrict__ key, int words, int n_qubits, int* __restrict__ flags) {
    bool ok = true;
    for (int w = 0; w < words; ++w) {
        const int lo = 64 * w;
        if (n_qubits <= lo) ok = ok && key[w] == 0ull;
        else if (n_qubits < lo + 64) ok = ok && (key[w] >> (n_qubits - lo)) == 0ull;
    }
    if (!ok && flags) atomicOr(flags, 1);
    return ok;
}

__device__ __forceinline__ bool key_in_sector(const uint64_t* __restrict__ key, int words, const Sector& sec) {
    if (!sec.enabled) return true;
    int na = 0, nb = 0;
    for (int w = 0; w < words; ++w) {
        const unsigned long long ev = (unsigned long long)sec.even[2 * w] | ((unsigned long long)sec.even[2 * w + 1] << 32);
        const unsigned long long od = (unsigned long long)sec.odd[2 * w] | ((unsigned long long)sec.odd[2 * w + 1] << 32);
        na += __popcll(key[w] & ev);
        nb += __popcll(key[w] & od);
    }
    return na == sec.n_alpha && nb == sec.n_beta;
}

__global__ void hash_init_kernel(HashSlot* slots, int64_t cap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) {
        // one 32 B store per slot
        ulonglong4 v = make_ulonglong4(kEmptyKey, kEmptyKey, 0ull, 0ull);
        reinterpret_cast<ulonglong4*>(slots)[i] = v;
    }
}

__device__ __forceinline__ ulonglong2 cas128(unsigned long long* addr, ulonglong2 cmp, ulonglong2 val) {
    ulonglong2 old;
    asm volatile(
        "{\n .reg .b128 d, c, s;\n mov.b128 c, {%2, %3};\n mov.b128 s, {%4, %5};\n"
        " atom.global.cas.b128 d, [%6], c, s;\n mov.b128 {%0, %1}, d;\n}"
        : "=l"(old.x), "=l"(old.y)
        : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(addr)
        : "memory");
    return old;
}

__global__ void hash_insert_kernel(HashSlot* slots, unsigned long long mask, int shift, const uint64_t* __restrict__ keys, int words,
                                   const void* __restrict__ psi, int psi_dtype, int64_t n, Sector sec, int* __restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !key_in_range(keys + i * words, words, sec.n_qubits, flags) || !key_in_sector(keys + i * words, words, sec)) return;
    const unsigned long long k0 = keys[i * words], k1 = words > 1 ? keys[i * words + 1] : 0ull;
    const double2 p = load_psi(psi, psi_dtype, i);
    unsigned long long h = hash_slot(k0, k1, shift);
    while (true) {
        HashSlot* sl = slots + h;
        const ulonglong2 old = cas128(sl->key, make_ulonglong2(kEmptyKey, kEmptyKey), make_ulonglong2(k0, k1));
        if ((old.x == kEmptyKey && old.y == kEmptyKey) || (old.x == k0 && old.y == k1)) {
            // duplicates of a key are summed (scipy's H[idx[:,None], idx] repeats the column)
            atomicAdd(&sl->re, p.x);
            atomicAdd(&sl->im, p.y);
            return;
        }
        h = (h + 1) & mask;
    }
}

__global__ void bucket_init_kernel(HashBucket* buckets, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 32 B quarter of a bucket
    if (i >= 4 * n) return;
    ulonglong4 v = (i & 3) == 0 ? make_ulonglong4(kEmptyKey63, kEmptyKey63, kEmptyKey63, kEmptyKey63) : make_ulonglong4(0, 0, 0, 0);
    reinterpret_cast<ulonglong4*>(buckets)[i] = v;
}

__global__ void bucket_insert_kernel(HashBucket* buckets, unsigned bmask, int bshift, const uint64_t* __restrict__ keys, int words,
                                     const void* __restrict__ psi, int psi_dtype, int64_t n, int keep_one, Sector sec, int* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !key_in_range(keys + i * words, words, sec.n_qubits, flags) || !key_in_sector(keys + i * words, words, sec)) return;
    const unsigned long long k = keys[i * words];
    const double2 p = load_psi(psi, psi_dtype, i);
    unsigned b = hash32(k, 0ull) >> bshift;
    while (true) {
        HashBucket* bk = buckets + b;
        for (int s = 0; s < 4; ++s) {
            const unsigned long long old = atomicCAS(&bk->key[s], kEmptyKey63, k);
            if (old == kEmptyKey63 || (old & kKeyMask63) == k) {
                if (keep_one) {  // copies carry the same amplitude: the thread that claimed the slot stores it
                    if (old == kEmptyKey63) bk->psi[s] = p;
                    return;
                }
                // duplicates of a key are summed (scipy's H[idx[:,None], idx] repeats the column)
                atomicAdd(&bk->psi[s].x, p.x);
                atomicAdd(&bk->psi[s].y, p.y);
                return;
            }
        }
        atomicOr(&bk->key[0], kOverflowFlag);  // full: later searches must continue with the next bucket
        b = (b + 1) & bmask;
    }
}

// Bloom filter over the in-sector table keys from the GF(2)-linear hashes of common.cuh: word = hw & mask (byte offset), bits
// rotl(kFilterPattern, hb).  small != nullptr: also the 2^15-word companion (same bits, the low 15 bits of the word index).
__global__ void filter_build_kernel(uint32_t* filter, uint32_t mask, uint32_t* small, const uint64_t* __restrict__ keys, int words, int64_t n, Sector sec,
                                    int* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !key_in_range(keys + i * words, words, sec.n_qubits, flags) || !key_in_sector(keys + i * words, words, sec)) return;
    uint32_t hw, hb;
    lin_hash_key64(keys + i * words, words, hw, hb);
    const uint32_t bits = filter_insert_bits(hb);
    atomicOr(&filter[(hw & mask) >> 2], bits);
    if (small) atomicOr(&small[(hw & (kFilterBytes - 4u)) >> 2], bits);
}

// Bank-binned order of a hash-lookup batch: a state's filter BANK is bits 2..6 of its word hash and the bank of a coupled
// state is bank(s) ^ bank(u) (linear hashes), so a warp whose lane l holds a state of bank l reads 32 different banks for
// every group.  perm[rank * 32 + bank] = row for the first R states of a bank, later ones go to an overflow region behind
// the 32 * R regular positions (a skewed batch stays correct, only conflict-prone).  perm is preset to -1, counters to 0.
template <int NW>
__global__ void __launch_bounds__(1024) bin_states_kernel(const uint64_t* __restrict__ states, int64_t M, int32_t* __restrict__ perm,
                                                          int32_t* __restrict__ counters, int64_t R) {
    __shared__ int cnt[32], base[32];
    if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bank = 0;
    int r = 0;
    if (m < M) {
        uint32_t s[NW], hw = 0, hb = 0;
        load_key<NW>(states, m, s);
#pragma unroll
        for (int w = 0; w < NW; ++w) lin_hash_word(s[w], w, hw, hb);
        bank = (hw >> 2) & 31u;
        r = atomicAdd(&cnt[bank], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) base[threadIdx.x] = atomicAdd(&counters[threadIdx.x], cnt[threadIdx.x]);
    __syncthreads();
    if (m < M) {
        const int64_t rank = (int64_t)base[bank] + r;
        const int64_t pos = rank < R ? rank * 32 + bank : 32 * R + atomicAdd(&counters[32], 1);
        perm[pos] = (int32_t)m;
    }
}

__global__ void widen_keys_kernel(const void* __restrict__ in, int itemsize, int64_t n, uint64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = itemsize == 4 ? (uint64_t)reinterpret_cast<const uint32_t*>(in)[i] : (uint64_t)reinterpret_cast<const uint16_t*>(in)[i];
}

__global__ void narrow_eloc_kernel(const double2* __restrict__ in, int64_t n, float2* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float2((float)in[i].x, (float)in[i].y);
}

__global__ void dense_scatter32_kernel(float2* dense, const uint64_t* __restrict__ keys, const float2* __restrict__ psi, int64_t n, Sector sec,
                                       int* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && key_in_range(keys + i, 1, sec.n_qubits, flags) && key_in_sector(keys + i, 1, sec)) dense[keys[i]] = psi[i];  // unique keys (caller's guarantee): a plain store
}

__global__ void dense_scatter_kernel(double2* dense, const uint64_t* __restrict__ keys, const void* __restrict__ psi,
                                     int psi_dtype, int64_t n, int overwrite, Sector sec, int* __restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !key_in_range(keys + i, 1, sec.n_qubits, flags) || !key_in_sector(keys + i, 1, sec)) return;
    const double2 p = load_psi(psi, psi_dtype, i);
    if (overwrite) { dense[keys[i]] = p; return; }  // copies of a key carry the same amplitude: keep one
    double* dst = reinterpret_cast<double*>(dense + keys[i]);
    atomicAdd(dst, p.x);  // duplicates of a key are summed (scipy's H[idx[:,None], idx] repeats the column)
    atomicAdd(dst + 1, p.y);
}

// ------------------------------------------------------------------------------------------ launch helpers
constexpr int kThreads = 256;

// sliced kernel launch shapes: threads per CTA and shared-memory tile capacity (two buffers per CTA)
// launch shapes [0..2] dense lookup: 1024/512/256 threads, 1/2/4 CTAs per SM (64 registers per thread);
//               [3..5] hash lookup : same thread counts, smaller tiles: 32-64 B/thread coupling queue, and a 128 KB Bloom filter
//                                    in the 1-CTA-per-SM shape.  (Measured alternatives: 512 x 1 CTA/SM with 128 registers
//                                    1.70 ms, 896 threads 1.06 ms, 1024 threads 0.99 ms on Li2O 1e5 — occupancy wins.)
constexpr int kSlicedThreads[6] = {1024, 512, 256, 1024, 512, 256};
constexpr int kSlicedCtasPerSm[6] = {1, 2, 4, 1, 2, 4};
constexpr bool kSlicedFilter[6] = {false, false, false, true, false, false};
constexpr size_t kSlicedCap[6] = {112640, 55296, 26624, 33792, 36864, 18432};  // [3]: 2 x 33 KB tiles + 32 KB queue + 128 KB filter
constexpr size_t kSlicedMaxBlob = 16384;
constexpr size_t kKoMaxBlobWords = 5;  // key-order stream: parity words (30 terms each) of a big group multiplied by psi together
static_assert(kSlicedCap[3] < 65536 && kSlicedCap[4] < 65536 && kSlicedCap[5] < 65536, "queue entries of the hash shapes hold 16-bit byte offsets into a tile");

static int tile_cap_for(int nw32, int64_t K) {
    // keep a tile <= ~56 KB so four CTAs of 256 threads fit one SM; whole table in one tile when it fits
    const int per_term = 8 + 8 * nw32 + 4;
    int cap = (56 * 1024 - 64) / per_term;
    cap = std::max(cap, 256);
    return (int)std::min<int64_t>(std::max<int64_t>(K, 1), cap);
}

}  // namespace naqs

using namespace naqs;

extern "C" {

const char* naqs_last_error(void) { return g_error.c_str(); }
int naqs_abi_version(void) { return NAQS_ABI_VERSION; }
int naqs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int64_t naqs_launch_count(void) { return g_launches.load(); }

int naqs_table_create(naqs_table_t** out, const uint64_t* h_xy, const uint64_t* h_yz, const double* h_coeff,
                      int64_t K, int words, int n_qubits, int n_alpha, int n_beta, int device) {
    NAQS_REQUIRE(out != nullptr, NAQS_ERR_ARG, "naqs_table_create: out is NULL");
    *out = nullptr;
    NAQS_REQUIRE(K >= 0 && (K == 0 || (h_xy && h_yz && h_coeff)), NAQS_ERR_ARG, "naqs_table_create: NULL term arrays");
    NAQS_REQUIRE(words == 1 || words == 2, NAQS_ERR_ARG, "naqs_table_create: words must be 1 or 2");
    NAQS_REQUIRE(n_qubits >= 1 && n_qubits <= 64 * words - 1, NAQS_ERR_ARG,
                 "naqs_table_create: n_qubits must be in [1, 63] (words=1) or [1, 127] (words=2)");
    NAQS_REQUIRE(K < (1ll << 31), NAQS_ERR_ARG, "naqs_table_create: too many terms");
    NAQS_REQUIRE((n_alpha < 0) == (n_beta < 0), NAQS_ERR_ARG, "naqs_table_create: give both n_alpha and n_beta or neither");
    NAQS_REQUIRE(naqs_device_count() > device && device >= 0, NAQS_ERR_CUDA, "naqs_table_create: no such CUDA device (this library has no CPU fallback)");
    DeviceGuard guard(device);
    NAQS_REQUIRE(guard.ok, NAQS_ERR_CUDA, "naqs_table_create: cudaSetDevice failed");

    auto* t = new naqs_table();
    t->device = device; t->words = words; t->n_qubits = n_qubits; t->n_alpha = n_alpha; t->n_beta = n_beta; t->K = K;
    // keys of up to 63 bits take the bucketed hash (bit 63 is its overflow flag); 64..127 qubits use 128-bit keys
    t->nw32 = n_qubits <= 32 ? 1 : (n_qubits <= 63 ? 2 : 4);
    const int NW = t->nw32;

    // group by XY mask: ascending multi-word value (np.unique), stable in k
    std::vector<int64_t> order(K);
    std::iota(order.begin(), order.end(), 0);
    auto key_less = [&](const uint64_t* a, const uint64_t* b) {
        for (int w = words - 1; w >= 0; --w) if (a[w] != b[w]) return a[w] < b[w];
        return false;
    };
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key_less(h_xy + a * words, h_xy + b * words); });
    std::vector<uint32_t> yz((size_t)NW * std::max<int64_t>(K, 1)), gxy, gstart;
    std::vector<double> coeff(std::max<int64_t>(K, 1));
    std::vector<const uint64_t*> gkeys;
    for (int64_t i = 0; i < K; ++i) {
        const int64_t k = order[i];
        if (i == 0 || std::memcmp(h_xy + k * words, h_xy + order[i - 1] * words, 8 * words) != 0) {
            gstart.push_back((uint32_t)i);
            gkeys.push_back(h_xy + k * words);
        }
        for (int w = 0; w < NW; ++w) yz[(size_t)w * K + i] = (uint32_t)(h_yz[k * words + w / 2] >> (32 * (w % 2)));
        coeff[i] = h_coeff[k];
    }
    const int64_t G = (int64_t)gstart.size();
    gstart.push_back((uint32_t)K);
    gxy.resize((size_t)NW * std::max<int64_t>(G, 1));
    for (int64_t g = 0; g < G; ++g)
        for (int w = 0; w < NW; ++w) gxy[(size_t)w * G + g] = (uint32_t)(gkeys[g][w / 2] >> (32 * (w % 2)));
    t->G = G;
    {   // Kyz for naqs_table_info
        std::vector<int64_t> o2(K);
        std::iota(o2.begin(), o2.end(), 0);
        std::sort(o2.begin(), o2.end(), [&](int64_t a, int64_t b) { return key_less(h_yz + a * words, h_yz + b * words); });
        int64_t n = 0;
        for (int64_t i = 0; i < K; ++i)
            if (i == 0 || std::memcmp(h_yz + o2[i] * words, h_yz + o2[i - 1] * words, 8 * words) != 0) ++n;
        t->Kyz = n;
    }
    // sector masks
    Sector& sec = t->sector;
    std::memset(&sec, 0, sizeof(sec));
    sec.enabled = n_alpha >= 0; sec.n_alpha = n_alpha; sec.n_beta = n_beta; sec.n_qubits = n_qubits;
    for (int q = 0; q < n_qubits; ++q) ((q % 2 == 0) ? sec.even : sec.odd)[q / 32] |= 1u << (q % 32);

    // tiles
    t->tile_cap = tile_cap_for(NW, K);
    std::vector<Tile> tiles;
    for (int64_t t0 = 0; t0 < K; t0 += t->tile_cap) {
        Tile tl;
        tl.t0 = (uint32_t)t0; tl.t1 = (uint32_t)std::min<int64_t>(K, t0 + t->tile_cap);
        // first group with gstart[g+1] > t0 ; one past the last group with gstart[g] < t1
        tl.g0 = (uint32_t)(std::upper_bound(gstart.begin(), gstart.end(), tl.t0) - gstart.begin() - 1);
        tl.g1 = (uint32_t)(std::lower_bound(gstart.begin(), gstart.end(), tl.t1) - gstart.begin());
        tiles.push_back(tl);
    }
    t->n_tiles = (int)tiles.size();
    // chunked tile list for the stored-row kernels: kMaxChunks contiguous group ranges of about equal term count
    std::vector<Tile> row_tiles;
    if (G >= 4 * kMaxChunks) {
        int64_t g = 0;
        int max_terms = 1;
        for (int c = 0; c < kMaxChunks; ++c) {
            t->row_chunk_lo[c] = (int)row_tiles.size();
            const int64_t want_end = K * (c + 1) / kMaxChunks;
            int64_t g_end = g;
            while (g_end < G && (c + 1 == kMaxChunks || (int64_t)gstart[(size_t)g_end + 1] <= want_end || g_end == g)) ++g_end;
            if (c + 1 == kMaxChunks) g_end = G;
            for (int64_t t0 = gstart[(size_t)g]; t0 < (int64_t)gstart[(size_t)g_end]; t0 += t->tile_cap) {
                Tile tl;
                tl.t0 = (uint32_t)t0; tl.t1 = (uint32_t)std::min<int64_t>(gstart[(size_t)g_end], t0 + t->tile_cap);
                tl.g0 = (uint32_t)(std::upper_bound(gstart.begin(), gstart.end(), tl.t0) - gstart.begin() - 1);
                tl.g1 = (uint32_t)(std::lower_bound(gstart.begin(), gstart.end(), tl.t1) - gstart.begin());
                max_terms = std::max<int>(max_terms, (int)(tl.t1 - tl.t0));
                row_tiles.push_back(tl);
            }
            g = g_end;
        }
        t->row_chunk_lo[kMaxChunks] = (int)row_tiles.size();
        t->n_row_chunks = kMaxChunks;
        t->row_tile_cap = max_terms;
    }

    // sliced (v2) stream: groups in ascending-XY order with their terms in reference order
    SlicedHost sh;
    std::vector<STile> stiles[6];
    {
        std::vector<HostGroup> hg((size_t)G);
        for (int64_t g = 0; g < G; ++g) {
            HostGroup& x = hg[(size_t)g];
            for (int w = 0; w < 4; ++w) x.u[w] = w < NW ? gxy[(size_t)w * G + g] : 0u;
            const uint32_t b = gstart[(size_t)g], e = gstart[(size_t)g + 1];
            x.yz.resize((size_t)(e - b) * NW);
            x.c.assign(coeff.begin() + b, coeff.begin() + e);
            for (uint32_t i = b; i < e; ++i)
                for (int w = 0; w < NW; ++w) x.yz[(size_t)(i - b) * NW + w] = yz[(size_t)w * K + i];
        }
        build_sliced_host(hg, n_qubits, NW, kSlicedMaxBlob, sh);
        for (int c = 0; c < 6; ++c) make_sliced_tiles(sh, kSlicedCap[c], stiles[c]);
        t->nn = sh.nn;
        t->stream_bytes = sh.stream.size();
    }
    KoHost kh;
    const bool have_ko = NW == 1 && n_qubits >= 5 && n_qubits <= 26 && K > 0;
    if (have_ko) {
        std::vector<HostGroup> hg((size_t)G);
        for (int64_t g = 0; g < G; ++g) {
            HostGroup& x = hg[(size_t)g];
            for (int w = 0; w < 4; ++w) x.u[w] = w < NW ? gxy[(size_t)w * G + g] : 0u;
            const uint32_t b = gstart[(size_t)g], e = gstart[(size_t)g + 1];
            x.yz.assign(yz.begin() + b, yz.begin() + e);
            x.c.assign(coeff.begin() + b, coeff.begin() + e);
        }
        build_ko_host(hg, n_qubits, kKoMaxBlobWords, kh);
        t->ko_n_hi = kh.n_hi; t->ko_r_total_pad = kh.r_total_pad; t->ko_stream_bytes = kh.stream.size();
        t->ko_units = kh.units;
    }
    t->coeff_f32_exact = true;
    for (int64_t k = 0; k < K; ++k) t->coeff_f32_exact = t->coeff_f32_exact && ((double)(float)h_coeff[k] == h_coeff[k] || h_coeff[k] != h_coeff[k]);
    // environment switches (A/B measurements) are read once, here — never on the launch path
    const char* algo_env = getenv("NAQS_ELOC_ALGO");
    t->algo = (algo_env && std::string(algo_env) == "direct") ? 1 : 0;
    t->env_no_keyorder = getenv("NAQS_ELOC_NO_KEYORDER") != nullptr;
    t->env_no_dense32 = getenv("NAQS_ELOC_NO_DENSE32") != nullptr;
    t->env_no_filter = getenv("NAQS_ELOC_NO_FILTER") != nullptr;
    t->ko_disabled = getenv("NAQS_ELOC_NO_KO3") != nullptr;
    t->env_no_bin = getenv("NAQS_ELOC_NO_BIN") != nullptr;
    t->env_static_tasks = getenv("NAQS_ELOC_STATIC_TASKS") != nullptr;
    if (const char* e = getenv("NAQS_ELOC_CHUNKS")) t->env_chunks = atoi(e);
    t->env_no_graph = getenv("NAQS_ELOC_NO_GRAPH") != nullptr;

    int rc = NAQS_OK;
    auto upload = [&](void** dptr, const void* src, size_t bytes) -> int {
        NAQS_CUDA(cudaMalloc(dptr, std::max<size_t>(bytes, 16)));
        if (bytes) NAQS_CUDA(cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice));
        return NAQS_OK;
    };
    if ((rc = upload((void**)&t->d_yz, yz.data(), (size_t)NW * K * 4)) ||
        (rc = upload((void**)&t->d_coeff, coeff.data(), (size_t)K * 8)) ||
        (rc = upload((void**)&t->d_gxy, gxy.data(), (size_t)NW * G * 4)) ||
        (rc = upload((void**)&t->d_gstart, gstart.data(), (size_t)(G + 1) * 4)) ||
        (rc = upload((void**)&t->d_tiles, tiles.data(), tiles.size() * sizeof(Tile))) ||
        (rc = upload((void**)&t->d_row_tiles, row_tiles.data(), row_tiles.size() * sizeof(Tile))) ||
        (rc = upload((void**)&t->d_stream, sh.stream.data(), sh.stream.size())) ||
        (rc = upload(&t->d_stiles[0], stiles[0].data(), stiles[0].size() * sizeof(STile))) ||
        (rc = upload(&t->d_stiles[1], stiles[1].data(), stiles[1].size() * sizeof(STile))) ||
        (rc = upload(&t->d_stiles[2], stiles[2].data(), stiles[2].size() * sizeof(STile))) ||
        (rc = upload(&t->d_stiles[3], stiles[3].data(), stiles[3].size() * sizeof(STile))) ||
        (rc = upload(&t->d_stiles[4], stiles[4].data(), stiles[4].size() * sizeof(STile))) ||
        (rc = upload(&t->d_stiles[5], stiles[5].data(), stiles[5].size() * sizeof(STile))) ||
        (have_ko && ((rc = upload((void**)&t->d_ko_stream, kh.stream.data(), kh.stream.size())) ||
                     (rc = upload((void**)&t->d_ko_ht, kh.ht.data(), kh.ht.size() * 4))))) {
        naqs_table_destroy(t);
        return rc;
    }
    for (int c = 0; c < 6; ++c) {
        t->n_stiles[c] = (int)stiles[c].size();
        // work estimate per tile in (state, group) pairs: 8 / 5 groups per A / B record, a big-group word ~ 1.5; a blob ends in
        // one un-batched lookup per state: cheap against the dense table or a shared-memory filter, two dependent global
        // reads (filter word, bucket) in the hash shapes without one
        const size_t rc = rec_bytes_C(sh.nn);
        const uint32_t blob_cost = c < 3 ? 2u : (kSlicedFilter[c] ? 4u : 12u);
        for (const STile& tl : stiles[c]) {
            uint32_t cost = tl.kind == kSecA ? 8u * tl.count + 1u : tl.kind == kSecB ? 5u * tl.count + 1u
                                             : (uint32_t)(blob_cost * tl.count + 3u * ((tl.bytes - kBlobHeader * tl.count) / rc) / 2u + 1u);
            t->stile_cost[c].push_back(cost);
        }
    }
    if (cudaStreamCreateWithFlags(&t->own_stream, cudaStreamNonBlocking) != cudaSuccess) t->own_stream = nullptr;
    if (cudaMalloc((void**)&t->d_flags, 4 * sizeof(int)) != cudaSuccess || cudaMemset(t->d_flags, 0, 4 * sizeof(int)) != cudaSuccess) {
        set_error("naqs_table_create: cudaMalloc of the flag word failed");
        naqs_table_destroy(t);
        return NAQS_ERR_ALLOC;
    }
    *out = t;
    return NAQS_OK;
}

int naqs_table_destroy(naqs_table_t* t) {
    if (!t) return NAQS_OK;
    DeviceGuard guard(t->device);
    cudaFree(t->d_yz); cudaFree(t->d_coeff); cudaFree(t->d_gxy); cudaFree(t->d_gstart);
    cudaFree(t->d_dense); cudaFree(t->d_dense32_raw); cudaFree(t->d_slots); cudaFree(t->d_buckets); cudaFree(t->d_filter); cudaFree(t->d_ws); cudaFree(t->d_stage);
    if (t->h_pinned) cudaFreeHost(t->h_pinned);
    for (auto& g : t->host_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (t->own_stream) cudaStreamDestroy(t->own_stream);
    cudaFree(t->d_tiles); cudaFree(t->d_row_tiles); cudaFree(t->d_binom); cudaFree(t->d_rank_tab); cudaFree(t->d_row_cc); cudaFree(t->d_stream); cudaFree(t->d_partial);
    for (int c = 0; c < 6; ++c) cudaFree(t->d_stiles[c]);
    cudaFree(t->d_ko_stream); cudaFree(t->d_ko_ht); cudaFree(t->d_flags); cudaFree(t->d_perm);
    if (t->h_flags) cudaFreeHost(t->h_flags);
    delete t;
    return NAQS_OK;
}

int naqs_table_info(const naqs_table_t* t, int64_t* info) {
    NAQS_REQUIRE(t && info, NAQS_ERR_ARG, "naqs_table_info: NULL argument");
    info[0] = t->K; info[1] = t->G; info[2] = t->Kyz; info[3] = t->words;
    info[4] = t->n_qubits; info[5] = t->n_alpha; info[6] = t->n_beta; info[7] = t->device;
    return NAQS_OK;
}

int naqs_lookup_build(naqs_table_t* t, const uint64_t* d_keys, const void* d_psi, int psi_dtype, int64_t n, int kind,
                      void* stream_) {
    NAQS_REQUIRE(t, NAQS_ERR_ARG, "naqs_lookup_build: NULL table");
    NAQS_REQUIRE(n >= 0 && (n == 0 || (d_keys && d_psi)), NAQS_ERR_ARG, "naqs_lookup_build: NULL keys/psi");
    NAQS_REQUIRE(psi_dtype == NAQS_C128 || psi_dtype == NAQS_C64, NAQS_ERR_DTYPE, "naqs_lookup_build: psi must be complex64 or complex128");
    DeviceGuard guard(t->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc_s = stream_handover(t, stream)) return rc_s;
    const bool dup_equal = (kind & NAQS_LOOKUP_DUPLICATES_EQUAL) != 0;
    const bool assume_unique = (kind & NAQS_LOOKUP_ASSUME_UNIQUE) != 0 || dup_equal;  // plain stores are exact in both cases
    kind &= 0xff;
    if (kind == NAQS_LOOKUP_AUTO) kind = (t->n_qubits <= 22) ? NAQS_LOOKUP_DENSE : NAQS_LOOKUP_HASH;
    t->dense32_valid = false;
    t->d_dense32_ext = nullptr;
    NAQS_REQUIRE(kind == NAQS_LOOKUP_DENSE || kind == NAQS_LOOKUP_HASH, NAQS_ERR_ARG, "naqs_lookup_build: bad kind");
    const int blocks = (int)((n + 255) / 256);
    t->filter_valid = false;
    int* const range_flags = t->quiet_range_flag ? nullptr : t->d_flags;
    int rc_f = NAQS_OK;
    // 2^15 words (the size that fits shared memory; >= 4 bits per key) while n <= 2^18;
    // larger batches get >= 16 bits per key, at most 2^22 words (16 MB, L2-resident) — and, up to 2.5 * 2^18 keys, a 2^15-word
    // companion for shared memory that still rejects more than half of the misses before anything is queued
    auto build_filter = [&](int wide) -> int {
        if (n <= 0 || t->env_no_filter) return NAQS_OK;
        int log2w = kFilterLog2WordsSmem;
        if (4 * n > (32ll << kFilterLog2WordsSmem))
            while (log2w < kFilterLog2WordsMax && (32ll << log2w) < 16 * n) ++log2w;
        const bool small = log2w > kFilterLog2WordsSmem && 2 * n <= (5ll << 18);
        if (t->filter_alloc_log2w < log2w) {
            cudaFree(t->d_filter); t->d_filter = nullptr; t->filter_alloc_log2w = -1;
            NAQS_CUDA(cudaMalloc((void**)&t->d_filter, ((size_t)4 << log2w) + kFilterBytes));
            t->filter_alloc_log2w = log2w;
        }
        t->filter_log2w = log2w;
        t->filter_small_valid = small;
        uint32_t* big = t->d_filter + (small ? (1u << kFilterLog2WordsSmem) : 0u);
        NAQS_CUDA(cudaMemsetAsync(t->d_filter, 0, ((size_t)4 << log2w) + (small ? kFilterBytes : 0), stream));
        (void)wide;
        filter_build_kernel<<<blocks, 256, 0, stream>>>(big, (4u << log2w) - 4u, small ? t->d_filter : nullptr, d_keys, t->words, n, t->sector, range_flags);
        NAQS_LAUNCHED();
        t->filter_valid = true;
        return NAQS_OK;
    };
    if (kind == NAQS_LOOKUP_DENSE) {
        NAQS_REQUIRE(t->n_qubits <= 30, NAQS_ERR_ARG, "naqs_lookup_build: dense lookup needs n_qubits <= 30");
        const int64_t entries = 1ll << t->n_qubits;
        // key-order walk + unique complex64 amplitudes: an 8-byte-per-entry table suffices (exact: the kernel widens
        // float -> double, as sparse_math.pyx:33-37 does); otherwise the complex128 table with duplicate summation
        const bool use32 = assume_unique && psi_dtype == NAQS_C64 && t->algo == 0 && n >= entries / 8 && t->n_qubits <= 26 &&
                           !t->env_no_keyorder && !t->env_no_dense32;
        if (use32) {
            if (t->dense32_entries < entries) {
                // the kernel forms entry addresses as (base ^ key * 8) ^ (u * 8) (emit_batch32): base aligned to the table size
                cudaFree(t->d_dense32_raw); t->d_dense32_raw = nullptr; t->d_dense32 = nullptr; t->dense32_entries = 0;
                const size_t bytes = (size_t)entries * sizeof(float2);
                NAQS_CUDA(cudaMalloc(&t->d_dense32_raw, 2 * bytes));
                t->d_dense32 = reinterpret_cast<float2*>(((uintptr_t)t->d_dense32_raw + bytes - 1) & ~(uintptr_t)(bytes - 1));
                t->dense32_entries = entries;
            }
            NAQS_CUDA(cudaMemsetAsync(t->d_dense32, 0, (size_t)entries * sizeof(float2), stream));
            if (n > 0) {
                dense_scatter32_kernel<<<blocks, 256, 0, stream>>>(t->d_dense32, d_keys, reinterpret_cast<const float2*>(d_psi), n, t->sector, range_flags);
                NAQS_LAUNCHED();
            }
            t->dense32_valid = true;
        } else {
            if (t->dense_entries < entries) {  // the 16-byte table is only allocated when it is the one in use
                cudaFree(t->d_dense); t->d_dense = nullptr; t->dense_entries = 0;
                NAQS_CUDA(cudaMalloc((void**)&t->d_dense, (size_t)entries * sizeof(double2)));
                t->dense_entries = entries;
            }
            NAQS_CUDA(cudaMemsetAsync(t->d_dense, 0, (size_t)entries * sizeof(double2), stream));
            if (n > 0) {
                dense_scatter_kernel<<<blocks, 256, 0, stream>>>(t->d_dense, d_keys, d_psi, psi_dtype, n, dup_equal ? 1 : 0, t->sector, range_flags);
                NAQS_LAUNCHED();
            }
        }
    } else if (t->nw32 <= 2) {
        // bucketed table: 4 slots per 128 B bucket; load factor <= 0.25 while it stays well inside L2, else <= 0.5
        int64_t nb = 256;
        while (4 * nb < 4 * n) nb <<= 1;
        if (nb * (int64_t)sizeof(HashBucket) > (64ll << 20) && 4 * nb >= 4 * n) nb >>= 1;
        NAQS_REQUIRE(nb <= (1ll << 30), NAQS_ERR_ARG, "naqs_lookup_build: too many keys for the hash lookup");
        if (t->bucket_alloc < nb) {
            cudaFree(t->d_buckets); t->d_buckets = nullptr; t->bucket_alloc = 0;
            NAQS_CUDA(cudaMalloc((void**)&t->d_buckets, (size_t)nb * sizeof(HashBucket)));
            t->bucket_alloc = nb;
        }
        t->n_buckets = nb;
        bucket_init_kernel<<<(unsigned)((4 * nb + 255) / 256), 256, 0, stream>>>(t->d_buckets, nb);
        NAQS_LAUNCHED();
        if (n > 0) {
            const LookupView lv = t->lookup();
            bucket_insert_kernel<<<blocks, 256, 0, stream>>>(t->d_buckets, lv.bmask, lv.bshift, d_keys, t->words, d_psi, psi_dtype, n, dup_equal ? 1 : 0, t->sector, range_flags);
            NAQS_LAUNCHED();
        }
        if ((rc_f = build_filter(0)) != NAQS_OK) return rc_f;
    } else {
        // load factor <= 0.25 while the table stays well inside L2 (32 B slots), else <= 0.5
        int64_t cap = 1024;
        while (cap < 4 * n) cap <<= 1;
        if (cap * (int64_t)sizeof(HashSlot) > (48ll << 20) && cap >= 4 * n) cap >>= 1;
        NAQS_REQUIRE(cap <= (1ll << 31), NAQS_ERR_ARG, "naqs_lookup_build: too many keys for the hash lookup (max 2^30)");
        if (t->hash_alloc < cap) {
            cudaFree(t->d_slots); t->d_slots = nullptr; t->hash_alloc = 0;
            NAQS_CUDA(cudaMalloc((void**)&t->d_slots, (size_t)cap * sizeof(HashSlot)));
            t->hash_alloc = cap;
        }
        t->hash_cap = cap;
        hash_init_kernel<<<(int)((cap + 255) / 256), 256, 0, stream>>>(t->d_slots, cap);
        NAQS_LAUNCHED();
        if (n > 0) {
            hash_insert_kernel<<<blocks, 256, 0, stream>>>(t->d_slots, (unsigned long long)(cap - 1), t->lookup().shift, d_keys, t->words,
                                                           d_psi, psi_dtype, n, t->sector, range_flags);
            NAQS_LAUNCHED();
        }
        if ((rc_f = build_filter(1)) != NAQS_OK) return rc_f;
    }
    t->lookup_kind = kind;
    t->lookup_n = n;
    return NAQS_OK;
}

}  // extern "C"

template <int NW>
static int launch_eloc(naqs_table_t* t, const uint64_t* d_states, const void* d_psi, int psi_dtype, int64_t M,
                       double* d_eloc, cudaStream_t stream) {
    constexpr int R = 4;
    const int64_t per_block = (int64_t)kThreads * R;
    const int64_t blocks = (M + per_block - 1) / per_block;
    // small batches: split the table into chunks on group boundaries (the list of the stored-row kernels) until about one
    // wave of threads is resident; partial sums are added in chunk order (deterministic)
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, t->device);
    int n_chunks = 1;
    while (t->n_row_chunks > 0 && n_chunks < t->n_row_chunks && (int64_t)n_chunks * blocks * kThreads < (int64_t)sm_count * 2048) n_chunks *= 2;
    ChunkBounds cb;
    const Tile* tiles = t->d_tiles;
    int cap = t->tile_cap;
    if (n_chunks > 1) {
        const int merge = t->n_row_chunks / n_chunks;
        for (int c = 0; c <= kMaxChunks; ++c) cb.lo[c] = t->row_chunk_lo[std::min(c * merge, t->n_row_chunks)];
        tiles = t->d_row_tiles;
        cap = t->row_tile_cap;
        const size_t need = (size_t)n_chunks * M * sizeof(double2);
        if (t->partial_bytes < need) {
            cudaFree(t->d_partial); t->d_partial = nullptr; t->partial_bytes = 0;
            NAQS_CUDA(cudaMalloc((void**)&t->d_partial, need));
            t->partial_bytes = need;
        }
    } else {
        cb.lo[0] = 0;
        for (int c = 1; c <= kMaxChunks; ++c) cb.lo[c] = t->n_tiles;
    }
    const size_t smem = tile_smem_bytes<NW>(cap);
    auto kern = eloc_direct_kernel<NW, R, kThreads>;
    NAQS_SMEM_ATTR(kern, smem, t->device);
    kern<<<dim3((unsigned)blocks, (unsigned)n_chunks), kThreads, smem, stream>>>(t->view(), tiles, cb, n_chunks, cap, t->sector, t->lookup(), d_states,
                                                                               d_psi, psi_dtype, M, reinterpret_cast<double2*>(d_eloc), t->d_partial);
    NAQS_LAUNCHED();
    if (n_chunks > 1) {
        eloc_finalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(t->d_partial, n_chunks, d_psi, psi_dtype, M, reinterpret_cast<double2*>(d_eloc));
        NAQS_LAUNCHED();
    }
    return NAQS_OK;
}

template <int NW, int NN, int CFG, int LK, bool SEC, bool KEYORDER, bool PSI32 = false>
static int launch_sliced_cfg(naqs_table_t* t, const uint64_t* d_states, const void* d_psi, int psi_dtype, int64_t M_rows,
                             double* d_eloc, cudaStream_t stream, int n_chunks, int sm_count) {
    // key-order mode walks all 2^N keys; otherwise one thread per row
    const int64_t M = KEYORDER ? (1ll << t->n_qubits) : M_rows;
    constexpr int TL = CFG + (LK == kLookHash ? 3 : 0);  // launch shape / tile list
    constexpr int THREADS = kSlicedThreads[TL];
    const int n_tiles = t->n_stiles[TL];
    // table chunks (grid.y): contiguous tile ranges of about equal estimated work — with at most one wave of CTAs the launch
    // lasts as long as its heaviest chunk, and A tiles (8 groups per 1.4 KB) are far heavier per byte than big-group tiles
    n_chunks = std::max(1, std::min(std::min(n_chunks, kMaxChunks), std::max(n_tiles, 1)));
    ChunkBounds cb;
    {
        const std::vector<uint32_t>& cost = t->stile_cost[TL];
        uint64_t total = 0, run = 0;
        for (uint32_t c : cost) total += c;
        int c = 0;
        cb.lo[0] = 0;
        for (int i = 0; i < n_tiles; ++i) {
            // close chunk c before tile i once it has its share, keeping enough tiles for the chunks that follow
            if (c + 1 < n_chunks && i > cb.lo[c] && (run * n_chunks >= total * (uint64_t)(c + 1) || n_tiles - i <= n_chunks - 1 - c)) cb.lo[++c] = i;
            run += cost[(size_t)i];
        }
        while (c < n_chunks) cb.lo[++c] = n_tiles;
        for (int k = n_chunks + 1; k <= kMaxChunks; ++k) cb.lo[k] = n_tiles;
    }
    int max_chunk_tiles = 0;
    for (int c = 0; c < n_chunks; ++c) max_chunk_tiles = std::max(max_chunk_tiles, cb.lo[c + 1] - cb.lo[c]);
    const size_t cap = kSlicedCap[TL];
    const bool resident = max_chunk_tiles <= 1;
    const int queue_cap = kSlicedFilter[TL] ? kQueueCapFilter : kQueueCap;
    const size_t queue_bytes = LK == kLookHash ? (size_t)queue_cap * 8 * THREADS : 0;
    const size_t queue_offset = resident ? cap : 2 * cap;
    // the Bloom filter is copied to shared memory only in the 1-CTA-per-SM shape (128 KB) and only at its smallest size;
    // otherwise the kernel consults it in global memory (L2) before a bucket / slot probe
    const bool use_filter = kSlicedFilter[TL] && t->filter_valid && (t->filter_log2w == kFilterLog2WordsSmem || t->filter_small_valid);
    const size_t filter_offset = queue_offset + queue_bytes;
    const size_t smem = filter_offset + (use_filter ? kFilterBytes : 0);
    auto kern = eloc_sliced_kernel<NW, NN, THREADS, kSlicedCtasPerSm[TL], LK, SEC, KEYORDER, PSI32>;
    NAQS_SMEM_ATTR(kern, 2 * cap + queue_bytes + (kSlicedFilter[TL] ? kFilterBytes : 0), t->device);
    LookupView lv = t->lookup();
    lv.filter_in_smem = use_filter ? 1 : 0;
    double2* partial = nullptr;
    uint32_t* need_bits = nullptr;
    if (n_chunks > 1 || KEYORDER) {
        const size_t bitmap_bytes = KEYORDER ? ((((size_t)M + 31) / 32 * 4 + 255) & ~(size_t)255) : 0;  // whole 32-bit words, >= 256 B
        const size_t need = (size_t)n_chunks * M * sizeof(double2) + bitmap_bytes;
        if (t->partial_bytes < need) {
            cudaFree(t->d_partial); t->d_partial = nullptr; t->partial_bytes = 0;
            NAQS_CUDA(cudaMalloc((void**)&t->d_partial, need));
            t->partial_bytes = need;
        }
        partial = t->d_partial;
        if (KEYORDER) {
            need_bits = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(t->d_partial) + (size_t)n_chunks * M * sizeof(double2));
            NAQS_CUDA(cudaMemsetAsync(need_bits, 0, bitmap_bytes, stream));
            mark_keys_kernel<<<(unsigned)((M_rows + 255) / 256), 256, 0, stream>>>(d_states, M_rows, need_bits, M, t->d_flags);
            NAQS_LAUNCHED();
        }
    }
    // hash lookup with the filter in shared memory: walk the rows in bank-binned order (conflict-free filter probes, see
    // bin_states_kernel); two standard deviations of slack per bank, what does not fit goes to the overflow region
    BinView bin{nullptr, nullptr, 0};
    int64_t n_pos = M;
    if (LK == kLookHash && use_filter && !KEYORDER && !t->env_no_bin && M >= 4096 && M < (1ll << 30)) {
        const int64_t R = M / 32 + 2 * (int64_t)std::sqrt((double)M / 32.0) + 4;
        const size_t need = (size_t)(64 + 32 * R + M) * sizeof(int32_t);
        if (t->perm_bytes < need) {
            cudaFree(t->d_perm); t->d_perm = nullptr; t->perm_bytes = 0;
            NAQS_CUDA(cudaMalloc((void**)&t->d_perm, need));
            t->perm_bytes = need;
        }
        NAQS_CUDA(cudaMemsetAsync(t->d_perm, 0, 64 * sizeof(int32_t), stream));
        NAQS_CUDA(cudaMemsetAsync(t->d_perm + 64, 0xff, (size_t)(32 * R + M) * sizeof(int32_t), stream));
        bin_states_kernel<NW><<<(unsigned)((M + 1023) / 1024), 1024, 0, stream>>>(d_states, M, t->d_perm + 64, t->d_perm, R);
        NAQS_LAUNCHED();
        bin = BinView{t->d_perm + 64, t->d_perm, 32 * R};
        n_pos = 32 * R + M;  // upper bound for the grid; the kernel reads the true count from the overflow counter
    }
    const int64_t n_blocks = (n_pos + THREADS - 1) / THREADS;
    const int slots = sm_count * kSlicedCtasPerSm[TL];
    // tasks = (chunk, state block); the hash walk deals them dynamically when a CTA gets more than one (uneven task lengths)
    const int64_t n_tasks = n_blocks * n_chunks;
    int* task_counter = nullptr;
    if (LK == kLookHash && n_tasks > slots && !t->env_static_tasks) {
        task_counter = t->d_flags + 1;
        NAQS_CUDA(cudaMemsetAsync(task_counter, 0, sizeof(int), stream));
    }
    dim3 grid((unsigned)std::min<int64_t>(n_tasks, slots), 1u);
    SlicedView sv{t->d_stream, (const STile*)t->d_stiles[TL], n_tiles, t->nn};
    kern<<<grid, THREADS, smem, stream>>>(sv, cb, (uint32_t)cap, (uint32_t)queue_offset, (uint32_t)queue_cap, (uint32_t)filter_offset, t->sector, lv, d_states, need_bits, d_psi, psi_dtype,
                                          M, bin, n_chunks, task_counter, reinterpret_cast<double2*>(d_eloc), partial);
    NAQS_LAUNCHED();
    if (KEYORDER) {
        eloc_rows_finalize_kernel<<<(unsigned)((M_rows + 255) / 256), 256, 0, stream>>>(partial, n_chunks, M, d_states, d_psi, psi_dtype,
                                                                                      M_rows, reinterpret_cast<double2*>(d_eloc));
        NAQS_LAUNCHED();
    } else if (n_chunks > 1) {
        eloc_finalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(partial, n_chunks, d_psi, psi_dtype, M,
                                                                            reinterpret_cast<double2*>(d_eloc));
        NAQS_LAUNCHED();
    }
    return NAQS_OK;
}

// ------------------------------------------------------------------------------------------ key-order kernel (keyorder.cuh)
constexpr int kKoThreads[3] = {1024, 512, 256};
constexpr int kKoCtasPerSm[3] = {1, 2, 4};
// dynamic shared memory available to one CTA when `ctas` of them share an SM (228 KB per SM, 1 KB reserved per CTA, 128 B static)
// (static = the mbarrier, padded to the 128-byte alignment of the dynamic window: cuobjdump reports SHARED:1152 = 1024 reserved + 128)
static size_t ko_smem_cap(int ctas) { return ctas == 1 ? (size_t)232448 - 128 : (size_t)233472 / ctas - 1024 - 128; }

// balanced split of the unit list into n_chunks contiguous runs by estimated cost (a run may only end after the last word of a
// blob) -> chunk descriptors; returns false when fewer runs are possible
static bool ko_split(const std::vector<KoUnit>& units, int n_chunks, KoChunks& out, size_t& max_bytes, uint32_t& max_words) {
    const int n = (int)units.size();
    uint64_t total = 0, run = 0;
    for (const auto& u : units) total += u.cost;
    std::vector<int> lo(1, 0);
    for (int i = 0; i < n; ++i) {
        const int c = (int)lo.size() - 1;
        if (c + 1 < n_chunks && i > lo[(size_t)c] && units[(size_t)i - 1].last && run * n_chunks >= total * (uint64_t)(c + 1)) lo.push_back(i);
        run += units[(size_t)i].cost;
    }
    if ((int)lo.size() != n_chunks) return false;
    lo.push_back(n);
    max_bytes = 0; max_words = 0;
    std::vector<uint32_t> off((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i) off[(size_t)i + 1] = off[(size_t)i] + units[(size_t)i].bytes;
    std::memset(&out, 0, sizeof(out));
    for (int k = 0; k < n_chunks; ++k) {
        KoChunk& c = out.c[k];
        const int a = lo[(size_t)k], b = lo[(size_t)k + 1];
        c.off = off[(size_t)a]; c.bytes = off[(size_t)b] - off[(size_t)a]; c.r0 = (uint32_t)a; c.n_words = (uint32_t)(b - a);
        c.off_b = c.off_c = c.bytes;
        for (int i = a; i < b; ++i) {
            const KoUnit& u = units[(size_t)i];
            if (u.kind == kSecA) ++c.n_a;
            else if (u.kind == kSecB) { if (!c.n_b) c.off_b = off[(size_t)i] - c.off; ++c.n_b; }
            else { if (!c.n_c) c.off_c = off[(size_t)i] - c.off; ++c.n_c; }
        }
        max_bytes = std::max<size_t>(max_bytes, c.bytes); max_words = std::max(max_words, c.n_words);
    }
    return true;
}

// Launch plan of the key-order kernel for this table (cached: it depends only on the key space and the table): the largest CTA
// shape whose grid — tasks (32 keys each) x table chunks — fills the machine to >= 85 %, every chunk resident in shared memory.
static void ko_plan(naqs_table_t* t, int sm_count) {
    KoPlan best;
    best.valid = -1;
    double best_eff = -1.0;
    const int64_t n_tasks = (1ll << t->n_qubits) / 32;
    for (int s = 0; s < 3 && t->d_ko_stream; ++s) {
        const int warps = kKoThreads[s] / 32;
        const int64_t cta_slots = (int64_t)sm_count * kKoCtasPerSm[s];
        const size_t cap = ko_smem_cap(kKoCtasPerSm[s]);
        for (int ch = 1; ch <= kMaxChunks; ++ch) {
            KoPlan p;
            size_t max_bytes; uint32_t max_words;
            if (!ko_split(t->ko_units, ch, p.chunks, max_bytes, max_words)) continue;
            const uint32_t tw_stride = (max_words + 31) / 32 * 32;
            const size_t tw_offset = (max_bytes + 127) & ~(size_t)127, smem = tw_offset + (size_t)warps * tw_stride * 4;
            if (smem > cap) continue;
            const int64_t ctas_full = (n_tasks + warps - 1) / warps;   // CTAs per chunk when every task has its own warp
            int64_t gx;
            double eff;
            if (ctas_full * ch <= cta_slots) {
                gx = ctas_full;
                eff = (double)(n_tasks * ch) / (double)(cta_slots * warps);
            } else {
                gx = std::max<int64_t>(1, cta_slots / ch);
                const int64_t waves = (n_tasks + gx * warps - 1) / (gx * warps);
                eff = (double)n_tasks / (double)(waves * gx * warps) * (double)(gx * ch) / (double)cta_slots;
            }
            eff -= 0.005 * ch;
            if (eff > best_eff + 1e-9) {
                best_eff = eff;
                best = p;
                best.valid = 1; best.shape = s; best.n_chunks = ch; best.grid_x = (int)gx;
                best.smem = (uint32_t)smem; best.tw_offset = (uint32_t)tw_offset; best.tw_stride = tw_stride;
            }
        }
        if (best_eff >= 0.85) break;
    }
    t->ko_plan = best;
}

template <int SHAPE>
static int launch_keyorder_shape(naqs_table_t* t, const KoPlan& p, const float2* dense32, const uint32_t* need, int64_t n_tasks, int64_t n_keys,
                                 double2* partial, cudaStream_t stream) {
    auto kern = eloc_keyorder_kernel<kKoThreads[SHAPE], kKoCtasPerSm[SHAPE]>;
    NAQS_SMEM_ATTR(kern, ko_smem_cap(kKoCtasPerSm[SHAPE]), t->device);
    KoView kv{t->d_ko_stream, t->d_ko_ht, t->ko_n_hi, t->ko_r_total_pad};
    kern<<<dim3((unsigned)p.grid_x, (unsigned)p.n_chunks), kKoThreads[SHAPE], p.smem, stream>>>(kv, p.chunks, p.tw_offset, p.tw_stride, dense32, need,
                                                                                                 n_tasks, n_keys, partial);
    NAQS_LAUNCHED();
    return NAQS_OK;
}

// key-order walk with the complex64 direct-address table: mark the row keys, run the kernel over all 2^N keys, finalise the rows
static int launch_keyorder(naqs_table_t* t, const uint64_t* d_states, const void* d_psi, int psi_dtype, int64_t M_rows, double* d_eloc,
                           cudaStream_t stream) {
    const KoPlan& p = t->ko_plan;
    const int64_t n_keys = 1ll << t->n_qubits, n_tasks = n_keys / 32;
    const size_t bitmap_bytes = ((size_t)n_tasks * 4 + 255) & ~(size_t)255;
    const size_t need_bytes = (size_t)p.n_chunks * n_keys * sizeof(double2) + bitmap_bytes;
    if (t->partial_bytes < need_bytes) {
        cudaFree(t->d_partial); t->d_partial = nullptr; t->partial_bytes = 0;
        NAQS_CUDA(cudaMalloc((void**)&t->d_partial, need_bytes));
        t->partial_bytes = need_bytes;
    }
    uint32_t* need_bits = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(t->d_partial) + (size_t)p.n_chunks * n_keys * sizeof(double2));
    NAQS_CUDA(cudaMemsetAsync(need_bits, 0, bitmap_bytes, stream));
    mark_keys_kernel<<<(unsigned)((M_rows + 255) / 256), 256, 0, stream>>>(d_states, M_rows, need_bits, n_keys, t->d_flags);
    NAQS_LAUNCHED();
    const float2* dense32 = t->d_dense32_ext ? t->d_dense32_ext : t->d_dense32;
    int rc;
    switch (p.shape) {
        case 0: rc = launch_keyorder_shape<0>(t, p, dense32, need_bits, n_tasks, n_keys, t->d_partial, stream); break;
        case 1: rc = launch_keyorder_shape<1>(t, p, dense32, need_bits, n_tasks, n_keys, t->d_partial, stream); break;
        default: rc = launch_keyorder_shape<2>(t, p, dense32, need_bits, n_tasks, n_keys, t->d_partial, stream); break;
    }
    if (rc) return rc;
    eloc_rows_finalize_kernel<<<(unsigned)((M_rows + 255) / 256), 256, 0, stream>>>(t->d_partial, p.n_chunks, n_keys, d_states, d_psi, psi_dtype,
                                                                                  M_rows, reinterpret_cast<double2*>(d_eloc));
    NAQS_LAUNCHED();
    return NAQS_OK;
}

template <int NW, int NN>
static int launch_sliced(naqs_table_t* t, const uint64_t* d_states, const void* d_psi, int psi_dtype, int64_t M, double* d_eloc,
                         cudaStream_t stream) {
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, t->device);
    // key-order mode: dense (direct-address) lookup and a batch that covers at least 1/8 of the key space
    const bool keyorder = NW == 1 && t->lookup_kind == NAQS_LOOKUP_DENSE && t->n_qubits <= 26 &&
                          (t->dense32_valid || (M >= (1ll << t->n_qubits) / 8 && !t->env_no_keyorder));
    const int64_t M_rows = M;
    if (keyorder) M = 1ll << t->n_qubits;  // launch shape is chosen for the number of threads that actually run
    // pick the launch shape: large CTAs when there are at least two waves of them, else smaller CTAs, and split the
    // table into chunks (grid.y) when even those cannot fill the machine
    // launch shape: the largest CTA (fewest tile switches, least table re-streaming) whose grid — state blocks x table
    // chunks (grid.y) — fills the machine to >= 85 %; otherwise the best fill found
    const int tl_off = t->lookup_kind == NAQS_LOOKUP_HASH ? 3 : 0;
    int cfg = 2, n_chunks = 1;
    double best_eff = -1.0;
    for (int c = 0; c < 3; ++c) {
        // the hash walk's 1024-thread shape arranges the rows by filter bank (bin_states_kernel): count the positions it walks
        const bool binned = tl_off == 3 && kSlicedFilter[c + tl_off] && !t->env_no_bin && M >= 4096 && t->filter_valid &&
                            (t->filter_log2w == kFilterLog2WordsSmem || t->filter_small_valid);
        const int64_t M_walk = binned ? M + 64 * (int64_t)std::sqrt((double)M / 32.0) + 128 : M;
        const int64_t n_blocks = (M_walk + kSlicedThreads[c + tl_off] - 1) / kSlicedThreads[c + tl_off];
        const int64_t slots = (int64_t)sm_count * kSlicedCtasPerSm[c + tl_off];
        const int max_chunks = n_blocks >= slots ? 1 : std::min(t->n_stiles[c + tl_off], 16);
        int ch_best = 1;
        double eff_best = -1.0;
        for (int ch = 1; ch <= max_chunks; ++ch) {
            const int64_t ctas = n_blocks * ch, waves = (ctas + slots - 1) / slots;
            const double eff = n_blocks >= slots ? 1.0 : (double)ctas / (double)(waves * slots) - 0.005 * ch;
            if (eff > eff_best) { eff_best = eff; ch_best = ch; }
        }
        // hash walk: tasks are dealt dynamically and differ in length, so aim at >= 8 tasks per CTA slot rather than at full waves
        // (measured on Li2O 1e5: 14 chunks 0.600 ms, 7 chunks 0.635 ms, 3 chunks 0.745 ms)
        if (tl_off == 3 && n_blocks < slots && !t->env_static_tasks)
            ch_best = (int)std::min<int64_t>(std::min<int64_t>(kMaxChunks, t->n_stiles[c + tl_off]), std::max<int64_t>(ch_best, (8 * slots + n_blocks - 1) / n_blocks));
        if (eff_best > best_eff + 1e-9) { best_eff = eff_best; cfg = c; n_chunks = ch_best; }
        if (eff_best >= 0.85) { cfg = c; n_chunks = ch_best; break; }
    }
    if (t->env_chunks > 0) n_chunks = t->env_chunks;  // A/B measurements
    // lookup structures built by this library hold in-sector keys only (key_in_sector above), so a coupled state outside the
    // sector simply misses: the kernel needs its own sector test only for a caller-owned table (naqs_lookup_attach_dense32)
    const bool hash = t->lookup_kind == NAQS_LOOKUP_HASH;
    const bool secf = t->sector.enabled != 0 && t->dense32_valid && t->d_dense32_ext != nullptr;
    if constexpr (NW == 1) {
        // complex64 table, no sector test needed in the kernel: the dedicated key-order kernel (keyorder.cuh)
        if (keyorder && t->dense32_valid && !secf && t->n_qubits >= 5 && t->d_ko_stream && !t->ko_disabled) {
            if (t->ko_plan.valid == 0) ko_plan(t, sm_count);
            if (t->ko_plan.valid == 1) return launch_keyorder(t, d_states, d_psi, psi_dtype, M_rows, d_eloc, stream);
        }
        if (keyorder) {
#define NAQS_KO(CFG, SEC, P32) launch_sliced_cfg<NW, NN, CFG, kLookDense, SEC, true, P32>(t, d_states, d_psi, psi_dtype, M_rows, d_eloc, stream, n_chunks, sm_count)
            switch (cfg * 4 + (secf ? 2 : 0) + (t->dense32_valid ? 1 : 0)) {
                case 0: return NAQS_KO(0, false, false);
                case 1: return NAQS_KO(0, false, true);
                case 3: return NAQS_KO(0, true, true);
                case 4: return NAQS_KO(1, false, false);
                case 5: return NAQS_KO(1, false, true);
                case 7: return NAQS_KO(1, true, true);
                case 8: return NAQS_KO(2, false, false);
                case 9: return NAQS_KO(2, false, true);
                default: return NAQS_KO(2, true, true);
            }
#undef NAQS_KO
        }
    }
    if (NW > 1 && !hash) { set_error("naqs_eloc: dense lookup needs n_qubits <= 30"); return NAQS_ERR_STATE; }
#define NAQS_SL(CFG, LK) launch_sliced_cfg<NW, NN, CFG, (NW > 1 ? kLookHash : LK), false, false>(t, d_states, d_psi, psi_dtype, M, d_eloc, stream, n_chunks, sm_count)
    switch (cfg * 2 + (hash ? 1 : 0)) {
        case 0: return NAQS_SL(0, kLookDense);
        case 1: return NAQS_SL(0, kLookHash);
        case 2: return NAQS_SL(1, kLookDense);
        case 3: return NAQS_SL(1, kLookHash);
        case 4: return NAQS_SL(2, kLookDense);
        default: return NAQS_SL(2, kLookHash);
    }
#undef NAQS_SL
}

extern "C" {

int naqs_table_set_precision(naqs_table_t* t, int bits) {
    NAQS_REQUIRE(t && (bits == 32 || bits == 64), NAQS_ERR_DTYPE, "naqs_table_set_precision: bits must be 32 or 64 (long double has no device type)");
    NAQS_REQUIRE(bits == 64 || t->coeff_f32_exact, NAQS_ERR_ARG,
                 "naqs_table_set_precision: float32 accumulation needs coefficients already rounded to float32 (couplings.astype(np.float32), hamiltonian.py:424)");
    t->f32 = bits == 32;
    if (t->f32) t->algo = 1;  // the sliced LUTs are built from float64 partial sums
    return NAQS_OK;
}

int naqs_table_set_algo(naqs_table_t* t, int algo) {
    NAQS_REQUIRE(t && (algo == 0 || algo == 1), NAQS_ERR_ARG, "naqs_table_set_algo: algo must be 0 (sliced) or 1 (direct)");
    NAQS_REQUIRE(algo == 1 || !t->f32, NAQS_ERR_STATE, "naqs_table_set_algo: a float32 table only has the direct formulation");
    t->algo = algo;
    return NAQS_OK;
}

int naqs_eloc(naqs_table_t* t, const uint64_t* d_states, const void* d_psi, int psi_dtype, int64_t M, double* d_eloc,
              void* stream_) {
    NAQS_REQUIRE(t, NAQS_ERR_ARG, "naqs_eloc: NULL table");
    NAQS_REQUIRE(M >= 0 && (M == 0 || (d_states && d_psi && d_eloc)), NAQS_ERR_ARG, "naqs_eloc: NULL buffers");
    NAQS_REQUIRE(psi_dtype == NAQS_C128 || psi_dtype == NAQS_C64, NAQS_ERR_DTYPE, "naqs_eloc: psi must be complex64 or complex128");
    NAQS_REQUIRE(t->lookup_kind != 0, NAQS_ERR_STATE, "naqs_eloc: call naqs_lookup_build first");
    NAQS_REQUIRE(M < (1ll << 40), NAQS_ERR_ARG, "naqs_eloc: batch too large");
    if (M == 0) return NAQS_OK;
    DeviceGuard guard(t->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc_s = stream_handover(t, stream)) return rc_s;
    if (t->algo == 0) {
        switch (t->nn) {
            case 5: return launch_sliced<1, 5>(t, d_states, d_psi, psi_dtype, M, d_eloc, stream);
            case 8: return launch_sliced<1, 8>(t, d_states, d_psi, psi_dtype, M, d_eloc, stream);
            case 16: return launch_sliced<2, 16>(t, d_states, d_psi, psi_dtype, M, d_eloc, stream);
            default: return launch_sliced<4, 32>(t, d_states, d_psi, psi_dtype, M, d_eloc, stream);
        }
    }
    switch (t->nw32) {
        case 1: return launch_eloc<1>(t, d_states, d_psi, psi_dtype, M, d_eloc, stream);
        case 2: return launch_eloc<2>(t, d_states, d_psi, psi_dtype, M, d_eloc, stream);
        default: return launch_eloc<4>(t, d_states, d_psi, psi_dtype, M, d_eloc, stream);
    }
}

int naqs_dense32_scatter(float* d_table, int64_t entries, const uint64_t* d_keys, const void* d_psi, int64_t n, void* stream) {
    NAQS_REQUIRE(n >= 0 && (n == 0 || (d_table && d_keys && d_psi)), NAQS_ERR_ARG, "naqs_dense32_scatter: NULL buffers");
    NAQS_REQUIRE(entries > 0 && (entries & (entries - 1)) == 0 && entries <= (1ll << 30), NAQS_ERR_ARG, "naqs_dense32_scatter: entries must be a power of two <= 2^30");
    if (n == 0) return NAQS_OK;
    Sector none;  // no table handle here: the fused kernel keeps its own sector test for caller-owned tables
    std::memset(&none, 0, sizeof(none));
    while ((1ll << none.n_qubits) < entries) ++none.n_qubits;  // keys >= entries are skipped (never used as an address)
    dense_scatter32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(d_table), d_keys,
                                                                                        reinterpret_cast<const float2*>(d_psi), n, none, nullptr);
    NAQS_LAUNCHED();
    return NAQS_OK;
}

int naqs_lookup_attach_dense32(naqs_table_t* t, const float* d_table, int64_t entries) {
    NAQS_REQUIRE(t && d_table, NAQS_ERR_ARG, "naqs_lookup_attach_dense32: NULL argument");
    NAQS_REQUIRE(t->nw32 == 1 && t->n_qubits <= 26 && entries == (1ll << t->n_qubits), NAQS_ERR_ARG,
                 "naqs_lookup_attach_dense32: the table must have exactly 2^n_qubits entries (n_qubits <= 26)");
    NAQS_REQUIRE(t->algo == 0, NAQS_ERR_STATE, "naqs_lookup_attach_dense32: needs the sliced formulation");
    NAQS_REQUIRE(((uintptr_t)d_table & ((uintptr_t)entries * sizeof(float2) - 1)) == 0, NAQS_ERR_ARG,
                 "naqs_lookup_attach_dense32: the table must be aligned to its size (8 * 2^n_qubits bytes)");
    t->lookup_kind = NAQS_LOOKUP_DENSE;
    t->lookup_n = entries;
    t->d_dense32_ext = reinterpret_cast<const float2*>(d_table);
    t->dense32_valid = true;
    t->filter_valid = false;
    return NAQS_OK;
}

int naqs_apply_h(naqs_table_t* t, const uint64_t* d_states, int64_t M, double* d_out, void* stream_) {
    NAQS_REQUIRE(t, NAQS_ERR_ARG, "naqs_apply_h: NULL table");
    NAQS_REQUIRE(M >= 0 && (M == 0 || (d_states && d_out)), NAQS_ERR_ARG, "naqs_apply_h: NULL buffers");
    NAQS_REQUIRE(t->lookup_kind != 0, NAQS_ERR_STATE, "naqs_apply_h: call naqs_lookup_build first (it holds the vector)");
    if (M == 0) return NAQS_OK;
    // same kernels as naqs_eloc; a NULL psi makes the finalisation write the raw row sums
    DeviceGuard guard(t->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc_s = stream_handover(t, stream)) return rc_s;
    if (t->algo == 0) {
        switch (t->nn) {
            case 5: return launch_sliced<1, 5>(t, d_states, nullptr, NAQS_C128, M, d_out, stream);
            case 8: return launch_sliced<1, 8>(t, d_states, nullptr, NAQS_C128, M, d_out, stream);
            case 16: return launch_sliced<2, 16>(t, d_states, nullptr, NAQS_C128, M, d_out, stream);
            default: return launch_sliced<4, 32>(t, d_states, nullptr, NAQS_C128, M, d_out, stream);
        }
    }
    switch (t->nw32) {
        case 1: return launch_eloc<1>(t, d_states, nullptr, NAQS_C128, M, d_out, stream);
        case 2: return launch_eloc<2>(t, d_states, nullptr, NAQS_C128, M, d_out, stream);
        default: return launch_eloc<4>(t, d_states, nullptr, NAQS_C128, M, d_out, stream);
    }
}

int naqs_eloc_host(naqs_table_t* t, const void* h_states, int key_itemsize, const void* h_psi, int psi_dtype, int64_t M,
                   const void* h_tkeys, const void* h_tpsi, int64_t T, int lookup_kind, void* h_eloc, int eloc_dtype) {
    int rc = naqs_eloc_host_begin(t, h_states, key_itemsize, h_psi, psi_dtype, M, h_tkeys, h_tpsi, T, lookup_kind, h_eloc, eloc_dtype);
    return rc ? rc : naqs_eloc_host_end(t);
}

int naqs_eloc_host_begin(naqs_table_t* t, const void* h_states, int key_itemsize, const void* h_psi, int psi_dtype, int64_t M,
                         const void* h_tkeys, const void* h_tpsi, int64_t T, int lookup_kind, void* h_eloc, int eloc_dtype) {
    NAQS_REQUIRE(t, NAQS_ERR_ARG, "naqs_eloc_host: NULL table");
    NAQS_REQUIRE(M >= 0 && (M == 0 || (h_states && h_psi && h_eloc)), NAQS_ERR_ARG, "naqs_eloc_host: NULL buffers");
    NAQS_REQUIRE(psi_dtype == NAQS_C128 || psi_dtype == NAQS_C64, NAQS_ERR_DTYPE, "naqs_eloc_host: psi must be complex64 or complex128");
    NAQS_REQUIRE(eloc_dtype == NAQS_C128 || eloc_dtype == NAQS_C64, NAQS_ERR_DTYPE, "naqs_eloc_host: E_loc must be complex64 or complex128");
    NAQS_REQUIRE(key_itemsize == 8 || ((key_itemsize == 4 || key_itemsize == 2) && t->words == 1), NAQS_ERR_DTYPE,
                 "naqs_eloc_host: keys must be 64-bit words, or int16/int32 indices for single-word keys (hilbert.py:405-410)");
    if (M == 0) return NAQS_OK;
    DeviceGuard guard(t->device);
    const size_t psz = psi_dtype == NAQS_C64 ? 8 : 16;
    const size_t kb_in = key_itemsize == 8 ? (size_t)8 * t->words : (size_t)key_itemsize, kb = (size_t)8 * t->words;
    const bool own_table = h_tkeys != nullptr;
    if (!own_table) T = M;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    // staging layout: raw keys | keys64 | psi | [table raw keys | table keys64 | table psi] | eloc128 | eloc64
    const size_t o_kraw = 0, o_k64 = al(M * kb_in), o_psi = o_k64 + al(M * kb), o_tkraw = o_psi + al(M * psz),
                 o_tk64 = o_tkraw + (own_table ? al(T * kb_in) : 0), o_tp = o_tk64 + (own_table ? al(T * kb) : 0),
                 o_out = o_tp + (own_table ? al(T * psz) : 0), o_out32 = o_out + al((size_t)M * 16), total = o_out32 + al((size_t)M * 8);
    if (t->stage_bytes < total) {
        cudaFree(t->d_stage); t->d_stage = nullptr; t->stage_bytes = 0;
        NAQS_CUDA(cudaMalloc(&t->d_stage, total));
        t->stage_bytes = total;
    }
    if (!t->h_flags) NAQS_CUDA(cudaHostAlloc((void**)&t->h_flags, sizeof(int), cudaHostAllocDefault));
    char* d = (char*)t->d_stage;
    cudaStream_t st = t->own_stream;
    const size_t out_bytes = (size_t)M * (eloc_dtype == NAQS_C64 ? 8 : 16);
    // small batch: stage through page-locked memory so that the stream sequence has fixed addresses and can be replayed as a graph
    const size_t in_bytes = M * (kb_in + psz) + (own_table ? T * (kb_in + psz) : 0);
    const bool small = !t->env_no_graph && st != nullptr && in_bytes + out_bytes <= ((size_t)1 << 19);
    t->pending_out_dst = nullptr;
    if (small) {
        const size_t p_k = 0, p_p = al(M * kb_in), p_tk = p_p + al(M * psz), p_tp = p_tk + (own_table ? al(T * kb_in) : 0),
                     p_out = p_tp + (own_table ? al(T * psz) : 0), p_total = p_out + al(out_bytes);
        if (t->pinned_bytes < p_total) {
            if (t->h_pinned) { NAQS_CUDA(cudaStreamSynchronize(st)); cudaFreeHost(t->h_pinned); }
            t->h_pinned = nullptr; t->pinned_bytes = 0;
            NAQS_CUDA(cudaHostAlloc(&t->h_pinned, std::max<size_t>(p_total, (size_t)1 << 16), cudaHostAllocDefault));
            t->pinned_bytes = std::max<size_t>(p_total, (size_t)1 << 16);
        }
        char* hp = (char*)t->h_pinned;
        std::memcpy(hp + p_k, h_states, M * kb_in);
        std::memcpy(hp + p_p, h_psi, M * psz);
        h_states = hp + p_k; h_psi = hp + p_p;
        if (own_table) {
            std::memcpy(hp + p_tk, h_tkeys, T * kb_in);
            std::memcpy(hp + p_tp, h_tpsi, T * psz);
            h_tkeys = hp + p_tk; h_tpsi = hp + p_tp;
        }
        t->pending_out_src = hp + p_out; t->pending_out_dst = h_eloc; t->pending_out_bytes = out_bytes;
        h_eloc = hp + p_out;
    }
    auto upload_keys = [&](const void* h, size_t o_raw, size_t o_64, int64_t n, const uint64_t** out) -> int {
        if (key_itemsize == 8) {
            NAQS_CUDA(cudaMemcpyAsync(d + o_64, h, n * kb, cudaMemcpyHostToDevice, st));
        } else {
            NAQS_CUDA(cudaMemcpyAsync(d + o_raw, h, n * kb_in, cudaMemcpyHostToDevice, st));
            widen_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d + o_raw, key_itemsize, n, (uint64_t*)(d + o_64));
            NAQS_LAUNCHED();
        }
        *out = (const uint64_t*)(d + o_64);
        return NAQS_OK;
    };
    auto enqueue = [&]() -> int {
        const uint64_t *d_keys = nullptr, *d_tk = nullptr;
        int rc = upload_keys(h_states, o_kraw, o_k64, M, &d_keys);
        if (rc) return rc;
        NAQS_CUDA(cudaMemcpyAsync(d + o_psi, h_psi, M * psz, cudaMemcpyHostToDevice, st));
        const void* d_tp = d + o_psi;
        d_tk = d_keys;
        if (own_table) {
            rc = upload_keys(h_tkeys, o_tkraw, o_tk64, T, &d_tk);
            if (rc) return rc;
            NAQS_CUDA(cudaMemcpyAsync(d + o_tp, h_tpsi, T * psz, cudaMemcpyHostToDevice, st));
            d_tp = d + o_tp;
        }
        rc = naqs_lookup_build(t, d_tk, d_tp, psi_dtype, T, lookup_kind, st);
        if (rc) return rc;
        rc = naqs_eloc(t, d_keys, d + o_psi, psi_dtype, M, (double*)(d + o_out), st);
        if (rc) return rc;
        if (eloc_dtype == NAQS_C64) {  // the reference hands E_loc to torch as float32 pairs (src/utils/complex.py:139-140)
            narrow_eloc_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>((const double2*)(d + o_out), M, (float2*)(d + o_out32));
            NAQS_LAUNCHED();
            NAQS_CUDA(cudaMemcpyAsync(h_eloc, d + o_out32, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
        } else {
            NAQS_CUDA(cudaMemcpyAsync(h_eloc, d + o_out, (size_t)M * 16, cudaMemcpyDeviceToHost, st));
        }
        NAQS_CUDA(cudaMemcpyAsync(t->h_flags, t->d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        return NAQS_OK;
    };
    if (!small) return enqueue();

    // ---- small batch: replay / capture the sequence as a CUDA graph -------------------------------------------------
    using HostGraph = naqs_table::HostGraph;
    auto lookup_state = [&](int64_t* s) {
        s[0] = t->lookup_kind; s[1] = t->lookup_n; s[2] = t->dense32_valid; s[3] = (int64_t)(uintptr_t)t->d_dense32_ext; s[4] = t->filter_valid;
        s[5] = t->filter_log2w; s[6] = t->filter_small_valid; s[7] = t->n_buckets; s[8] = t->hash_cap;
    };
    auto buffer_ptrs = [&](const void** q) {
        q[0] = t->d_stage; q[1] = t->d_ws; q[2] = t->d_partial; q[3] = t->d_dense; q[4] = t->d_dense32; q[5] = t->d_buckets; q[6] = t->d_slots;
        q[7] = t->d_filter; q[8] = t->d_perm; q[9] = t->h_pinned; q[10] = t->d_flags; q[11] = t->h_flags;
    };
    const int64_t sig[10] = {M, T, key_itemsize, psi_dtype, eloc_dtype, lookup_kind, own_table ? 1 : 0, t->algo, t->f32, 0};
    HostGraph* g = nullptr;
    for (auto& e : t->host_graphs)
        if (std::memcmp(e.sig, sig, sizeof(sig)) == 0) { g = &e; break; }
    const void* now[12];
    buffer_ptrs(now);
    if (g && g->exec && std::memcmp(g->ptrs, now, sizeof(now)) == 0) {
        if (int rc_s = stream_handover(t, st)) return rc_s;
        t->lookup_kind = (int)g->lookup_state[0]; t->lookup_n = g->lookup_state[1]; t->dense32_valid = g->lookup_state[2] != 0;
        t->d_dense32_ext = reinterpret_cast<const float2*>((uintptr_t)g->lookup_state[3]); t->filter_valid = g->lookup_state[4] != 0;
        t->filter_log2w = (int)g->lookup_state[5]; t->filter_small_valid = g->lookup_state[6] != 0; t->n_buckets = g->lookup_state[7]; t->hash_cap = g->lookup_state[8];
        NAQS_CUDA(cudaGraphLaunch(g->exec, st));
        g_launches.fetch_add(g->n_launches, std::memory_order_relaxed);
        g->last_use = ++t->host_graph_clock;
        return NAQS_OK;
    }
    if (!g) {  // first call with this signature: plain run (it allocates whatever the sequence needs), remember the signature
        if (t->host_graphs.size() >= 8) {
            size_t victim = 0;
            for (size_t i = 1; i < t->host_graphs.size(); ++i) if (t->host_graphs[i].last_use < t->host_graphs[victim].last_use) victim = i;
            if (t->host_graphs[victim].exec) { NAQS_CUDA(cudaStreamSynchronize(st)); cudaGraphExecDestroy(t->host_graphs[victim].exec); }
            t->host_graphs.erase(t->host_graphs.begin() + (long)victim);
        }
        HostGraph e;
        std::memcpy(e.sig, sig, sizeof(sig));
        e.last_use = ++t->host_graph_clock;
        t->host_graphs.push_back(e);
        return enqueue();
    }
    g->last_use = ++t->host_graph_clock;
    if (g->failed) return enqueue();
    if (g->exec) { NAQS_CUDA(cudaStreamSynchronize(st)); cudaGraphExecDestroy(g->exec); g->exec = nullptr; }  // a buffer moved: capture again
    // second call: capture.  Nothing allocates now (same sizes as the first call); should anything go wrong the plain path runs.
    if (int rc_s = stream_handover(t, st)) return rc_s;
    const int64_t launches_before = g_launches.load();
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
        const int rc = enqueue();
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        ok = rc == NAQS_OK && ce == cudaSuccess && graph != nullptr;
    }
    if (ok) ok = cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    const void* after[12];
    buffer_ptrs(after);
    if (ok && std::memcmp(after, now, sizeof(now)) != 0) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; ok = false; }
    if (!ok) {
        cudaGetLastError();
        g->exec = nullptr; g->failed = true;
        g_launches.store(launches_before);
        return enqueue();
    }
    g->n_launches = g_launches.load() - launches_before;
    std::memcpy(g->ptrs, now, sizeof(now));
    lookup_state(g->lookup_state);
    NAQS_CUDA(cudaGraphLaunch(g->exec, st));
    return NAQS_OK;
}

int naqs_eloc_host_end(naqs_table_t* t) {
    NAQS_REQUIRE(t, NAQS_ERR_ARG, "naqs_eloc_host_end: NULL table");
    if (!t->h_flags) return NAQS_OK;  // nothing was begun (or an empty batch)
    DeviceGuard guard(t->device);
    cudaStream_t st = t->own_stream;
    NAQS_CUDA(cudaStreamSynchronize(st));
    if (t->pending_out_dst) {
        std::memcpy(t->pending_out_dst, t->pending_out_src, t->pending_out_bytes);
        t->pending_out_dst = nullptr;
    }
    if (*t->h_flags & 1) {
        *t->h_flags = 0;
        NAQS_CUDA(cudaMemsetAsync(t->d_flags, 0, sizeof(int), st));
        set_error("naqs_eloc_host: a state index lies outside [0, 2^n_qubits) (the reference raises IndexError); its row is NaN");
        return NAQS_ERR_INDEX;
    }
    return NAQS_OK;
}

int naqs_table_check(naqs_table_t* t, void* stream_) {
    NAQS_REQUIRE(t, NAQS_ERR_ARG, "naqs_table_check: NULL table");
    DeviceGuard guard(t->device);
    cudaStream_t st = (cudaStream_t)stream_;
    int h_flags = 0;
    NAQS_CUDA(cudaMemcpyAsync(&h_flags, t->d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    NAQS_CUDA(cudaStreamSynchronize(st));
    if (h_flags & 1) {
        NAQS_CUDA(cudaMemsetAsync(t->d_flags, 0, sizeof(int), st));
        set_error("naqs_table_check: a key outside [0, 2^n_qubits) was passed since the last check (the reference raises IndexError); "
                  "it was ignored as a table key and its row is NaN");
        return NAQS_ERR_INDEX;
    }
    return NAQS_OK;
}

}  // extern "C"
