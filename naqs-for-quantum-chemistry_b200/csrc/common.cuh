// common.cuh — shared declarations of libnaqs_eloc (sm_100a only).
//
// Internal layout of the device-resident Pauli table:
//   terms are stored GROUP-MAJOR: groups = unique XY masks in ascending order (np.unique order of
//   the reference, src/optimizer/hamiltonian.py:248), terms of a group in ascending reference index k,
//   so a serial walk over a group reproduces the summation order of
//   src_cpp/hamiltonian_math.pyx:31-34 bit for bit.
//   Masks are held as NW32 32-bit words (NW32 = 1 for N<=32, 2 for N<=63, 4 for N<=127) in
//   struct-of-arrays form so that a warp reads one word per term with a single broadcast LDS.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/naqs_eloc.h"

namespace naqs {

void set_error(const std::string& msg);
extern std::atomic<int64_t> g_launches;

#define NAQS_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            naqs::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                \
            return e_ == cudaErrorMemoryAllocation ? NAQS_ERR_ALLOC : NAQS_ERR_CUDA;            \
        }                                                                                       \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize sticks to (function, device): raise it once per launch site and device, not per call
#define NAQS_SMEM_ATTR(kern, bytes, device)                                                     \
    do {                                                                                        \
        static int attr_set_[64];                                                               \
        const int b_ = (int)(bytes);                                                            \
        if (attr_set_[(device) & 63] < b_) {                                                    \
            NAQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, b_)); \
            attr_set_[(device) & 63] = b_;                                                      \
        }                                                                                       \
    } while (0)

#define NAQS_REQUIRE(cond, code, msg)                                                           \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            naqs::set_error(msg);                                                               \
            return code;                                                                        \
        }                                                                                       \
    } while (0)

// count + check a kernel launch
#define NAQS_LAUNCHED()                                                                         \
    do {                                                                                        \
        naqs::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
        NAQS_CUDA(cudaGetLastError());                                                          \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Sector description passed by value to kernels.  enabled == 0 => every key is "in sector".
struct Sector {
    uint32_t even[4];
    uint32_t odd[4];
    int n_alpha, n_beta, enabled, n_qubits;
};

// Group-major term table view (device pointers).
struct TableView {
    const uint32_t* yz;      // [NW32][K]   SoA: word w of term i at yz[w*K + i]
    const double* coeff;     // [K]
    const uint32_t* gxy;     // [NW32][G]
    const uint32_t* gstart;  // [G+1]       term offsets of each group
    int K, G;
    int f32;                 // 1: every partial sum of H_ij is rounded to float32 (the reference's dtype=np.float32 kernel)
};

// Table chunks of a sliced launch (grid.y): chunk c walks tiles [lo[c], lo[c + 1]) — contiguous ranges of about equal WORK.
constexpr int kMaxChunks = 16;
struct ChunkBounds {
    int lo[kMaxChunks + 1];
};

// Key-order kernel (keyorder.cuh): one unit per parity word of its stream, chunk descriptors passed by value, cached launch plan
struct KoUnit {
    uint32_t kind;      // kSecA / kSecB / kSecC
    uint32_t bytes;
    uint32_t cost;      // estimated work (group-equivalents), for balanced chunks
    uint32_t last;      // C: last word of its blob (a chunk may only end after one)
};
struct KoChunk {
    uint32_t off, bytes;       // byte range in the stream (multiples of 16)
    uint32_t n_a, n_b, n_c;    // A records, B records, C words
    uint32_t off_b, off_c;     // byte offsets of the first B record / C word relative to `off`
    uint32_t r0, n_words;      // first parity word (index into HT) and their number (= n_a + n_b + n_c)
};
struct KoChunks {
    KoChunk c[kMaxChunks];
};
struct KoPlan {
    int valid = 0;           // 0 = not planned yet, 1 = usable, -1 = the key-order kernel cannot run this table
    int shape = 0;           // 0: 1024 x 1 CTA/SM, 1: 512 x 2, 2: 256 x 4
    int n_chunks = 1, grid_x = 1;
    uint32_t smem = 0, tw_offset = 0, tw_stride = 0;
    KoChunks chunks{};
};

// A shared-memory tile of the term table: terms [t0, t1) and the groups touching them [g0, g1).
// A group may straddle two tiles; its partial sum is carried in registers.
struct Tile {
    uint32_t t0, t1;
    uint32_t g0, g1;
};

// Hash slot: one 32-byte sector per probe.
struct alignas(32) HashSlot {
    unsigned long long key[2];
    double re, im;
};
static_assert(sizeof(HashSlot) == 32, "HashSlot must be one 32 B sector");
constexpr unsigned long long kEmptyKey = ~0ull;

// Bucketed hash for keys of <= 63 bits (all molecules): one probe = one 32-byte sector holding the 4 keys of a bucket.
// A bucket that ever overflowed carries a flag in bit 63 of key[0]; without it a lookup ends after ONE sector read
// whether it hits or misses (misses dominate a sparse VMC batch), so there is no dependent probe chain.
struct alignas(128) HashBucket {
    unsigned long long key[4];  // kEmptyKey63 when free; bit 63 of key[0] = "bucket overflowed into the next one"
    double2 psi[4];
    unsigned long long pad[4];
};
static_assert(sizeof(HashBucket) == 128, "HashBucket must be one 128 B line");
constexpr unsigned long long kKeyMask63 = 0x7FFFFFFFFFFFFFFFull;
constexpr unsigned long long kEmptyKey63 = kKeyMask63;
constexpr unsigned long long kOverflowFlag = 0x8000000000000000ull;

struct LookupView {
    const double2* dense;    // [2^N] (kind DENSE)
    const HashSlot* slots;   // [cap]  (kind HASH)
    unsigned long long mask; // cap - 1
    int kind;
    int shift;               // 32 - log2(cap): slot = hash32 >> shift
    const HashBucket* buckets;  // [n_buckets] (kind HASH, keys <= 63 bits)
    unsigned bmask;          // n_buckets - 1
    int bshift;              // 32 - log2(n_buckets)
    const uint32_t* filter;  // blocked Bloom filter over the table keys (filter_mask / 4 + 1 words) or nullptr
    const float2* dense32;   // [2^N] complex64 copy of the dense table (unique keys + complex64 psi only) or nullptr
    uint32_t filter_mask;    // byte offset of a key's filter word = lin-hash word & filter_mask (= 4 * n_words - 4)
    const uint32_t* filter_small;  // 2^14-word companion of a larger filter (same hashes, the low 14 word bits) or nullptr
    int filter_in_smem;      // the launch copies the filter into shared memory (it has 2^14 words and the shape has room)
    int* flags;              // bit 0: a key outside [0, 2^n_qubits) was seen (naqs_table_check reports and clears it)
};



__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
// Multiplicative (Fibonacci) hashing: the top bits of key * odd-constant depend on every key bit.  Two independent
// products give the bucket index / filter word and the two bit positions inside that word — 2 IMAD + a few SHF per key
// of <= 32 bits.  Table capacities are powers of two <= 2^31.
__host__ __device__ inline uint32_t hash32(unsigned long long k0, unsigned long long k1) {
    return (uint32_t)k0 * 0x9E3779B1u + (uint32_t)(k0 >> 32) * 0x85EBCA6Bu + (uint32_t)k1 * 0xC2B2AE35u +
           (uint32_t)(k1 >> 32) * 0x27D4EB2Fu;
}
__host__ __device__ inline unsigned long long hash_slot(unsigned long long k0, unsigned long long k1, int shift) {
    return (unsigned long long)(hash32(k0, k1) >> shift);
}

// Bloom filter of the hash lookups.  In a sparse VMC batch almost every coupled state is NOT in the table; the filter
// answers those without the global sector read of a bucket / slot.  Blocked: both bits of a key live in ONE 32-bit word
// (one load per test).  2^15 words (128 KB, >= 4 bits per key) up to 2^18 keys — that size is copied into shared memory by
// every CTA of the 1024-thread launch shape and consulted for every (state, group) pair (about 2 % false positives at
// 1e5 keys: three bits per key); larger batches get >= 16 bits per key, up to 2^22 words (16 MB, L2-resident),
// consulted from global memory before a bucket / slot probe.
//
// The hashes are GF(2)-LINEAR in the key (XOR of one pseudo-random 64-bit column per set key bit): hash(s ^ u) =
// hash(s) ^ hash(u), so the fused kernel keeps hash(s) per thread, reads hash(u) with the group's flip mask (warp-uniform)
// and pays two XORs per coupled state instead of two multiplies and shifts.  A random linear map is a universal hash
// family: the false-positive rate is that of an ideal hash in expectation.  Second benefit: the shared-memory BANK of the
// filter word is bits 2..6 of the word hash, so bank(s ^ u) = bank(s) ^ bank(u) — a warp whose 32 states have pairwise
// different bank(s) probes 32 different banks for EVERY group (bin_states_kernel arranges the batch that way).
//   hw: byte offset of the filter word (bits 0-1 clear; a filter of 2^L words uses hw & (4 * 2^L - 4))
//   hb: low 5 bits = position b of the key's first bit; the other two sit 11 and 21 positions further (mod 32):
//       set  word |= rotl(kFilterPattern, b),   test  (~rotr(word, b) & kFilterPattern) == 0
constexpr int kFilterLog2WordsSmem = 15, kFilterLog2WordsMax = 22;
constexpr uint32_t kFilterBytes = (1u << kFilterLog2WordsSmem) * 4;  // the shared-memory copy
constexpr uint32_t kFilterPattern = 0x00200801u;  // bits 0, 11, 21
__host__ __device__ inline void lin_hash_word(uint32_t x, int word_index, uint32_t& hw, uint32_t& hb) {
    while (x) {
#ifdef __CUDA_ARCH__
        const int i = __ffs((int)x) - 1;
#else
        const int i = __builtin_ctz(x);
#endif
        x &= x - 1;
        const unsigned long long m = mix64(0x9E3779B97F4A7C15ull * (unsigned long long)(word_index * 32 + i + 1));
        hw ^= (uint32_t)m & ~3u;
        hb ^= (uint32_t)(m >> 32);
    }
}
__host__ __device__ inline void lin_hash_key64(const uint64_t* key, int words, uint32_t& hw, uint32_t& hb) {
    hw = 0; hb = 0;
    for (int w = 0; w < words; ++w) {
        lin_hash_word((uint32_t)key[w], 2 * w, hw, hb);
        lin_hash_word((uint32_t)(key[w] >> 32), 2 * w + 1, hw, hb);
    }
}
__host__ __device__ inline uint32_t filter_insert_bits(uint32_t hb) {
    const uint32_t b = hb & 31u;
    return (kFilterPattern << b) | (kFilterPattern >> ((32u - b) & 31u));
}
__host__ __device__ inline bool filter_test_word(uint32_t word, uint32_t hb) {
    const uint32_t b = hb & 31u;
    const uint32_t r = (word >> b) | (word << ((32u - b) & 31u));
    return (~r & kFilterPattern) == 0u;
}

}  // namespace naqs

struct naqs_table {
    int device = 0, words = 1, nw32 = 1, n_qubits = 0, n_alpha = -1, n_beta = -1;
    int64_t K = 0, G = 0, Kyz = 0;
    naqs::Sector sector{};
    // device arrays
    uint32_t* d_yz = nullptr;
    double* d_coeff = nullptr;
    uint32_t* d_gxy = nullptr;
    uint32_t* d_gstart = nullptr;
    naqs::Tile* d_tiles = nullptr;
    int n_tiles = 0, tile_cap = 0;
    long long* d_binom = nullptr;  // C(n, k) table for the restricted-index ranker (lazy)
    int32_t* d_rank_tab = nullptr; // per-spin rank tables (<= 32 qubits with a sector): [2^n_even | 2^n_odd] int32, built with d_binom
    long long rank_n_b = 0;        // C(n_odd, n_beta)
    // stored-row kernels: the same terms cut into kMaxChunks chunks on group boundaries (grid.y of rows_kernel), so that a small
    // batch (a VMC sector: 10^4 rows) still fills the machine; row_chunk_lo[c] = first tile of base chunk c
    naqs::Tile* d_row_tiles = nullptr;
    int row_chunk_lo[naqs::kMaxChunks + 1] = {0};
    int n_row_chunks = 0, row_tile_cap = 0;
    // per-(chunk, row) counts of the last chunked naqs_rows_count: a naqs_rows_fill for the same (d_states, M, stream) that follows
    // it reuses them once instead of recounting (its indptr must stem from those counts anyway)
    int32_t* d_row_cc = nullptr;
    size_t row_cc_bytes = 0;
    const void* row_cc_states = nullptr;
    int64_t row_cc_M = 0;
    int row_cc_chunks = 0;
    cudaStream_t row_cc_stream = nullptr;
    bool row_cc_valid = false;
    // sliced (v2) formulation: byte stream + tile lists for the 1024/512/256-thread launch shapes
    int algo = 0;                  // 0 = sliced (default), 1 = direct
    int f32 = 0;                   // naqs_table_set_precision(32): float32 accumulation of H_ij (direct formulation only)
    bool coeff_f32_exact = false;  // every coefficient is representable in float32
    unsigned char* d_stream = nullptr;
    size_t stream_bytes = 0;
    void* d_stiles[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [0..2] dense, [3..5] hash tile lists
    int n_stiles[6] = {0, 0, 0, 0, 0, 0};
    std::vector<uint32_t> stile_cost[6];   // host: estimated work per tile (group-equivalents), for balanced table chunks
    int nn = 0;
    // key-order (v3) stream: single-word keys, n_qubits in [5, 26] (keyorder.cuh); the host copy of its unit list drives the chunking
    unsigned char* d_ko_stream = nullptr;
    size_t ko_stream_bytes = 0;
    uint32_t* d_ko_ht = nullptr;
    int ko_n_hi = 0, ko_r_total_pad = 0;
    std::vector<naqs::KoUnit> ko_units;
    bool env_no_keyorder = false, env_no_dense32 = false, env_no_filter = false;  // A/B switches, read once at table creation
    bool ko_disabled = false;          // NAQS_ELOC_NO_KO3 (A/B against the generic key-order walk of sliced.cuh)
    naqs::KoPlan ko_plan{};            // cached launch plan (the key space, hence the shape, is fixed per table)
    double2* d_partial = nullptr;
    size_t partial_bytes = 0;
    // lookup
    int lookup_kind = 0;  // 0 = none
    int64_t lookup_n = 0;
    double2* d_dense = nullptr;
    int64_t dense_entries = 0;
    const float2* d_dense32_ext = nullptr;  // caller-owned complex64 dense table (naqs_lookup_attach_dense32), e.g. all-reduced
    float2* d_dense32 = nullptr;   // complex64 dense table (key-order walk with unique complex64 amplitudes), aligned to its size
    void* d_dense32_raw = nullptr; // the allocation d_dense32 points into
    int64_t dense32_entries = 0;
    bool dense32_valid = false;
    naqs::HashSlot* d_slots = nullptr;     // 128-bit keys: 32 B slots, linear probing
    int64_t hash_cap = 0, hash_alloc = 0;
    naqs::HashBucket* d_buckets = nullptr; // <= 63-bit keys: 128 B buckets of 4
    int64_t n_buckets = 0, bucket_alloc = 0;
    uint32_t* d_filter = nullptr;          // Bloom filter storage (2^filter_log2w words, allocated for filter_alloc_log2w)
    bool filter_valid = false;
    int filter_log2w = naqs::kFilterLog2WordsSmem, filter_alloc_log2w = -1;
    bool filter_small_valid = false;       // d_filter = [2^14-word companion][2^filter_log2w words] when set
    // generic workspace (scan / sort temporaries, host-path staging)
    void* d_ws = nullptr;
    size_t ws_bytes = 0;
    void* d_stage = nullptr;  // device staging for naqs_eloc_host
    size_t stage_bytes = 0;
    void* h_pinned = nullptr;   // page-locked staging of the small-batch (CUDA graph) form of naqs_eloc_host
    size_t pinned_bytes = 0;
    // small batches through naqs_eloc_host are launch-bound (ten stream operations for a 224-state LiH batch): the whole sequence —
    // uploads from the page-locked staging, lookup build, fused kernel, download — is captured once per call signature into a
    // CUDA graph and replayed.  An entry is valid only while every buffer it baked in is still the current one.
    struct HostGraph {
        int64_t sig[10] = {0};       // M, T, key_itemsize, psi_dtype, eloc_dtype, lookup_kind (with flags), own_table, algo, f32, 0
        const void* ptrs[12] = {nullptr};
        cudaGraphExec_t exec = nullptr;
        bool failed = false;         // capture did not work for this signature: use the plain path
        int64_t n_launches = 0;      // kernels in the graph (naqs_launch_count stays meaningful)
        int64_t lookup_state[9] = {0};
        unsigned long long last_use = 0;
    };
    std::vector<HostGraph> host_graphs;
    unsigned long long host_graph_clock = 0;
    bool env_no_graph = false;
    const void* pending_out_src = nullptr;  // staged E_loc of a begun small-batch call -> copied to pending_out_dst in _end
    void* pending_out_dst = nullptr;
    size_t pending_out_bytes = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t last_stream = nullptr;  // stream of the last call that touched the per-table buffers (stream_handover)
    bool last_stream_valid = false;
    int* h_flags = nullptr;   // page-locked copy target of d_flags[0] for the host-buffer entry
    int* d_flags = nullptr;   // [4] device flags (bit 0 of [0]: key out of range), see naqs_table_check
    int32_t* d_perm = nullptr;      // bank-binned order of a hash-lookup batch (bin_states_kernel) + 33 counters in front
    size_t perm_bytes = 0;
    bool quiet_range_flag = false;  // naqs_table_exchange pads shards with out-of-range keys on purpose
    bool env_no_bin = false, env_static_tasks = false;
    int env_chunks = 0;

    naqs::TableView view() const { return naqs::TableView{d_yz, d_coeff, d_gxy, d_gstart, (int)K, (int)G, f32}; }
    naqs::LookupView lookup() const {
        int shift = 32;
        for (int64_t c = hash_cap; c > 1; c >>= 1) --shift;
        int bshift = 32;
        for (int64_t c = n_buckets; c > 1; c >>= 1) --bshift;
        return naqs::LookupView{d_dense, d_slots, (unsigned long long)(hash_cap - 1), lookup_kind, shift,
                                d_buckets, (unsigned)(n_buckets > 0 ? n_buckets - 1 : 0), bshift,
                                filter_valid ? d_filter + (filter_small_valid ? (1u << naqs::kFilterLog2WordsSmem) : 0u) : nullptr,
                                dense32_valid ? (d_dense32_ext ? d_dense32_ext : d_dense32) : nullptr,
                                (uint32_t)((4u << filter_log2w) - 4u), filter_valid && filter_small_valid ? d_filter : nullptr, 0, d_flags};
    }
};

namespace naqs {
int ensure_ws(naqs_table* t, size_t bytes);
}
