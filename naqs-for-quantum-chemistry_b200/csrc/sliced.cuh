// sliced.cuh — the "nibble-sliced parity + group LUT" formulation of the fused E_loc kernel (v2).
//
// Why: in the direct formulation every (state, term) pair costs AND + POPC + shift + XOR + DADD, and POPC
// issues at 16 lanes/clk/SM on sm_100a (measured: 4.0e12 couplings/s ceiling, profiles/pipe_peaks_*.json).
// Here the 32 parities of a WORD of terms are produced at once for one state,
//     P_w(s) = XOR_i  T[w][i][ nibble_i(s) ]            (one 4-byte LDS per 4 qubits, no POPC),
// where T[w][i][v] holds, for each of the 32 terms of word w, the parity of popcount(yz_term & (v << 4i)).
// XY groups of <= 4 (<= 6) terms occupy 4 (6) adjacent bits of a word, and the group's matrix element
//     H[s, s^u] = sum_k c_k (-1)^{p_k}     (k ascending, src_cpp/hamiltonian_math.pyx:31-34)
// is read from a 16- (64-) entry table indexed by those parity bits.  Every table entry was produced by
// the same serial fp64 additions, in the same order, as the reference performs — so H is bit-identical —
// but the device does one LDS.64 instead of n x (POPC, XOR, DADD).  Larger groups (the diagonal, the
// 2-flip groups) are cut into consecutive chunks of 6 terms with one 64-entry table each; the kernel adds the
// chunk sums in term order (one LDS.64 + DADD per 6 terms).  That re-associates the reference's serial sum at
// the chunk boundaries: H of a big group agrees with the reference to a few ulp, not bit for bit — E_loc is
// specified to 1e-12 relative; the direct formulation (rows / CSR / hij_dense, NAQS_ELOC_ALGO=direct) keeps the
// fully serial order and is bit-exact for every group.
//
// Stream layout (device memory, staged into shared memory tile by tile with cp.async.bulk + mbarrier):
//   A record (8 groups x 4 bits):  T[NN][16] u32 | U[8][NW] u32 (+ U << 3 when NW == 1) | HW[8] HB[8] u32 | LUT[8][16] f64
//   B record (5 groups x 6 bits):  T[NN][16] u32 | U[8][NW] u32 (+ U << 3 when NW == 1) | HW[8] HB[8] u32 | LUT[5][64] f64
//   C blob   (<= 150 terms of one big group; a longer group is several blobs with the same u, each contributing
//             H_blob * psi(s ^ u) on its own):  header{n_words, HW, HB, -, u[4]} | n_words x ( T[NN][16] u32 | LUT[5][64] f64 )
//   HW / HB: the GF(2)-linear Bloom-filter hashes of the flip mask u (lin_hash, common.cuh).  Linear means
//   hash(s ^ u) = hash(s) ^ hash(u): the filter test of a coupled state costs two XORs with the per-thread hash(s).
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "eloc_kernels.cuh"

namespace naqs {

enum { kSecA = 0, kSecB = 1, kSecC = 2 };

struct STile {
    uint32_t kind;      // kSecA / kSecB / kSecC
    uint32_t count;     // records (A, B) or blobs (C)
    uint32_t bytes;     // multiple of 16
    uint32_t pad;
    uint64_t offset;    // byte offset into the stream, multiple of 16
};

struct SlicedView {
    const unsigned char* stream;
    const STile* tiles;
    int n_tiles;
    int nn;             // nibbles per key (5, 8, 16 or 32)
};

// ------------------------------------------------------------------------------------------ host builder
struct HostGroup {
    uint32_t u[4];
    std::vector<uint32_t> yz;  // [n][NW]
    std::vector<double> c;
};

inline int nibbles_for(int n_qubits) { return n_qubits <= 20 ? 5 : (n_qubits <= 32 ? 8 : (n_qubits <= 63 ? 16 : 32)); }
// flip masks of a record: 8 slots of nw words; single-word keys carry a second copy shifted left by 3 (= byte offsets into
// the complex64 direct-address table, whose base is aligned to its size: entry address = (base ^ key * 8) ^ (u * 8))
inline size_t u_bytes(int nw) { return (nw == 1 ? 64 : (size_t)32 * nw) + 64; }  // masks + their filter hashes HW[8], HB[8]
inline size_t rec_bytes_A(int nn, int nw) { return (size_t)64 * nn + u_bytes(nw) + 8 * 16 * 8; }
inline size_t rec_bytes_B(int nn, int nw) { return (size_t)64 * nn + u_bytes(nw) + 5 * 64 * 8; }
inline size_t rec_bytes_C(int nn) { return (size_t)64 * nn + 5 * 64 * 8; }  // one word of a big group: 5 chunks of <= 6 terms
constexpr size_t kBlobHeader = 32;

struct SlicedHost {
    std::vector<unsigned char> stream;
    int nn = 0, nw = 0;
    size_t n_rec_a = 0, n_rec_b = 0;
    struct Blob { size_t offset, bytes; };
    std::vector<Blob> blobs;
    size_t off_a = 0, off_b = 0, off_c = 0;
};

// parity nibble tables of up to 32 terms (yz pointers may be null = padding term, parity 0)
inline void fill_nibble_tables(unsigned char* dst, int nn, int nw, const uint32_t* const* yz_of_bit, int n_bits) {
    uint32_t* T = reinterpret_cast<uint32_t*>(dst);
    for (int i = 0; i < nn; ++i)
        for (int v = 0; v < 16; ++v) {
            uint32_t word = 0;
            for (int b = 0; b < n_bits; ++b) {
                const uint32_t* yz = yz_of_bit[b];
                if (!yz) continue;
                const int w = (4 * i) / 32, sh = (4 * i) % 32;
                if (w >= nw) continue;
                const uint32_t nib = (yz[w] >> sh) & 15u;
                word |= (uint32_t)(__builtin_popcount(nib & (uint32_t)v) & 1) << b;
            }
            T[i * 16 + v] = word;
        }
}

// serial signed sum in reference order: pattern bit t set => term t enters with a minus sign
inline double lut_entry(const double* c, int n, unsigned pattern) {
    volatile double acc = 0.0;  // volatile: every addition individually rounded, no re-association
    for (int t = 0; t < n; ++t) acc = acc + (((pattern >> t) & 1u) ? -c[t] : c[t]);
    return acc;
}

inline void build_sliced_host(const std::vector<HostGroup>& groups, int n_qubits, int nw, size_t max_blob_bytes, SlicedHost& out) {
    const int nn = nibbles_for(n_qubits);
    out.nn = nn; out.nw = nw;
    std::vector<const HostGroup*> ga, gb, gc;
    for (const auto& g : groups) {
        const size_t n = g.c.size();
        (n <= 4 ? ga : (n <= 6 ? gb : gc)).push_back(&g);
    }
    const size_t ra = rec_bytes_A(nn, nw), rb = rec_bytes_B(nn, nw), rc = rec_bytes_C(nn);
    out.n_rec_a = (ga.size() + 7) / 8;
    out.n_rec_b = (gb.size() + 4) / 5;
    auto& S = out.stream;
    S.clear();
    out.off_a = 0;
    S.resize(out.n_rec_a * ra, 0);
    auto pack = [&](const std::vector<const HostGroup*>& gs, size_t n_rec, size_t rec, int per, int bits, size_t base) {
        for (size_t r = 0; r < n_rec; ++r) {
            unsigned char* p = S.data() + base + r * rec;
            const uint32_t* yz_of_bit[32] = {nullptr};
            uint32_t* U = reinterpret_cast<uint32_t*>(p + 64 * nn);
            double* L = reinterpret_cast<double*>(p + 64 * nn + u_bytes(nw));
            for (int j = 0; j < per; ++j) {
                const size_t gi = r * per + j;
                const HostGroup* g = gi < gs.size() ? gs[gi] : nullptr;
                const int n = g ? (int)g->c.size() : 0;
                for (int t = 0; t < n; ++t) yz_of_bit[j * bits + t] = g->yz.data() + (size_t)t * nw;
                for (int w = 0; w < nw; ++w) U[j * nw + w] = g ? g->u[w] : 0u;
                if (nw == 1) U[8 + j] = g ? g->u[0] << 3 : 0u;
                {
                    uint32_t* HWB = reinterpret_cast<uint32_t*>(p + 64 * nn + u_bytes(nw) - 64);
                    uint32_t hw = 0, hb = 0;
                    if (g) for (int w = 0; w < nw; ++w) lin_hash_word(g->u[w], w, hw, hb);
                    HWB[j] = hw; HWB[8 + j] = hb;
                }
                for (unsigned pat = 0; pat < (1u << bits); ++pat) L[j * (1 << bits) + pat] = g ? lut_entry(g->c.data(), n, pat) : 0.0;
            }
            fill_nibble_tables(p, nn, nw, yz_of_bit, per * bits);
        }
    };
    pack(ga, out.n_rec_a, ra, 8, 4, out.off_a);
    out.off_b = S.size();
    S.resize(S.size() + out.n_rec_b * rb, 0);
    pack(gb, out.n_rec_b, rb, 5, 6, out.off_b);
    out.off_c = S.size();
    // big groups (> 6 terms): consecutive chunks of 6 terms, each with its own 64-entry LUT of serially accumulated signed
    // sums; 5 chunks (30 terms) share a parity word; blobs of at most max_words words, each blob a self-contained
    // pseudo-group (same u).  The kernel adds the chunk sums of a blob in order, so H_ij of a big group is the reference's
    // serial sum re-associated at the chunk / blob boundaries (exact in real arithmetic, within a few ulp in floating
    // point); the direct formulation keeps the fully serial order.
    const size_t max_words = std::max<size_t>(1, (max_blob_bytes - kBlobHeader) / rc);
    for (const HostGroup* g : gc) {
        const size_t n = g->c.size(), total_words = (n + 29) / 30;
        for (size_t w0 = 0; w0 < total_words; w0 += max_words) {
            const size_t nwords = std::min(max_words, total_words - w0);
            const size_t off = S.size();
            S.resize(off + kBlobHeader + nwords * rc, 0);
            uint32_t* hdr = reinterpret_cast<uint32_t*>(S.data() + off);
            hdr[0] = (uint32_t)nwords;
            hdr[1] = hdr[2] = 0;
            for (int w = 0; w < nw; ++w) { hdr[4 + w] = g->u[w]; lin_hash_word(g->u[w], w, hdr[1], hdr[2]); }
            for (size_t q = 0; q < nwords; ++q) {
                unsigned char* p = S.data() + off + kBlobHeader + q * rc;
                const uint32_t* yz_of_bit[32] = {nullptr};
                double* L = reinterpret_cast<double*>(p + 64 * nn);
                for (int j = 0; j < 5; ++j) {
                    const size_t t0 = std::min(n, (w0 + q) * 30 + (size_t)j * 6), t1 = std::min(n, t0 + 6);
                    for (size_t t = t0; t < t1; ++t) yz_of_bit[j * 6 + (t - t0)] = g->yz.data() + t * nw;
                    for (unsigned pat = 0; pat < 64; ++pat) L[j * 64 + pat] = lut_entry(g->c.data() + t0, (int)(t1 - t0), pat);  // empty chunk: +0.0
                }
                fill_nibble_tables(p, nn, nw, yz_of_bit, 30);
            }
            out.blobs.push_back({off, kBlobHeader + nwords * rc});
        }
    }
}

// tiles for a given shared-memory buffer capacity
inline void make_sliced_tiles(const SlicedHost& h, size_t cap, std::vector<STile>& tiles) {
    tiles.clear();
    const size_t ra = rec_bytes_A(h.nn, h.nw), rb = rec_bytes_B(h.nn, h.nw);
    auto add_records = [&](uint32_t kind, size_t base, size_t n_rec, size_t rec) {
        const size_t per = std::max<size_t>(1, cap / rec);
        for (size_t r = 0; r < n_rec; r += per) {
            const size_t n = std::min(per, n_rec - r);
            tiles.push_back(STile{kind, (uint32_t)n, (uint32_t)(n * rec), 0, (uint64_t)(base + r * rec)});
        }
    };
    add_records(kSecA, h.off_a, h.n_rec_a, ra);
    add_records(kSecB, h.off_b, h.n_rec_b, rb);
    size_t i = 0;
    while (i < h.blobs.size()) {
        size_t bytes = 0, n = 0;
        while (i + n < h.blobs.size() && (n == 0 || bytes + h.blobs[i + n].bytes <= cap)) { bytes += h.blobs[i + n].bytes; ++n; }
        tiles.push_back(STile{kSecC, (uint32_t)n, (uint32_t)bytes, 0, (uint64_t)h.blobs[i].offset});
        i += n;
    }
}

// ------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int NW, int NN>
__device__ __forceinline__ void state_nibbles(const uint32_t (&s)[NW], uint32_t (&nib)[NN]) {
#pragma unroll
    for (int i = 0; i < NN; ++i) {
        const int w = (4 * i) / 32, sh = (4 * i) % 32;
        nib[i] = w < NW ? ((s[w] >> sh) & 15u) * 4u : 0u;  // byte offset inside the 64-byte nibble row
    }
}

template <int NN>
__device__ __forceinline__ uint32_t parity_word(const unsigned char* __restrict__ T, const uint32_t (&nib)[NN]) {
    uint32_t p = 0;
#pragma unroll
    for (int i = 0; i < NN; ++i) p ^= *reinterpret_cast<const uint32_t*>(T + i * 64 + nib[i]);
    return p;
}

// ------------------------------------------------------------------------------------------ kernel
constexpr int kLookDense = 0, kLookHash = 1;

// 64-bit address of entry `key` of the complex128 dense table (ptxas lowers the wide multiply-add to a LEA pair)
__device__ __forceinline__ const double2* dense_entry(const double2* __restrict__ base, uint32_t key) {
    unsigned long long addr;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(addr) : "r"(key), "l"(base));
    return reinterpret_cast<const double2*>(addr);
}

// complex64 direct-address table whose base is aligned to its size (a power of two): the byte address of entry s ^ u is
// (base ^ s * 8) ^ (u * 8) — ONE LOP3 per group from the per-thread word a0 = low32(base) ^ s * 8 and the pre-shifted flip
// mask u8 of the record; the high address word is shared.  Sector test on the shifted key with shifted masks (sec8).
template <bool SEC, int B>
__device__ __forceinline__ void emit_batch32(const double (&h)[B], const uint32_t (&u8)[B], uint32_t a0, uint32_t base_hi, bool valid,
                                             const Sector& sec8, double& e_re, double& e_im) {
    float2 q[B];
    double hh[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const uint32_t lo = a0 ^ u8[b];
        unsigned long long addr;
        asm("mov.b64 %0, {%1, %2};" : "=l"(addr) : "r"(lo), "r"(base_hi));
        q[b] = __ldg(reinterpret_cast<const float2*>(addr));
        hh[b] = h[b];
        if constexpr (SEC) {
            const uint32_t j8[1] = {lo};  // bits above the table size belong to the base and are masked off by sec8
            hh[b] = ((h[b] != 0.0) & valid && in_sector<1>(j8, sec8)) ? h[b] : 0.0;
        }
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
        e_re = __fma_rn(hh[b], (double)q[b].x, e_re);
        e_im = __fma_rn(hh[b], (double)q[b].y, e_im);
    }
}

// Accumulate H[b] * psi_table(s ^ u_b) for a batch of B groups (dense complex128 table).  All B table reads are issued before any
// is consumed (memory-level parallelism); lanes whose H is exactly 0.0 (hamiltonian.py:363) or whose coupled
// state leaves the sector (hamiltonian.py:328) read entry 0 / slot 0 instead — a single broadcast sector —
// and contribute H * psi = 0 exactly, so no branch is needed.
template <int NW, bool SEC, bool KEYORDER, int B>
__device__ __forceinline__ void emit_batch(const double (&h)[B], const uint32_t (&u)[B], const uint32_t (&s)[NW],
                                           bool valid, const Sector& sec, const LookupView& lv, double& e_re, double& e_im) {
    static_assert(NW == 1, "the dense lookup holds keys of <= 30 bits");
    double hh[B];
    double2 p[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const uint32_t j[1] = {s[0] ^ u[b]};
        if constexpr (KEYORDER && !SEC) {
            // key-order walk: the 32 lanes hold 32 consecutive keys, so s ^ u stays inside one aligned 32-entry block of
            // the table for every lane — loading it even where h == 0 costs no extra line and needs neither test nor select
            p[b] = __ldg(dense_entry(lv.dense, j[0]));
            hh[b] = h[b];
        } else {
            bool on = (h[b] != 0.0) & valid;
            if constexpr (SEC) on = on && in_sector<1>(j, sec);
            p[b] = __ldg(dense_entry(lv.dense, (KEYORDER || on) ? j[0] : 0u));
            hh[b] = SEC ? (on ? h[b] : 0.0) : h[b];  // without a sector filter "off" already means h == 0 (or an invalid lane)
        }
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
        e_re = __fma_rn(hh[b], p[b].x, e_re);
        e_im = __fma_rn(hh[b], p[b].y, e_im);
    }
}

// Bloom-filter test of a coupled state (launch shape with the filter in shared memory): a clear bit proves the key is not
// in the table — ~90 % of the couplings of a large-sector batch end here without touching global memory.  Cheap enough
// (one LDS, two IMAD, a few shifts) to run for EVERY (state, group) pair in the light path, before anything is queued.
// The same test against the filter in GLOBAL memory (larger batches, or shapes without room for the copy) runs in the
// probe rounds, where it replaces most bucket reads (HBM sectors of a table far larger than L2) by L2 hits.
template <int NW>
__device__ __forceinline__ bool filter_pass(const uint32_t (&j)[NW], const uint32_t* __restrict__ filt, int wshift) {
    unsigned long long k0, k1;
    key_words64<NW>(j, k0, k1);
    uint32_t w, b1, b2;
    filter_word_bits(k0, k1, hash32(k0, k1), wshift, w, b1, b2);
    const uint32_t word = filt[w];
    return ((word >> b1) & (word >> b2) & 1u) != 0;
}

// Hash-lookup ("heavy") epilogue of up to B couplings of one thread: optional Bloom-filter test (`filt`: the filter in
// global memory when the light path could not consult a shared-memory copy, else nullptr), probe of the bucketed table
// (keys <= 63 bits: one 256-bit load reads the four keys of a 128-byte bucket) or of the 32-byte-slot table (wider keys),
// complex multiply-add.  Called for couplings whose H is not exactly 0.0 (hamiltonian.py:363) — and, in the filter shape,
// whose coupled state passed the filter — which the per-thread queue of the kernel below makes dense across the warp.
// The first probes of all B couplings are issued before any is examined, so their L2 latencies overlap.  SEC: sector test
// on the coupled state (hamiltonian.py:328); unused by the library's own tables, which hold in-sector keys only.
template <int NW, bool SEC, int B>
__device__ __forceinline__ void heavy_lookup(int n, const double (&h)[B], const uint32_t* const (&u)[B], const uint32_t (&s)[NW],
                                             const Sector& sec, const LookupView& lv, const uint32_t* __restrict__ filt, int filt_wshift,
                                             double& e_re, double& e_im) {
    unsigned long long k0[B], k1[B];
    unsigned slot[B];
    bool on[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        on[b] = b < n;
        uint32_t j[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) j[w] = on[b] ? (s[w] ^ u[b][w]) : s[w];
        if constexpr (SEC) on[b] = on[b] & in_sector<NW>(j, sec);
        key_words64<NW>(j, k0[b], k1[b]);
        if (filt) on[b] = on[b] & filter_pass<NW>(j, filt, filt_wshift);  // a clear bit proves the key is not in the table
        if constexpr (NW <= 2) slot[b] = hash32(k0[b], 0ull) >> lv.bshift;
        else slot[b] = (unsigned)hash_slot(k0[b], k1[b], lv.shift);
    }
    if constexpr (NW <= 2) {
        // bucketed table: the 4 keys of a bucket are one sector; all B first probes are in flight together
        BucketKeys bk[B];
#pragma unroll
        for (int b = 0; b < B; ++b) bk[b] = load_bucket_keys(lv.buckets + (on[b] ? slot[b] : 0u));
#pragma unroll
        for (int b = 0; b < B; ++b) {
            if (on[b]) {
                while (true) {
                    const int hs = bucket_match(bk[b], k0[b]);
                    if (hs >= 0) {
                        const double2 p = __ldg(&lv.buckets[slot[b]].psi[hs]);
                        e_re = __fma_rn(h[b], p.x, e_re);
                        e_im = __fma_rn(h[b], p.y, e_im);
                        break;
                    }
                    if (!(bk[b].k[0] & kOverflowFlag)) break;  // bucket never overflowed: definite miss after one sector
                    slot[b] = (slot[b] + 1) & lv.bmask;
                    bk[b] = load_bucket_keys(lv.buckets + slot[b]);
                }
            }
        }
    } else {
        ulonglong2 kk[B];
#pragma unroll
        for (int b = 0; b < B; ++b) kk[b] = __ldg(reinterpret_cast<const ulonglong2*>(lv.slots + (on[b] ? slot[b] : 0u)));
#pragma unroll
        for (int b = 0; b < B; ++b) {
            if (on[b]) {
                while (true) {  // 128-bit keys: 32 B slots, linear probing
                    if (kk[b].x == k0[b] && kk[b].y == k1[b]) {
                        const double2 p = __ldg(reinterpret_cast<const double2*>(lv.slots + slot[b]) + 1);
                        e_re = __fma_rn(h[b], p.x, e_re);
                        e_im = __fma_rn(h[b], p.y, e_im);
                        break;
                    }
                    if (kk[b].x == kEmptyKey && kk[b].y == kEmptyKey) break;
                    slot[b] = (slot[b] + 1) & (unsigned)lv.mask;
                    kk[b] = __ldg(reinterpret_cast<const ulonglong2*>(lv.slots + slot[b]));
                }
            }
        }
    }
}

constexpr int kQueueCap = 16;  // pending couplings per thread (hash mode)

// One state per thread.  Each CTA owns state blocks blockIdx.x, blockIdx.x + gridDim.x, ... and walks the
// tiles [tile_lo, tile_hi) of its chunk (blockIdx.y) for each of them; with a single tile the table stays
// resident in shared memory for the CTA's lifetime.
//
// KEYORDER (dense lookup, batch dense in key space): thread m IS key m (states == nullptr, M = 2^N); `need` is a bitmap
// of the keys that occur as rows; the raw sums S[k] = sum_u H[k, k^u] psi(k^u) go to partial[chunk * M + k] and
// eloc_rows_finalize_kernel turns them into E_loc per row.  A warp then holds 32 consecutive keys and every table
// read of a group falls into one aligned 512-byte block: 4 L1 lines per request instead of ~11 scattered sectors.
template <int NW, int NN, int THREADS, int CTAS_PER_SM, int LK, bool SEC, bool KEYORDER, bool PSI32>
__global__ void __launch_bounds__(THREADS, CTAS_PER_SM)
eloc_sliced_kernel(SlicedView sv, const __grid_constant__ ChunkBounds chunks, uint32_t buf_bytes, uint32_t queue_offset, uint32_t filter_offset, Sector sec, LookupView lv,
                   const uint64_t* __restrict__ states, const uint32_t* __restrict__ need, const void* __restrict__ psi,
                   int psi_dtype, int64_t M, double2* __restrict__ out, double2* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar[2];
    constexpr int UB = NW == 1 ? 64 : 32 * NW;  // u_bytes(NW)
    constexpr int REC_A = 64 * NN + UB + 1024, REC_B = 64 * NN + UB + 2560, REC_C = 64 * NN + 2560;

    const int tile_lo = chunks.lo[blockIdx.y], tile_hi = chunks.lo[blockIdx.y + 1];
    const bool resident = (tile_hi - tile_lo) <= 1;
    if (threadIdx.x == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t* sfilt = nullptr;  // Bloom filter of the table keys in shared memory (hash lookup, 1024-thread shape, <= 2^17 keys)
    const uint32_t* gfilt = nullptr;  // ... or in global memory (L2), consulted in the probe rounds
    if constexpr (LK == kLookHash) {
        if (lv.filter && lv.filter_in_smem) {
            uint4* dst = reinterpret_cast<uint4*>(smem + filter_offset);
            const uint4* src = reinterpret_cast<const uint4*>(lv.filter_small ? lv.filter_small : lv.filter);
            for (uint32_t i = threadIdx.x; i < kFilterBytes / 16; i += THREADS) dst[i] = __ldg(src + i);
            sfilt = reinterpret_cast<const uint32_t*>(smem + filter_offset);
            if (lv.filter_small) gfilt = lv.filter;  // the full-size filter screens the survivors once more before a bucket read
        } else {
            gfilt = lv.filter;
        }
    }
    __syncthreads();
    // complex64 direct-address table (PSI32): base split into words, sector masks shifted like the keys (see emit_batch32)
    const uint32_t base_lo = (uint32_t)reinterpret_cast<unsigned long long>(lv.dense32);
    const uint32_t base_hi = (uint32_t)(reinterpret_cast<unsigned long long>(lv.dense32) >> 32);
    Sector sec8 = sec;
    sec8.even[0] <<= 3; sec8.odd[0] <<= 3;
    uint32_t phase0 = 0, phase1 = 0;
    bool have_resident = false;

    auto issue = [&](int t, int b) {  // thread 0 only
        const STile tl = sv.tiles[t];
        mbar_expect_tx(&mbar[b], tl.bytes);
        bulk_g2s(smem + (size_t)b * buf_bytes, sv.stream + tl.offset, tl.bytes, &mbar[b]);
    };

    const int64_t n_blocks = (M + THREADS - 1) / THREADS;
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const int64_t m = blk * THREADS + threadIdx.x;
        bool valid = m < M;
        uint32_t s[NW];
        if constexpr (KEYORDER) {
            s[0] = valid ? (uint32_t)m : 0u;  // threads past the key space read entry 0 (table reads are unconditional)
#pragma unroll
            for (int w = 1; w < NW; ++w) s[w] = 0;
            valid = valid && ((need[m >> 5] >> (m & 31)) & 1u);
        } else {
            if (valid) load_key<NW>(states, m, s);
            else {
#pragma unroll
                for (int w = 0; w < NW; ++w) s[w] = 0;
            }
        }
        uint32_t nib[NN];
        state_nibbles<NW, NN>(s, nib);
        const uint32_t a0 = base_lo ^ (s[0] << 3);  // PSI32 only
        double e_re = 0.0, e_im = 0.0;

        // hash mode: a coupling survives the light path when H != 0 AND (filter shape) its coupled state passes the Bloom
        // filter in shared memory — a few % of the (state, group) pairs of a large-sector batch.  Survivors are parked in a
        // per-thread queue (shared memory, [slot][thread] layout, one 32-bit word each: LUT byte offset | flip-mask offset
        // inside the current tile) and resolved in warp-wide rounds, so the expensive part (bucket probe in global memory,
        // key compare) runs with many lanes busy.  The queue is drained before a tile buffer is released.
        constexpr uint32_t QSTRIDE = THREADS * 4;  // bytes between consecutive queue slots of one thread
        unsigned char* const q0 = smem + queue_offset + threadIdx.x * 4;
        unsigned char* qtail = q0;            // the queue holds (qtail - q0) / QSTRIDE couplings
        constexpr int PB = 2;  // couplings resolved per thread and round
        auto pop_round = [&](const unsigned char* __restrict__ buf) {
            const int n = min((int)((uint32_t)(qtail - q0) / QSTRIDE), PB);
            double h[PB];
            const uint32_t* u[PB];
#pragma unroll
            for (int b = 0; b < PB; ++b) {
                // always read a valid slot (the oldest one when the queue is shorter than b + 1): no branch, lanes are masked by n
                const unsigned char* slot = b < n ? qtail - (b + 1) * QSTRIDE : q0;
                const uint32_t e = *reinterpret_cast<const uint32_t*>(slot);
                h[b] = *reinterpret_cast<const double*>(buf + (e & 0xffffu));
                u[b] = reinterpret_cast<const uint32_t*>(buf + (e >> 16) * 4u);
            }
            qtail -= n * QSTRIDE;
            heavy_lookup<NW, SEC, PB>(n, h, u, s, sec, lv, gfilt, lv.filter_wshift, e_re, e_im);  // a shared-memory filter was consulted before queueing
        };
        auto drain = [&](const unsigned char* __restrict__ buf) {  // before a tile buffer is released: its offsets die with it
            while (__any_sync(0xffffffffu, qtail != q0)) pop_round(buf);
        };
        // branch-free: the entry is always written at the tail, the tail only advances for a live coupling
        const uint32_t q_adv = valid ? QSTRIDE : 0u;  // an invalid lane never keeps an entry
        auto push = [&](double h, uint32_t entry) {
            *reinterpret_cast<uint32_t*>(qtail) = entry;
            qtail += (h != 0.0) ? q_adv : 0u;
        };
        // the same with the shared-memory Bloom filter consulted first; U = the record's flip masks, group j
        auto push_filtered = [&](double h, const uint32_t* __restrict__ U, int j, const uint4& ua, const uint4& ub, uint32_t entry) {
            *reinterpret_cast<uint32_t*>(qtail) = entry;
            uint32_t k[NW];
            if constexpr (NW == 1) {        // masks of the 8 groups preloaded as two vectors
                const uint32_t uv[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
                k[0] = s[0] ^ uv[j];
            } else if constexpr (NW == 2) {
                const uint2 v = reinterpret_cast<const uint2*>(U)[j];
                k[0] = s[0] ^ v.x; k[1] = s[1] ^ v.y;
            } else {
                const uint4 v = reinterpret_cast<const uint4*>(U)[j];
                k[0] = s[0] ^ v.x; k[1] = s[1] ^ v.y; k[2] = s[2] ^ v.z; k[3] = s[3] ^ v.w;
            }
            const bool pass = filter_pass<NW>(k, sfilt, 32 - kFilterLog2WordsSmem);  // the shared-memory copy always has 2^14 words
            qtail += (pass && h != 0.0) ? q_adv : 0u;
        };

        auto process = [&](const unsigned char* __restrict__ buf, const uint32_t tl_kind, const uint32_t tl_count) {
            // queue entry of group 0 of the first record: LUT byte offset | (flip-mask offset / 4) << 16; one add per record
            constexpr uint32_t EB_FIRST = (uint32_t)(64 * NN + UB) | ((uint32_t)(64 * NN / 4) << 16);
            constexpr uint32_t EB_STEP_A = (uint32_t)REC_A | ((uint32_t)(REC_A / 4) << 16), EB_STEP_B = (uint32_t)REC_B | ((uint32_t)(REC_B / 4) << 16);
            [[maybe_unused]] uint32_t ebase = EB_FIRST;
            if (tl_kind == kSecA) {
                const unsigned char* rec = buf;
                for (uint32_t r = 0; r < tl_count; ++r, rec += REC_A) {
                    const uint32_t P = parity_word<NN>(rec, nib);
                    const uint32_t* U = reinterpret_cast<const uint32_t*>(rec + 64 * NN);
                    const unsigned char* L = rec + 64 * NN + UB;
                    if constexpr (LK == kLookDense) {
                        const uint32_t* Ux = PSI32 ? U + 8 : U;  // complex64 table: flip masks as byte offsets (u * 8)
                        const uint4 ua = *reinterpret_cast<const uint4*>(Ux), ub = *reinterpret_cast<const uint4*>(Ux + 4);
                        const uint32_t uu[2][4] = {{ua.x, ua.y, ua.z, ua.w}, {ub.x, ub.y, ub.z, ub.w}};
#pragma unroll
                        for (int j0 = 0; j0 < 8; j0 += 4) {
                            double h[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int j = j0 + jj;
                                const uint32_t off = j == 0 ? ((P << 3) & 0x78u) : ((P >> (4 * j - 3)) & 0x78u);
                                h[jj] = *reinterpret_cast<const double*>(L + j * 128 + off);
                            }
                            if constexpr (PSI32) emit_batch32<SEC, 4>(h, uu[j0 / 4], a0, base_hi, valid, sec8, e_re, e_im);
                            else emit_batch<NW, SEC, KEYORDER, 4>(h, uu[j0 / 4], s, valid, sec, lv, e_re, e_im);
                        }
                    } else {
                        // entry = byte offset of the LUT entry | (flip-mask offset / 4) << 16, both relative to the tile buffer
                        // (tiles of the hash shapes are < 64 KB); ebase follows the record pointer
                        if (sfilt) {  // warp-uniform
                            uint4 ua = make_uint4(0, 0, 0, 0), ub = ua;
                            if constexpr (NW == 1) { ua = *reinterpret_cast<const uint4*>(U); ub = *reinterpret_cast<const uint4*>(U + 4); }
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const uint32_t off = j == 0 ? ((P << 3) & 0x78u) : ((P >> (4 * j - 3)) & 0x78u);
                                push_filtered(*reinterpret_cast<const double*>(L + j * 128 + off), U, j, ua, ub, ebase + j * (128u + ((uint32_t)NW << 16)) + off);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const uint32_t off = j == 0 ? ((P << 3) & 0x78u) : ((P >> (4 * j - 3)) & 0x78u);
                                push(*reinterpret_cast<const double*>(L + j * 128 + off), ebase + j * (128u + ((uint32_t)NW << 16)) + off);
                            }
                        }
                        ebase += EB_STEP_A;
                        while (__any_sync(0xffffffffu, qtail > q0 + (kQueueCap - 8) * QSTRIDE)) pop_round(buf);
                    }
                }
            } else if (tl_kind == kSecB) {
                const unsigned char* rec = buf;
                for (uint32_t r = 0; r < tl_count; ++r, rec += REC_B) {
                    const uint32_t P = parity_word<NN>(rec, nib);
                    const uint32_t* U = reinterpret_cast<const uint32_t*>(rec + 64 * NN);
                    const unsigned char* L = rec + 64 * NN + UB;
                    if constexpr (LK == kLookDense) {
                        const uint32_t* Ux = PSI32 ? U + 8 : U;
                        const uint4 ua = *reinterpret_cast<const uint4*>(Ux);
                        const uint32_t uu[5] = {ua.x, ua.y, ua.z, ua.w, Ux[4]};
                        double h[5];
#pragma unroll
                        for (int j = 0; j < 5; ++j) {
                            const uint32_t off = j == 0 ? ((P << 3) & 0x1f8u) : ((P >> (6 * j - 3)) & 0x1f8u);
                            h[j] = *reinterpret_cast<const double*>(L + j * 512 + off);
                        }
                        if constexpr (PSI32) emit_batch32<SEC, 5>(h, uu, a0, base_hi, valid, sec8, e_re, e_im);
                        else emit_batch<NW, SEC, KEYORDER, 5>(h, uu, s, valid, sec, lv, e_re, e_im);
                    } else {
                        if (sfilt) {  // warp-uniform
                            uint4 ua = make_uint4(0, 0, 0, 0), ub = ua;
                            if constexpr (NW == 1) { ua = *reinterpret_cast<const uint4*>(U); ub = *reinterpret_cast<const uint4*>(U + 4); }
#pragma unroll
                            for (int j = 0; j < 5; ++j) {
                                const uint32_t off = j == 0 ? ((P << 3) & 0x1f8u) : ((P >> (6 * j - 3)) & 0x1f8u);
                                push_filtered(*reinterpret_cast<const double*>(L + j * 512 + off), U, j, ua, ub, ebase + j * (512u + ((uint32_t)NW << 16)) + off);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 5; ++j) {
                                const uint32_t off = j == 0 ? ((P << 3) & 0x1f8u) : ((P >> (6 * j - 3)) & 0x1f8u);
                                push(*reinterpret_cast<const double*>(L + j * 512 + off), ebase + j * (512u + ((uint32_t)NW << 16)) + off);
                            }
                        }
                        ebase += EB_STEP_B;
                        while (__any_sync(0xffffffffu, qtail > q0 + (kQueueCap - 8) * QSTRIDE)) pop_round(buf);
                    }
                }
            } else {
                const unsigned char* p = buf;
                for (uint32_t b = 0; b < tl_count; ++b) {
                    const uint32_t* hdr = reinterpret_cast<const uint32_t*>(p);
                    const uint32_t n_words = hdr[0];
                    double acc = 0.0;  // every blob is self-contained: sum_u-blobs (H_blob * psi(s ^ u)) — tiles of one group may run in different CTAs (table chunks)
                    const unsigned char* rec = p + kBlobHeader;
                    for (uint32_t q = 0; q < n_words; ++q, rec += REC_C) {
                        const uint32_t P = parity_word<NN>(rec, nib);
                        const unsigned char* L = rec + 64 * NN;
#pragma unroll
                        for (int j = 0; j < 5; ++j) {  // chunk sums in term order
                            const uint32_t off = j == 0 ? ((P << 3) & 0x1f8u) : ((P >> (6 * j - 3)) & 0x1f8u);
                            acc = __dadd_rn(acc, *reinterpret_cast<const double*>(L + j * 512 + off));
                        }
                    }
                    {
                        double h[1] = {acc};
                        if constexpr (LK == kLookDense) {
                            if constexpr (PSI32) {
                                const uint32_t u1[1] = {hdr[4] << 3};
                                emit_batch32<SEC, 1>(h, u1, a0, base_hi, valid, sec8, e_re, e_im);
                            } else {
                                const uint32_t u1[1] = {hdr[4]};
                                emit_batch<NW, SEC, KEYORDER, 1>(h, u1, s, valid, sec, lv, e_re, e_im);
                            }
                        }
                        else {
                            const uint32_t* uu[1] = {hdr + 4};
                            heavy_lookup<NW, SEC, 1>((h[0] != 0.0 && valid) ? 1 : 0, h, uu, s, sec, lv, sfilt ? sfilt : gfilt,
                                                     sfilt ? 32 - kFilterLog2WordsSmem : lv.filter_wshift, e_re, e_im);
                        }
                    }
                    p += kBlobHeader + (size_t)n_words * REC_C;
                }
            }
        };

        // (measured alternative: per-buffer "empty" mbarriers released warp by warp + prefetch across state blocks instead of the
        // CTA-wide barrier below — no gain on Li2O, 2 % slower on N2, so the simple form stays)
        if (threadIdx.x == 0 && tile_lo < tile_hi && !(resident && have_resident)) issue(tile_lo, 0);
        for (int t = tile_lo; t < tile_hi; ++t) {
            const int b = resident ? 0 : ((t - tile_lo) & 1);
            if (!resident && threadIdx.x == 0 && t + 1 < tile_hi) issue(t + 1, b ^ 1);
            if (!(resident && have_resident)) {
                if (b == 0) { mbar_wait(&mbar[0], phase0); phase0 ^= 1; }
                else        { mbar_wait(&mbar[1], phase1); phase1 ^= 1; }
            }
            have_resident = resident;
            const uint32_t tl_kind = sv.tiles[t].kind, tl_count = sv.tiles[t].count;
            process(smem + (size_t)b * buf_bytes, tl_kind, tl_count);
            if constexpr (LK == kLookHash) drain(smem + (size_t)b * buf_bytes);
            if (!resident) __syncthreads();  // every thread is done with buffer b before it is refilled
        }

        if (valid) {
            if (KEYORDER || partial) partial[(int64_t)blockIdx.y * M + m] = make_double2(e_re, e_im);
            else out[m] = finalize_row(make_double2(e_re, e_im), psi, psi_dtype, m);
        }
    }
}

// sum the per-chunk partial sums in chunk order, divide by psi(s) and conjugate (energy.py:248)
__global__ void eloc_finalize_kernel(const double2* __restrict__ partial, int n_chunks, const void* __restrict__ psi, int psi_dtype,
                                     int64_t M, double2* __restrict__ out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double re = 0.0, im = 0.0;
    for (int c = 0; c < n_chunks; ++c) {
        const double2 p = partial[(int64_t)c * M + m];
        re = __dadd_rn(re, p.x); im = __dadd_rn(im, p.y);
    }
    out[m] = finalize_row(make_double2(re, im), psi, psi_dtype, m);
}

// key-order mode helpers: mark the keys that occur as rows; gather S[key_m] per row, divide and conjugate
__global__ void mark_keys_kernel(const uint64_t* __restrict__ states, int64_t M, uint32_t* __restrict__ need) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const unsigned long long k = states[m];
    atomicOr(&need[k >> 5], 1u << (k & 31));
}

__global__ void eloc_rows_finalize_kernel(const double2* __restrict__ partial, int n_chunks, int64_t n_keys,
                                          const uint64_t* __restrict__ states, const void* __restrict__ psi, int psi_dtype,
                                          int64_t M, double2* __restrict__ out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const unsigned long long k = states[m];
    double re = 0.0, im = 0.0;
    for (int c = 0; c < n_chunks; ++c) {
        const double2 p = partial[(int64_t)c * n_keys + k];
        re = __dadd_rn(re, p.x); im = __dadd_rn(im, p.y);
    }
    out[m] = finalize_row(make_double2(re, im), psi, psi_dtype, m);
}

}  // namespace naqs
