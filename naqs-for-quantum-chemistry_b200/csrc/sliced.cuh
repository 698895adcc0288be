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
//   A record (8 groups x 4 bits):  T[NN][16] u32 | U[8][NW] u32 (+ U << 3 when NW == 1) | HW[8] HB[8] u32 | Z[16] u32 | LUT[8][16] f64
//   B record (5 groups x 6 bits):  T[NN][16] u32 | U[8][NW] u32 (+ U << 3 when NW == 1) | HW[8] HB[8] u32 | Z[16] u32 | LUT[5][64] f64
//   C blob   (<= 150 terms of one big group; a longer group is several blobs with the same u, each contributing
//             H_blob * psi(s ^ u) on its own):  header{n_words, HW, HB, -, u[4]} | n_words x ( T[NN][16] u32 | LUT[5][64] f64 )
//   HW / HB: the GF(2)-linear Bloom-filter hashes of the flip mask u (lin_hash, common.cuh).  Linear means
//   hash(s ^ u) = hash(s) ^ hash(u): the filter test of a coupled state costs two XORs with the per-thread hash(s).
//   Z: which LUT entries are NOT exactly 0.0 — all the light pass of the hash walk needs from H (hamiltonian.py:363):
//   A: Z[j] = 16-bit mask of group j in both halves of the word (a rotate by the group's parity bits, whatever sits in
//   bit 4 of the amount, lands on the entry's flag); B: Z[2j], Z[2j+1] = low / high half of group j's 64-bit mask.
#pragma once
#include <algorithm>
#include <cstring>
#include <type_traits>
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
inline size_t u_bytes(int nw) { return (nw == 1 ? 64 : (size_t)32 * nw) + 128; }  // masks + their filter hashes HW[8], HB[8] + zero masks Z[16]
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
                for (unsigned pat = 0; pat < (1u << bits); ++pat) L[j * (1 << bits) + pat] = g ? lut_entry(g->c.data(), n, pat) : 0.0;
                {
                    uint32_t* HWB = reinterpret_cast<uint32_t*>(p + 64 * nn + u_bytes(nw) - 128);
                    uint32_t hw = 0, hb = 0;
                    if (g) for (int w = 0; w < nw; ++w) lin_hash_word(g->u[w], w, hw, hb);
                    HWB[j] = hw; HWB[8 + j] = hb;
                    uint64_t z = 0;  // bit e: LUT entry e of this group is not exactly 0.0
                    for (unsigned pat = 0; pat < (1u << bits); ++pat) z |= (uint64_t)(L[j * (1 << bits) + pat] != 0.0) << pat;
                    if (bits == 4) HWB[16 + j] = (uint32_t)z | ((uint32_t)z << 16);
                    else { HWB[16 + 2 * j] = (uint32_t)z; HWB[16 + 2 * j + 1] = (uint32_t)(z >> 32); }
                }
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
// One poll of the barrier phase / the spin loop around it.  ptxas treats the try_wait predicate as warp-uniform (no BSSY / BSYNC
// around the loop in the SASS, whether the branch is written in C++ or inside the asm) and removes any __syncwarp() placed after it.
// compute-sanitizer memcheck reports, for the 128-bit hash walk only, lanes of one warp reading shared memory through stale
// warp-uniform registers (open); plain hardware reproduces the oracle on every row.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
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

// Bloom-filter test of a coupled state s ^ u from the linear hashes (xw, xb) = hash(s) ^ hash(u) (common.cuh): a clear bit
// proves the key is not in the table — ~90 % of the couplings of a large-sector batch end here without touching global
// memory.  `filt` is indexed in BYTES; mask = 4 * n_words - 4.
// mask |= bit  iff  every pattern bit of r (the rotated filter word) is set AND bit 0 of t (the rotated zero mask) is set.
// Spelled in PTX so that it stays two LOP3 + compare + predicated OR (ptxas otherwise expands it into two compare / select chains).
__device__ __forceinline__ void survivor_bit(uint32_t& mask, uint32_t r, uint32_t t, uint32_t bit) {
    asm("{\n .reg .b32 a, v;\n .reg .pred p;\n"
        " lop3.b32 a, %1, %2, 0, 0x0c;\n"   // a = ~r & pattern
        " lop3.b32 v, a, %3, 1, 0xf2;\n"    // v = a | (~t & 1)
        " setp.eq.u32 p, v, 0;\n"
        " @p or.b32 %0, %0, %4;\n}"
        : "+r"(mask) : "r"(r), "r"(kFilterPattern), "r"(t), "r"(bit));
}

__device__ __forceinline__ bool filter_pass(uint32_t xw, uint32_t xb, const unsigned char* __restrict__ filt, uint32_t mask) {
    const uint32_t word = *reinterpret_cast<const uint32_t*>(filt + (xw & mask));
    return (~__funnelshift_r(word, word, xb) & kFilterPattern) == 0u;  // rotate right by xb & 31: the key's bits land on the pattern
}

// Hash-lookup ("heavy") epilogue of up to B couplings of one thread: optional Bloom-filter test (`filt`: the filter in
// global memory when the light path could not consult a shared-memory copy, else nullptr), probe of the bucketed table
// (keys <= 63 bits: one 256-bit load reads the four keys of a 128-byte bucket) or of the 32-byte-slot table (wider keys),
// complex multiply-add.  Called for couplings whose H is not exactly 0.0 (hamiltonian.py:363) — and, in the filter shape,
// whose coupled state passed the filter — which the per-thread queue of the kernel below makes dense across the warp.
// The first probes of all B couplings are issued before any is examined, so their L2 latencies overlap.  SEC: sector test
// on the coupled state (hamiltonian.py:328); unused by the library's own tables, which hold in-sector keys only.
// (xw, xb)[b]: linear filter hashes of the coupled state.
template <int NW, bool SEC, int B>
__device__ __forceinline__ void heavy_lookup(int n, const double (&h)[B], const uint32_t* const (&u)[B], const uint32_t (&xw)[B], const uint32_t (&xb)[B],
                                             const uint32_t (&s)[NW], const Sector& sec, const LookupView& lv, const unsigned char* __restrict__ filt,
                                             uint32_t filt_mask, double& e_re, double& e_im) {
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
        if (filt) on[b] = on[b] & filter_pass(xw[b], xb[b], filt, filt_mask);  // a clear bit proves the key is not in the table
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

// pending RECORDS per thread (hash mode): 8-byte entries {parity word, survivor mask | record}; a ring (FIFO), so a thread
// resolves its couplings in record order whatever the rest of its warp does — E_loc is reproducible bit for bit
constexpr int kQueueCap = 8, kQueueCapFilter = 4;  // without / with the filter in shared memory (survivors: ~60 % / ~1.5 % of the pairs)

// Bank-binned order of a hash-lookup batch (bin_states_kernel): position p holds row perm[p] (-1 = empty); the first
// `base` positions are 32-wide rows whose lane l holds a state with filter bank l, the rest is the overflow of full bins.
struct BinView {
    const int32_t* perm;      // nullptr: rows are walked in their own order
    const int32_t* counters;  // [33]: states per bank, [32] = overflow count
    int64_t base;
};

template <int NW>
__device__ __forceinline__ bool key_in_range(const uint32_t (&s)[NW], int n_qubits) {
    const int w = n_qubits >> 5, r = n_qubits & 31;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        if (i > w) ok = ok && s[i] == 0u;
        else if (i == w) ok = ok && (s[i] >> r) == 0u;
    }
    return ok;
}

// One state per thread.  Work is cut into TASKS = (table chunk, state block), chunk-major: task = chunk * n_blocks + block.
// A CTA walks the tiles [tile_lo, tile_hi) of the task's chunk for the task's block of states; with a single tile per chunk
// the table stays resident in shared memory while the chunk does not change.  Tasks are dealt statically (CTA i takes tasks
// i, i + gridDim.x, ...) or, when `task_counter` is given, dynamically from an atomic counter — the hash walk's tasks differ
// in length (survivors, overflowing buckets), and with dynamic dealing the launch ends within one task of the ideal.
//
// KEYORDER (dense lookup, batch dense in key space): thread m IS key m (states == nullptr, M = 2^N); `need` is a bitmap
// of the keys that occur as rows; the raw sums S[k] = sum_u H[k, k^u] psi(k^u) go to partial[chunk * M + k] and
// eloc_rows_finalize_kernel turns them into E_loc per row.  A warp then holds 32 consecutive keys and every table
// read of a group falls into one aligned 512-byte block: 4 L1 lines per request instead of ~11 scattered sectors.
//
// Hash lookup (LK == kLookHash), the sparse-batch case (a VMC batch of a large sector): almost every coupled state is NOT
// in the table, so the walk is split into a LIGHT pass over every (state, group) pair — H from the group LUT, exact-zero
// test (hamiltonian.py:363), Bloom-filter test in shared memory from linear hashes (two XORs, one LDS, one rotate) — that
// only records, per record of 8 (5) groups, a bit mask of the survivors, and a HEAVY pass that resolves the survivors
// (a few % of the pairs) against the bucketed table in global memory, one coupling per lane and round, from a per-thread
// queue of {parity word, survivor mask | record offset} entries.
template <int NW, int NN, int THREADS, int CTAS_PER_SM, int LK, bool SEC, bool KEYORDER, bool PSI32>
__global__ void __launch_bounds__(THREADS, CTAS_PER_SM)
eloc_sliced_kernel(SlicedView sv, const __grid_constant__ ChunkBounds chunks, uint32_t buf_bytes, uint32_t queue_offset, uint32_t queue_cap, uint32_t filter_offset, Sector sec, LookupView lv,
                   const uint64_t* __restrict__ states, const uint32_t* __restrict__ need, const void* __restrict__ psi,
                   int psi_dtype, int64_t M, BinView bin, int n_chunks, int* __restrict__ task_counter, double2* __restrict__ out,
                   double2* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ int64_t s_task;
    constexpr int UB = (NW == 1 ? 64 : 32 * NW) + 128;  // u_bytes(NW): flip masks, their filter hashes HW[8] HB[8], zero masks Z[16]
    constexpr int HOFF = UB - 128, ZOFF = UB - 64;
    constexpr int REC_A = 64 * NN + UB + 1024, REC_B = 64 * NN + UB + 2560, REC_C = 64 * NN + 2560;
    static_assert(REC_A % 16 == 0 && REC_B % 16 == 0, "queue entries address records in 16-byte units");

    if (threadIdx.x == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const unsigned char* sfilt = nullptr;  // Bloom filter of the table keys in shared memory (hash lookup, 1024-thread shape, <= 2.5 * 2^18 keys)
    const unsigned char* gfilt = nullptr;  // ... or in global memory (L2), consulted in the probe rounds
    if constexpr (LK == kLookHash) {
        if (lv.filter && lv.filter_in_smem) {
            uint4* dst = reinterpret_cast<uint4*>(smem + filter_offset);
            const uint4* src = reinterpret_cast<const uint4*>(lv.filter_small ? lv.filter_small : lv.filter);
            for (uint32_t i = threadIdx.x; i < kFilterBytes / 16; i += THREADS) dst[i] = __ldg(src + i);
            sfilt = smem + filter_offset;
            if (lv.filter_small) gfilt = reinterpret_cast<const unsigned char*>(lv.filter);  // the full-size filter screens the survivors once more before a bucket read
        } else {
            gfilt = reinterpret_cast<const unsigned char*>(lv.filter);
        }
    }
    __syncthreads();
    // complex64 direct-address table (PSI32): base split into words, sector masks shifted like the keys (see emit_batch32)
    const uint32_t base_lo = (uint32_t)reinterpret_cast<unsigned long long>(lv.dense32);
    const uint32_t base_hi = (uint32_t)(reinterpret_cast<unsigned long long>(lv.dense32) >> 32);
    Sector sec8 = sec;
    sec8.even[0] <<= 3; sec8.odd[0] <<= 3;
    uint32_t phase0 = 0, phase1 = 0;
    int resident_chunk = -1;  // chunk whose single tile currently sits in buffer 0

    auto issue = [&](int t, int b) {  // thread 0 only
        const STile tl = sv.tiles[t];
        mbar_expect_tx(&mbar[b], tl.bytes);
        bulk_g2s(smem + (size_t)b * buf_bytes, sv.stream + tl.offset, tl.bytes, &mbar[b]);
    };

    // positions to walk: the rows themselves, or the bank-binned arrangement of them (hash lookup with the filter in shared memory)
    const int64_t n_pos = bin.perm ? bin.base + (int64_t)__ldg(bin.counters + 32) : M;
    const int64_t n_blocks = (n_pos + THREADS - 1) / THREADS;
    const int64_t n_tasks = n_blocks * n_chunks;
    for (int64_t task = blockIdx.x;; task += gridDim.x) {
        if (task_counter) {  // dynamic dealing
            __syncthreads();  // everyone is done with s_task of the previous round
            if (threadIdx.x == 0) s_task = (int64_t)atomicAdd(task_counter, 1);
            __syncthreads();
            task = s_task;
        }
        if (task >= n_tasks) break;
        const int chunk = (int)(task / n_blocks);
        const int64_t blk = task - (int64_t)chunk * n_blocks;
        const int tile_lo = chunks.lo[chunk], tile_hi = chunks.lo[chunk + 1];
        const bool resident = (tile_hi - tile_lo) <= 1;
        const bool have_resident = resident && resident_chunk == chunk;
        int64_t m = blk * THREADS + threadIdx.x;
        bool valid = m < n_pos;
        if (bin.perm) {
            const int32_t row = valid ? __ldg(bin.perm + m) : -1;
            valid = row >= 0;
            m = row;
        }
        uint32_t s[NW];
        if constexpr (KEYORDER) {
            s[0] = valid ? (uint32_t)m : 0u;  // threads past the key space read entry 0 (table reads are unconditional)
#pragma unroll
            for (int w = 1; w < NW; ++w) s[w] = 0;
            valid = valid && ((need[m >> 5] >> (m & 31)) & 1u);
        } else {
            if (valid) {
                load_key<NW>(states, m, s);
                if (!key_in_range<NW>(s, sec.n_qubits)) {  // the reference raises IndexError; here: flag (naqs_table_check), row -> NaN
                    atomicOr(lv.flags, 1);
                    const double2 nan2 = make_double2(__longlong_as_double(0x7ff8000000000000ll), 0.0);
                    if (partial) partial[(int64_t)chunk * M + m] = nan2;
                    else out[m] = nan2;
                    valid = false;
                }
            }
            if (!valid) {
#pragma unroll
                for (int w = 0; w < NW; ++w) s[w] = 0;
            }
        }
        uint32_t nib[NN];
        state_nibbles<NW, NN>(s, nib);
        const uint32_t a0 = base_lo ^ (s[0] << 3);  // PSI32 only
        double e_re = 0.0, e_im = 0.0;

        // ---- hash mode state: linear filter hashes of the state, the per-thread queue ([slot][thread] layout, 8-byte entries
        // {P, meta}: P = parity word of the record, meta = survivor mask (bits 0-7) | record offset / 16 (bits 8-19) | B record
        // (bit 31)) and the entry being resolved.  The queue is drained before a tile buffer is released.
        [[maybe_unused]] uint32_t hsw = 0, hsb = 0;
        if constexpr (LK == kLookHash) {
#pragma unroll
            for (int w = 0; w < NW; ++w) lin_hash_word(s[w], w, hsw, hsb);
        }
        constexpr uint32_t QSTRIDE = THREADS * 8;  // bytes between consecutive queue slots of one thread
        unsigned char* const q0 = smem + queue_offset + threadIdx.x * 8;
        const uint32_t qwrap = queue_cap * QSTRIDE - QSTRIDE;  // ring of queue_cap (a power of two) slots: offset & qwrap
        uint32_t qhead = 0, qtail = 0;        // byte offsets (multiples of QSTRIDE, free-running); the ring holds (qtail - qhead) / QSTRIDE entries
        uint32_t cur_P = 0, cur_meta = 0;     // entry being resolved: (cur_meta & 0xff) = couplings still pending
        // one coupling per lane: take the lowest pending group of the current entry (refilled from the ring when exhausted)
        auto resolve_round = [&](const unsigned char* __restrict__ buf) {
            if ((cur_meta & 0xffu) == 0u && qhead != qtail) {
                const uint2 e = *reinterpret_cast<const uint2*>(q0 + (qhead & qwrap));
                qhead += QSTRIDE;
                cur_P = e.x; cur_meta = e.y;
            }
            const uint32_t pend = cur_meta & 0xffu;
            const int n = pend ? 1 : 0;
            const uint32_t j = pend ? (uint32_t)__ffs((int)pend) - 1u : 0u;
            cur_meta = pend ? (cur_meta & (cur_meta - 1u)) : cur_meta;  // clear the lowest pending bit (the borrow stays inside the mask field)
            const unsigned char* rec = buf + ((cur_meta >> 8) & 0xfffu) * 16u;
            const uint32_t bits = (cur_meta >> 31) ? 6u : 4u;
            const uint32_t idx = (cur_P >> (j * bits)) & ((1u << bits) - 1u);
            const double h[1] = {*reinterpret_cast<const double*>(rec + 64 * NN + UB + (((j << bits) + idx) << 3))};
            const uint32_t* U = reinterpret_cast<const uint32_t*>(rec + 64 * NN);
            const uint32_t* HWB = reinterpret_cast<const uint32_t*>(rec + 64 * NN + HOFF);
            const uint32_t* u[1] = {U + j * NW};
            const uint32_t xw[1] = {hsw ^ HWB[j]}, xb[1] = {hsb ^ HWB[8 + j]};
            heavy_lookup<NW, SEC, 1>(n, h, u, xw, xb, s, sec, lv, gfilt, lv.filter_mask, e_re, e_im);  // a shared-memory filter was consulted before queueing
        };
        auto pending = [&]() { return qhead != qtail || (cur_meta & 0xffu) != 0u; };
        auto drain = [&](const unsigned char* __restrict__ buf) {  // before a tile buffer is released: its offsets die with it
            while (__any_sync(0xffffffffu, pending())) resolve_round(buf);
        };
        const uint32_t q_adv = valid ? QSTRIDE : 0u;  // an invalid lane never keeps an entry
        // light pass over one record: survivor mask of its groups (H != 0 and, with the filter in shared memory, filter passed).
        // Nothing of H itself is read: bit `entry` of the group's zero mask Z says whether the LUT entry is exactly 0.0.
        auto light_record = [&](const unsigned char* __restrict__ buf, const unsigned char* __restrict__ rec, uint32_t rec_meta, auto kind_tag) {
            constexpr bool KB = decltype(kind_tag)::value;
            constexpr int G = KB ? 5 : 8, LB = KB ? 6 : 4;
            const uint32_t P = parity_word<NN>(rec, nib);
            const uint32_t* Z = reinterpret_cast<const uint32_t*>(rec + 64 * NN + ZOFF);
            uint32_t z[10];
            {
                const uint4 za = *reinterpret_cast<const uint4*>(Z), zb = *reinterpret_cast<const uint4*>(Z + 4);
                z[0] = za.x; z[1] = za.y; z[2] = za.z; z[3] = za.w; z[4] = zb.x; z[5] = zb.y; z[6] = zb.z; z[7] = zb.w;
                if constexpr (KB) { const uint2 zc = *reinterpret_cast<const uint2*>(Z + 8); z[8] = zc.x; z[9] = zc.y; }
                else { z[8] = z[9] = 0; }
            }
            // flag of group j in bit 0 of nz(j)
            auto nz = [&](int j) -> uint32_t {
                const uint32_t sh = j == 0 ? P : (P >> (LB * j));
                if constexpr (KB) {
                    const uint32_t half = (sh & 32u) ? z[2 * j + 1] : z[2 * j];  // 64-entry mask: bit 5 of the entry picks the word
                    return __funnelshift_r(half, half, sh);
                } else {
                    return __funnelshift_r(z[j], z[j], sh);  // both halves hold the mask: bit 4 of the amount is irrelevant
                }
            };
            uint32_t mask = 0;
            if (sfilt) {  // warp-uniform
                const uint32_t* HWB = reinterpret_cast<const uint32_t*>(rec + 64 * NN + HOFF);
                const uint4 wa = *reinterpret_cast<const uint4*>(HWB), wb = *reinterpret_cast<const uint4*>(HWB + 4);
                const uint4 ba = *reinterpret_cast<const uint4*>(HWB + 8), bb = *reinterpret_cast<const uint4*>(HWB + 12);
                const uint32_t hw[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w}, hb[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                uint32_t word[G];
#pragma unroll
                for (int j = 0; j < G; ++j)  // the shared-memory copy always has 2^15 words
                    word[j] = *reinterpret_cast<const uint32_t*>(sfilt + ((hsw ^ hw[j]) & (kFilterBytes - 4u)));
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    // straight-line bit logic (no lazy evaluation): kill != 0 <=> a filter bit is clear or the LUT entry is exactly 0.0
                    const uint32_t r = __funnelshift_r(word[j], word[j], hsb ^ hb[j]);
                    survivor_bit(mask, r, nz(j), 1u << j);
                }
            } else {
#pragma unroll
                for (int j = 0; j < G; ++j) mask |= (nz(j) & 1u) << j;
            }
            // branch-free push: the entry is always written at the tail, the tail only advances when a group survived
            const uint32_t meta = mask | rec_meta;  // record offset / 16 << 8 | B record << 31 (kept by the record loop)
            *reinterpret_cast<uint2*>(q0 + (qtail & qwrap)) = make_uint2(P, meta);
            qtail += mask ? q_adv : 0u;
            while (__any_sync(0xffffffffu, qtail - qhead > qwrap)) resolve_round(buf);  // a ring is full: the next push needs its tail slot
        };

        auto process = [&](const unsigned char* __restrict__ buf, const uint32_t tl_kind, const uint32_t tl_count) {
            if (tl_kind == kSecA) {
                const unsigned char* rec = buf;
                [[maybe_unused]] uint32_t rec_meta = 0;
                for (uint32_t r = 0; r < tl_count; ++r, rec += REC_A, rec_meta += (REC_A / 16) << 8) {
                    if constexpr (LK == kLookDense) {
                        const uint32_t P = parity_word<NN>(rec, nib);
                        const uint32_t* U = reinterpret_cast<const uint32_t*>(rec + 64 * NN);
                        const unsigned char* L = rec + 64 * NN + UB;
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
                        light_record(buf, rec, rec_meta, std::false_type{});
                    }
                }
            } else if (tl_kind == kSecB) {
                const unsigned char* rec = buf;
                [[maybe_unused]] uint32_t rec_meta = 0x80000000u;
                for (uint32_t r = 0; r < tl_count; ++r, rec += REC_B, rec_meta += (REC_B / 16) << 8) {
                    if constexpr (LK == kLookDense) {
                        const uint32_t P = parity_word<NN>(rec, nib);
                        const uint32_t* U = reinterpret_cast<const uint32_t*>(rec + 64 * NN);
                        const unsigned char* L = rec + 64 * NN + UB;
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
                        light_record(buf, rec, rec_meta, std::true_type{});
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
                            const uint32_t xw[1] = {hsw ^ hdr[1]}, xb[1] = {hsb ^ hdr[2]};
                            heavy_lookup<NW, SEC, 1>((h[0] != 0.0 && valid) ? 1 : 0, h, uu, xw, xb, s, sec, lv, sfilt ? sfilt : gfilt,
                                                     sfilt ? kFilterBytes - 4u : lv.filter_mask, e_re, e_im);
                        }
                    }
                    p += kBlobHeader + (size_t)n_words * REC_C;
                }
            }
        };

        // (measured alternative: per-buffer "empty" mbarriers released warp by warp + prefetch across state blocks instead of the
        // CTA-wide barrier below — no gain on Li2O, 2 % slower on N2, so the simple form stays)
        // a resident tile is read without CTA barriers: before another chunk's tile replaces it, every thread must have left it
        // (dynamic dealing has barriers at the top of the task loop)
        if (!have_resident && resident_chunk != -1 && !task_counter) __syncthreads();
        if (threadIdx.x == 0 && tile_lo < tile_hi && !have_resident) issue(tile_lo, 0);
        for (int t = tile_lo; t < tile_hi; ++t) {
            const int b = resident ? 0 : ((t - tile_lo) & 1);
            if (!resident && threadIdx.x == 0 && t + 1 < tile_hi) issue(t + 1, b ^ 1);
            if (!have_resident) {
                if (b == 0) { mbar_wait(&mbar[0], phase0); phase0 ^= 1; }
                else        { mbar_wait(&mbar[1], phase1); phase1 ^= 1; }
            }
            const uint32_t tl_kind = sv.tiles[t].kind, tl_count = sv.tiles[t].count;
            process(smem + (size_t)b * buf_bytes, tl_kind, tl_count);
            if constexpr (LK == kLookHash) drain(smem + (size_t)b * buf_bytes);
            if (!resident) __syncthreads();  // every thread is done with buffer b before it is refilled
        }
        resident_chunk = (resident && tile_lo < tile_hi) ? chunk : -1;

        if (valid) {
            if (KEYORDER || partial) partial[(int64_t)chunk * M + m] = make_double2(e_re, e_im);
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
__global__ void mark_keys_kernel(const uint64_t* __restrict__ states, int64_t M, uint32_t* __restrict__ need, int64_t n_keys, int* __restrict__ flags) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const unsigned long long k = states[m];
    if (k >= (unsigned long long)n_keys) { atomicOr(flags, 1); return; }  // out of range: flagged, never used as an index
    atomicOr(&need[k >> 5], 1u << (k & 31));
}

__global__ void eloc_rows_finalize_kernel(const double2* __restrict__ partial, int n_chunks, int64_t n_keys,
                                          const uint64_t* __restrict__ states, const void* __restrict__ psi, int psi_dtype,
                                          int64_t M, double2* __restrict__ out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const unsigned long long k = states[m];
    if (k >= (unsigned long long)n_keys) { out[m] = make_double2(__longlong_as_double(0x7ff8000000000000ll), 0.0); return; }
    double re = 0.0, im = 0.0;
    for (int c = 0; c < n_chunks; ++c) {
        const double2 p = partial[(int64_t)c * n_keys + k];
        re = __dadd_rn(re, p.x); im = __dadd_rn(im, p.y);
    }
    out[m] = finalize_row(make_double2(re, im), psi, psi_dtype, m);
}

}  // namespace naqs
