// keyorder.cuh — the key-order walk of the fused E_loc kernel (v3): dense complex64 lookup, batch dense in key space.
//
// A warp owns a TASK = 32 consecutive keys (key = task * 32 + lane); every table read psi[key ^ u] of an XY group then
// falls into one aligned 256-byte block of the complex64 direct-address table.  What changed against the generic sliced
// kernel (sliced.cuh), all to cut L1TEX wavefronts — the pipe that binds this kernel (profiles/):
//   * parity word of a record = TL[lane] ^ TW[task]: the key's low 5 bits ARE the lane, so their contribution is a 32-entry
//     table read conflict-free by the 32 lanes (one wavefront), and everything above bit 4 is warp-uniform: its contribution
//     TW[r] (one word per parity word r of the chunk) is computed once per task from 5-bit slice tables HT in global
//     memory (L2) into a warp-private shared-memory strip and read back with a broadcast LDS.  2 LDS per 32 terms
//     instead of one per nibble (5 for N2).
//   * table reads are predicated per lane on H != 0.0 (and on the key being a row): exact zeros are 2/3 of all matrix
//     elements on an unrestricted space and they cluster by lane, so a read touches 1.13 lines on average instead of 2;
//     the float -> double conversions and FMAs of an all-zero group are predicated off with it.  (hamiltonian.py:363 drops
//     exact zeros, so skipping them is the reference's own semantics — a non-finite amplitude cannot leak into a row
//     that does not couple to it.)
//   * the whole chunk of the table is resident in shared memory for the CTA's lifetime (N2: 165 KB in one chunk); warps
//     never meet at a CTA barrier after the initial load, tasks are dealt round-robin to warps.
// H is read from the same group LUTs as in sliced.cuh (entries = the reference's serial fp64 sums, hamiltonian_math.pyx:31-34;
// groups of > 6 terms in 6-term chunks added in term order).
//
// Stream layout (device memory, bulk-copied into shared memory once per CTA):
//   A record (8 groups x 4 bits):  TL[32] u32 | U8[8] u32 (flip masks << 3 = byte offsets into the table) | LUT[8][16] f64
//   B record (5 groups x 6 bits):  TL[32] u32 | U8[8] u32 (5 used)                                         | LUT[5][64] f64
//   C word   (30 terms of one big group):   TL[32] u32 | {u8, last word of its blob?, -, -}                   | LUT[5][64] f64
//   HT[j][v][r]: contribution of key bits [5j+9 : 5j+5] == v to parity word r (j = 0 .. n_hi-1), r-contiguous (coalesced)
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "eloc_kernels.cuh"
#include "sliced.cuh"

namespace naqs {

constexpr size_t kKoRecA = 128 + 32 + 8 * 16 * 8;   // 1184
constexpr size_t kKoRecB = 128 + 32 + 5 * 64 * 8;   // 2720
constexpr size_t kKoRecC = 128 + 16 + 5 * 64 * 8;   // 2704: TL | {u8, last, -, -} | LUT

struct KoView {
    const unsigned char* stream;
    const uint32_t* ht;  // [n_hi][32][r_total_pad]
    int n_hi;            // hi slices: ceil((n_qubits - 5) / 5)
    int r_total_pad;     // parity words of the whole stream, padded to a multiple of 32
};

struct KoHost {
    std::vector<unsigned char> stream;
    std::vector<uint32_t> ht;
    std::vector<KoUnit> units;   // in stream order; unit i is parity word i
    int n_hi = 0, r_total_pad = 0;
};

// parity of popcount(yz & (v << shift)) for the 32 term-bits of a word, v = 0..31
inline void ko_slice_table(uint32_t* dst, size_t stride, const uint32_t* const* yz_of_bit, int n_bits, int shift) {
    for (int v = 0; v < 32; ++v) {
        uint32_t word = 0;
        for (int b = 0; b < n_bits; ++b) {
            const uint32_t* yz = yz_of_bit[b];
            if (!yz || shift >= 32) continue;
            word |= (uint32_t)(__builtin_popcount((yz[0] >> shift) & 31u & (uint32_t)v) & 1) << b;
        }
        dst[(size_t)v * stride] = word;
    }
}

// Groups (single-word keys, 5 <= n_qubits <= 26) -> key-order stream, one unit per parity word.
inline void build_ko_host(const std::vector<HostGroup>& groups, int n_qubits, size_t max_blob_words, KoHost& out) {
    std::vector<const HostGroup*> ga, gb, gc;
    for (const auto& g : groups) {
        const size_t n = g.c.size();
        (n <= 4 ? ga : (n <= 6 ? gb : gc)).push_back(&g);
    }
    out.n_hi = std::max(0, (n_qubits - 5 + 4) / 5);
    const size_t n_rec_a = (ga.size() + 7) / 8, n_rec_b = (gb.size() + 4) / 5;
    size_t r_total = n_rec_a + n_rec_b;
    for (const HostGroup* g : gc) r_total += (g->c.size() + 29) / 30;
    out.r_total_pad = (int)std::max<size_t>(32, (r_total + 31) / 32 * 32);
    out.ht.assign((size_t)std::max(out.n_hi, 1) * 32 * out.r_total_pad, 0u);
    auto& S = out.stream;
    S.clear();
    out.units.clear();
    uint32_t r = 0;
    auto hi_tables = [&](const uint32_t* const* yz_of_bit, int n_bits, uint32_t rr) {
        for (int j = 0; j < out.n_hi; ++j)
            ko_slice_table(out.ht.data() + (size_t)j * 32 * out.r_total_pad + rr, (size_t)out.r_total_pad, yz_of_bit, n_bits, 5 + 5 * j);
    };
    auto pack = [&](const std::vector<const HostGroup*>& gs, size_t n_rec, size_t rec, int per, int bits, uint32_t kind) {
        for (size_t q = 0; q < n_rec; ++q, ++r) {
            const size_t off = S.size();
            S.resize(off + rec, 0);
            out.units.push_back(KoUnit{kind, (uint32_t)rec, (uint32_t)per, 1u});
            unsigned char* p = S.data() + off;
            const uint32_t* yz_of_bit[32] = {nullptr};
            uint32_t* U8 = reinterpret_cast<uint32_t*>(p + 128);
            double* L = reinterpret_cast<double*>(p + 160);
            for (int j = 0; j < per; ++j) {
                const size_t gi = q * per + j;
                const HostGroup* g = gi < gs.size() ? gs[gi] : nullptr;
                const int nt = g ? (int)g->c.size() : 0;
                for (int t = 0; t < nt; ++t) yz_of_bit[j * bits + t] = g->yz.data() + t;  // nw == 1
                U8[j] = g ? g->u[0] << 3 : 0u;
                for (unsigned pat = 0; pat < (1u << bits); ++pat) L[j * (1 << bits) + pat] = g ? lut_entry(g->c.data(), nt, pat) : 0.0;
            }
            ko_slice_table(reinterpret_cast<uint32_t*>(p), 1, yz_of_bit, per * bits, 0);
            hi_tables(yz_of_bit, per * bits, r);
        }
    };
    pack(ga, n_rec_a, kKoRecA, 8, 4, kSecA);
    pack(gb, n_rec_b, kKoRecB, 5, 6, kSecB);
    // big groups: 30 terms (five 6-term chunks) per word; a blob = at most max_blob_words consecutive words whose chunk sums
    // are added in term order and multiplied by psi(key ^ u) once, at its last word (a longer group is several blobs)
    for (const HostGroup* g : gc) {
        const size_t n = g->c.size(), total_words = (n + 29) / 30;
        for (size_t q = 0; q < total_words; ++q, ++r) {
            const size_t off = S.size();
            S.resize(off + kKoRecC, 0);
            const bool last = q + 1 == total_words || (q + 1) % max_blob_words == 0;
            out.units.push_back(KoUnit{kSecC, (uint32_t)kKoRecC, last ? 3u : 2u, last ? 1u : 0u});
            unsigned char* p = S.data() + off;
            uint32_t* hdr = reinterpret_cast<uint32_t*>(p + 128);
            hdr[0] = g->u[0] << 3;
            hdr[1] = last ? 1u : 0u;
            const uint32_t* yz_of_bit[32] = {nullptr};
            double* L = reinterpret_cast<double*>(p + 144);
            for (int j = 0; j < 5; ++j) {
                const size_t t0 = std::min(n, q * 30 + (size_t)j * 6), t1 = std::min(n, t0 + 6);
                for (size_t t = t0; t < t1; ++t) yz_of_bit[j * 6 + (t - t0)] = g->yz.data() + t;
                for (unsigned pat = 0; pat < 64; ++pat) L[j * 64 + pat] = lut_entry(g->c.data() + t0, (int)(t1 - t0), pat);  // empty chunk: +0.0
            }
            ko_slice_table(reinterpret_cast<uint32_t*>(p), 1, yz_of_bit, 30, 0);
            hi_tables(yz_of_bit, 30, r);
        }
    }
}

// ------------------------------------------------------------------------------------------ kernel
// e += H_b * psi(key ^ u_b) for a batch of groups against the complex64 table aligned to its size (entry address = a0 ^ u8,
// see sliced.cuh).  The table read of a group is predicated per lane on (key is a row) && H_b != 0.0, and the reads of the whole
// batch are issued before any is consumed.  The read destinations q are caller-owned registers that keep their previous — finite
// — contents where the predicate is off, so the arithmetic needs no predicate or select: H_b == 0 adds exactly 0, and lanes
// whose key is not a row are never stored.  (ptxas turns predicated conversions / FMAs into unconditional ones plus selects,
// measured; this form has neither.)  Amplitudes are assumed finite.
struct KoRegs {
    float2 q[5];
};

template <int B>
__device__ __forceinline__ void ko_emit(const double (&h)[B], const uint32_t (&u8)[B], uint32_t a0, uint32_t base_hi, bool valid, KoRegs& r,
                                        double& e_re, double& e_im) {
#pragma unroll
    for (int b = 0; b < B; ++b) {
        unsigned long long addr;
        asm("mov.b64 %0, {%1, %2};" : "=l"(addr) : "r"(a0 ^ u8[b]), "r"(base_hi));
        if (valid && h[b] != 0.0) r.q[b] = __ldg(reinterpret_cast<const float2*>(addr));
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
        e_re = __fma_rn(h[b], (double)r.q[b].x, e_re);
        e_im = __fma_rn(h[b], (double)r.q[b].y, e_im);
    }
}

// grid = (CTAs per chunk, n_chunks).  partial[c * n_keys + key] receives the raw row sums of chunk c (eloc_rows_finalize_kernel
// adds the chunks in order, divides by psi and conjugates).  Tasks (32 consecutive keys) are dealt round-robin to the warps of
// the CTAs of a chunk; there is no CTA-wide synchronisation after the chunk has landed in shared memory.
template <int THREADS, int CTAS_PER_SM>
__global__ void __launch_bounds__(THREADS, CTAS_PER_SM)
eloc_keyorder_kernel(KoView kv, const __grid_constant__ KoChunks chunks, uint32_t tw_offset, uint32_t tw_stride, const float2* __restrict__ dense32,
                     const uint32_t* __restrict__ need, int64_t n_tasks, int64_t n_keys, double2* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar;
    constexpr int WARPS = THREADS / 32;
    // chunk descriptor fields, read from the constant bank (warp-uniform)
    const struct { uint32_t off, bytes, n_a, n_b, n_c, off_b, off_c, r0, n_words; } ck = {
        chunks.c[blockIdx.y].off, chunks.c[blockIdx.y].bytes, chunks.c[blockIdx.y].n_a, chunks.c[blockIdx.y].n_b, chunks.c[blockIdx.y].n_c,
        chunks.c[blockIdx.y].off_b, chunks.c[blockIdx.y].off_c, chunks.c[blockIdx.y].r0, chunks.c[blockIdx.y].n_words};
    if (threadIdx.x == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&mbar, ck.bytes);
        bulk_g2s(smem, kv.stream + ck.off, ck.bytes, &mbar);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: keeps the record loops on the uniform datapath
    const uint32_t lane4 = lane * 4u;
    uint32_t* const tw = reinterpret_cast<uint32_t*>(smem + tw_offset) + (size_t)warp * tw_stride;
    const uint32_t base_lo = (uint32_t)reinterpret_cast<unsigned long long>(dense32);
    const uint32_t base_hi = (uint32_t)(reinterpret_cast<unsigned long long>(dense32) >> 32);
    const unsigned char* const rec_b0 = smem + ck.off_b;
    const unsigned char* const rec_c0 = smem + ck.off_c;
    bool loaded = false;
    KoRegs q;
#pragma unroll
    for (int b = 0; b < 5; ++b) q.q[b] = make_float2(0.f, 0.f);

    for (int64_t task = (int64_t)blockIdx.x * WARPS + warp; task < n_tasks; task += (int64_t)gridDim.x * WARPS) {
        const uint32_t need_word = __shfl_sync(0xffffffffu, need[task], 0);
        if (need_word == 0) continue;  // none of the 32 keys is a row (warp-uniform)
        // warp-uniform part of every parity word of the chunk: XOR of the 5-bit slice tables at the task's bits
        for (uint32_t r = lane; r < ck.n_words; r += 32) {
            uint32_t x = 0;
            for (int j = 0; j < kv.n_hi; ++j) {
                const uint32_t v = (uint32_t)(task >> (5 * j)) & 31u;
                x ^= __ldg(kv.ht + ((size_t)j * 32 + v) * kv.r_total_pad + ck.r0 + r);
            }
            tw[r] = x;
        }
        __syncwarp();
        if (!loaded) { mbar_wait(&mbar, 0); loaded = true; }
        const uint32_t key = (uint32_t)task * 32u + lane;
        const bool valid = (need_word >> lane) & 1u;
        const uint32_t a0 = base_lo ^ (key << 3);
        double e_re = 0.0, e_im = 0.0;
        const uint32_t* twp = tw;
        for (uint32_t i = 0; i < ck.n_a; ++i) {
            const unsigned char* rec = smem + (size_t)i * kKoRecA;
            const uint32_t P = *reinterpret_cast<const uint32_t*>(rec + lane4) ^ twp[i];
            const uint4 ua = *reinterpret_cast<const uint4*>(rec + 128), ub = *reinterpret_cast<const uint4*>(rec + 144);
            const unsigned char* L = rec + 160;
            double h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t off = j == 0 ? ((P << 3) & 0x78u) : ((P >> (4 * j - 3)) & 0x78u);
                h[j] = *reinterpret_cast<const double*>(L + j * 128 + off);
            }
            {
                const uint32_t uu[4] = {ua.x, ua.y, ua.z, ua.w};
                ko_emit<4>(h, uu, a0, base_hi, valid, q, e_re, e_im);
            }
#pragma unroll
            for (int j = 4; j < 8; ++j) {
                const uint32_t off = (P >> (4 * j - 3)) & 0x78u;
                h[j - 4] = *reinterpret_cast<const double*>(L + j * 128 + off);
            }
            {
                const uint32_t uu[4] = {ub.x, ub.y, ub.z, ub.w};
                ko_emit<4>(h, uu, a0, base_hi, valid, q, e_re, e_im);
            }
        }
        twp += ck.n_a;
        for (uint32_t i = 0; i < ck.n_b; ++i) {
            const unsigned char* rec = rec_b0 + (size_t)i * kKoRecB;
            const uint32_t P = *reinterpret_cast<const uint32_t*>(rec + lane4) ^ twp[i];
            const uint4 ua = *reinterpret_cast<const uint4*>(rec + 128);
            const uint32_t u4 = *reinterpret_cast<const uint32_t*>(rec + 144);
            const unsigned char* L = rec + 160;
            double h[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const uint32_t off = j == 0 ? ((P << 3) & 0x1f8u) : ((P >> (6 * j - 3)) & 0x1f8u);
                h[j] = *reinterpret_cast<const double*>(L + j * 512 + off);
            }
            const uint32_t uu[5] = {ua.x, ua.y, ua.z, ua.w, u4};
            ko_emit<5>(h, uu, a0, base_hi, valid, q, e_re, e_im);
        }
        twp += ck.n_b;
        double acc = 0.0;
        for (uint32_t i = 0; i < ck.n_c; ++i) {
            const unsigned char* w = rec_c0 + (size_t)i * kKoRecC;
            const uint32_t P = *reinterpret_cast<const uint32_t*>(w + lane4) ^ twp[i];
            const uint2 hdr = *reinterpret_cast<const uint2*>(w + 128);
            const unsigned char* L = w + 144;
#pragma unroll
            for (int j = 0; j < 5; ++j) {  // chunk sums in term order
                const uint32_t off = j == 0 ? ((P << 3) & 0x1f8u) : ((P >> (6 * j - 3)) & 0x1f8u);
                acc = __dadd_rn(acc, *reinterpret_cast<const double*>(L + j * 512 + off));
            }
            if (hdr.y) {  // last word of a blob (warp-uniform)
                const double h1[1] = {acc};
                const uint32_t u1[1] = {hdr.x};
                ko_emit<1>(h1, u1, a0, base_hi, valid, q, e_re, e_im);
                acc = 0.0;
            }
        }
        if (valid) partial[(int64_t)blockIdx.y * n_keys + key] = make_double2(e_re, e_im);
        __syncwarp();  // every lane is done with the TW strip before the next task overwrites it
    }
    if (!loaded) mbar_wait(&mbar, 0);  // never leave while the bulk copy into this CTA's shared memory is in flight
}

}  // namespace naqs
