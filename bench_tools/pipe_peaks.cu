// pipe_peaks.cu — measures the B200 ceilings the E_loc roofline needs and MEASURED_PEAKS.json lacks
// (SURVEY.md §8d): thread-op rates of POPC / LOP3 / IMAD / DADD, the issue-limited rate of the direct
// sign-accumulate sequence, broadcast-LDS rate, and random 16-byte gather rates from L2-resident and
// HBM-resident tables.  Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

__global__ void k_popc(uint32_t* out, uint32_t seed) {
    uint32_t x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 2654435761u + i + seed;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
    uint32_t s = 0;
    for (int i = 0; i < ILP; ++i) s ^= x[i];
    if (s == 0xdeadbeef) out[0] = s;
}
__global__ void k_lop3(uint32_t* out, uint32_t seed) {
    uint32_t x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 2654435761u + i + seed;
    const uint32_t a = seed * 3 + 1, b = seed * 7 + 5;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[i]) : "r"(a), "r"(b));
    uint32_t s = 0;
    for (int i = 0; i < ILP; ++i) s ^= x[i];
    if (s == 0xdeadbeef) out[0] = s;
}
__global__ void k_imad(uint32_t* out, uint32_t seed) {
    uint32_t x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 2654435761u + i + seed;
    const uint32_t a = seed * 3 + 1;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = x[i] * a + 12345u;
    uint32_t s = 0;
    for (int i = 0; i < ILP; ++i) s ^= x[i];
    if (s == 0xdeadbeef) out[0] = s;
}
__global__ void k_dadd(uint32_t* out, double c) {
    double x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __dadd_rn(x[i], c);
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 1.2345) out[0] = 1;
}
// the direct formulation's per-coupling sequence: AND, POPC, shift, XOR-into-high-word, DADD
__global__ void k_direct(uint32_t* out, uint32_t seed, double c) {
    uint32_t s[ILP]; double acc[ILP];
    for (int i = 0; i < ILP; ++i) { s[i] = threadIdx.x * 2654435761u + i + seed; acc[i] = 0; }
    uint32_t yz = seed * 97 + 13;
    const int hi = __double2hiint(c), lo = __double2loint(c);
    for (int it = 0; it < ITERS; ++it) {
        yz = yz * 1664525u + 1013904223u;  // one uniform op per ILP couplings (stands for the broadcast LDS)
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = __dadd_rn(acc[i], __hiloint2double(hi ^ (int)(__popc(s[i] & yz) << 31), lo));
    }
    double t = 0;
    for (int i = 0; i < ILP; ++i) t += acc[i];
    if (t == 1.2345) out[0] = 1;
}
// sign from a precomputed parity word (bit-sliced formulation): shift, XOR-into-high-word, DADD
__global__ void k_sliced(uint32_t* out, uint32_t seed, double c) {
    double acc[ILP]; uint32_t p = threadIdx.x * 2654435761u + seed;
    for (int i = 0; i < ILP; ++i) acc[i] = 0;
    const int hi = __double2hiint(c), lo = __double2loint(c);
    for (int it = 0; it < ITERS; ++it) {
        p = p * 1664525u + 1013904223u;
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = __dadd_rn(acc[i], __hiloint2double(hi ^ (int)((p << (31 - i)) & 0x80000000u), lo));
    }
    double t = 0;
    for (int i = 0; i < ILP; ++i) t += acc[i];
    if (t == 1.2345) out[0] = 1;
}
__global__ void k_lds_bcast(uint32_t* out, uint32_t seed) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
    uint32_t x = 0;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x ^= sm[(it * ILP + i) & 4095];   // same address for the whole warp
    if (x == 0xdeadbeef) out[0] = x;
}
__global__ void k_gather(const double2* __restrict__ tbl, uint64_t mask, uint32_t* out, int per_thread) {
    uint64_t h = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    double s = 0;
    for (int it = 0; it < per_thread; it += 4) {
        double2 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { h = h * 6364136223846793005ull + 1442695040888963407ull; v[j] = __ldg(tbl + ((h >> 20) & mask)); }
#pragma unroll
        for (int j = 0; j < 4; ++j) s += v[j].x + v[j].y;
    }
    if (s == 1.2345) out[0] = 1;
}

template <class F>
static float time_ms(F launch, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint32_t* d_out; CK(cudaMalloc(&d_out, 64));
    const int blocks = sms * 8, threads = 256;
    const double n_ops = (double)blocks * threads * ITERS * ILP;
    auto rate = [&](float ms) { return n_ops / (ms * 1e-3); };
    float t_popc = time_ms([&] { k_popc<<<blocks, threads>>>(d_out, 1); });
    float t_lop3 = time_ms([&] { k_lop3<<<blocks, threads>>>(d_out, 1); });
    float t_imad = time_ms([&] { k_imad<<<blocks, threads>>>(d_out, 1); });
    float t_dadd = time_ms([&] { k_dadd<<<blocks, threads>>>(d_out, 1e-3); });
    float t_dir = time_ms([&] { k_direct<<<blocks, threads>>>(d_out, 1, 0.37); });
    float t_sli = time_ms([&] { k_sliced<<<blocks, threads>>>(d_out, 1, 0.37); });
    float t_lds = time_ms([&] { k_lds_bcast<<<blocks, threads>>>(d_out, 3); });
    CK(cudaGetLastError());
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", prop.name, sms, prop.clockRate);
    printf(" \"popc_ops_per_s\": %.4e, \"lop3_ops_per_s\": %.4e, \"imad_ops_per_s\": %.4e, \"dadd_ops_per_s\": %.4e,\n",
           rate(t_popc), rate(t_lop3), rate(t_imad), rate(t_dadd));
    printf(" \"direct_couplings_per_s\": %.4e, \"sliced_sign_dadd_per_s\": %.4e, \"lds_broadcast_per_s\": %.4e,\n",
           rate(t_dir), rate(t_sli), rate(t_lds));
    // random 16 B gathers
    const size_t sizes_mb[3] = {16, 64, 4096};
    for (int s = 0; s < 3; ++s) {
        const size_t entries = sizes_mb[s] * 1024 * 1024 / 16;
        double2* tbl; CK(cudaMalloc(&tbl, entries * 16)); CK(cudaMemset(tbl, 0, entries * 16));
        const int per_thread = 256, gb = sms * 16;
        float t = time_ms([&] { k_gather<<<gb, 256>>>(tbl, entries - 1, d_out, per_thread); });
        printf(" \"gather16B_%zuMB_per_s\": %.4e,\n", sizes_mb[s], (double)gb * 256 * per_thread / (t * 1e-3));
        cudaFree(tbl);
    }
    printf(" \"iters\": %d}\n", ITERS);
    return 0;
}
