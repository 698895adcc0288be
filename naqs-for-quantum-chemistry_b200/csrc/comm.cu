// comm.cu — the exchange steps of the sharded local energy behind the C ABI: one process per GPU.
//
// The reference has no distributed code; the path needs exactly two exchanges per step:
//   (1) every rank must see the (key, psi) pairs of ALL ranks before it can walk its own rows  -> naqs_table_exchange
//   (2) the five fp64 statistics sums of the loss are global                                  -> naqs_stats_allreduce
// Both are written as PUSH kernels over peer memory — except that DENSE shards (a rank's pairs alone would fill a good part
// of the direct-address table, e.g. 10^6 rows of a 2^20 key space per rank) go through an all-reduce of the table instead: a
// push delivers every pair to every peer (volume ~ n_local * world), the all-reduce moves the table once whatever the
// number of ranks and can be reduced inside the NVSwitch.  The statistics always use the push kernel.
// Push kernels run over CUDA IPC mappings of one region per rank (NVLink / NVSwitch underneath): a rank stores its own contribution straight into every peer's buffer, raises a flag there, and waits for the flags of its
// peers — no reduction is needed (copies of a key carry the same amplitude by contract, and the sums are added locally in
// rank order, so the result is bitwise identical on every rank).  One kernel per exchange, no host synchronisation, no NCCL
// on the data path.  NCCL (loaded at run time from the process, the library torch.distributed uses) does the plumbing:
// the one-time all-gather of the IPC handles, and the all-gather of (key, psi) for key spaces too large for a
// direct-address table (hash lookup).
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace naqs {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi& nccl_api() {
    static NcclApi api;
    if (api.ok || api.lib) return api;
    // the copy already mapped into the process (torch's) first, then whatever the loader finds
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
        if (api.lib) break;
    }
    if (!api.lib)
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
    if (!api.lib) return api;
    auto sym = [&](const char* n) { return dlsym(api.lib, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce && api.GetErrorString;
    return api;
}

#define NAQS_NCCL(call)                                                                              \
    do {                                                                                             \
        ncclResult_t r_ = (call);                                                                    \
        if (r_ != ncclSuccess) {                                                                     \
            naqs::set_error(std::string(#call) + ": " + naqs::nccl_api().GetErrorString(r_));        \
            return NAQS_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)

constexpr int kMaxRanks = 64;
// Region layout (identical on every rank except for the table offsets, which each rank aligns in its own address space):
//   [0, 1024)            flags: int32 [4 kinds][kMaxRanks] — kind 0 table push, kind 1 statistics, kind 2 "my pairs are in my table",
//                        kind 3 "my slice of the merged table is in every table"; flag[k][r] = last epoch rank r completed
//   [1024, 1024 + 8192)  statistics slots: double [2 parities][kMaxRanks][8]
//   [16384, ...)         three direct-address complex64 tables (2^N entries each), each aligned to its size: two alternate
//                        between the epochs of the push exchange, the third belongs to the all-reduce exchange
constexpr size_t kFlagsOff = 0, kStatsOff = 1024, kTablesOff = 16384;

struct PeerInfo {
    cudaIpcMemHandle_t handle;
    unsigned long long table_off[2];
};

}  // namespace naqs

struct naqs_comm {
    int world = 1, rank = 0, device = 0;
    ncclComm_t nccl = nullptr;
    bool own_nccl = false;
    char* region = nullptr;         // this rank's region
    size_t region_bytes = 0;
    int64_t table_entries = 0;
    std::vector<char*> peer;        // mapped base of every rank's region; peer[rank] == region
    std::vector<naqs::PeerInfo> info;
    char** d_peer = nullptr;        // device copies for the kernels
    unsigned long long* d_table_off = nullptr;  // [world][2]
    int* d_done = nullptr;          // CTA completion counter of the push kernel
    unsigned epoch_table = 0, epoch_stats = 0;
    // NCCL path (hash lookup): gather buffers
    void* d_gather = nullptr;
    size_t gather_bytes = 0;
    // push-gather path (hash lookup, the default): a second peer-mapped region holding, per parity, the (key, psi) slots of all ranks
    char* gregion = nullptr;
    size_t gregion_bytes = 0;
    int64_t g_cap = 0;              // entries per rank slot
    size_t g_kb = 0, g_pb = 0;      // bytes per key / per amplitude the region was laid out for
    std::vector<char*> gpeer;
    char** d_gpeer = nullptr;
    int* d_gdone = nullptr;
    unsigned epoch_gather = 0;
    // peer mapping (CUDA IPC) is probed once, collectively; without it every exchange takes its NCCL form
    int ipc_state = 0;              // 0 = not probed, 1 = available on every rank, -1 = unavailable (or NAQS_COMM_NO_IPC)
    void* d_local_tbl_raw = nullptr;  // NCCL all-reduce form without a peer region: a local table aligned to its size
    float2* d_local_tbl = nullptr;
    int64_t local_tbl_entries = 0;
};

namespace naqs {

__device__ __forceinline__ int ld_flag(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Every rank stores its (key, psi) pairs into the current table of EVERY rank (its own included), then the last CTA to finish
// publishes "rank `rank` completed epoch e" in every peer's flag row and waits until all peers have published e in its own.
__global__ void push_table_kernel(char* const* __restrict__ peer, const unsigned long long* __restrict__ table_off, int world, int rank, int parity,
                                  const uint64_t* __restrict__ keys, const float2* __restrict__ psi, int64_t n, int64_t entries, int epoch,
                                  int* __restrict__ done, int* __restrict__ err_flags) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        if (k >= (unsigned long long)entries) { if (err_flags) atomicOr(err_flags, 1); continue; }
        const float2 v = psi[i];
        for (int r = 0; r < world; ++r) {
            const int rr = (rank + r) % world;  // start with the own table, spread the peers
            reinterpret_cast<float2*>(peer[rr] + table_off[2 * rr + parity])[k] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = atomicAdd(done, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *done = 0;
    __threadfence_system();
    if ((int)threadIdx.x < world) {
        st_flag(reinterpret_cast<int*>(peer[threadIdx.x] + kFlagsOff) + rank, epoch);                       // tell peer threadIdx.x
        const int* mine = reinterpret_cast<const int*>(peer[rank] + kFlagsOff) + threadIdx.x;
        while (ld_flag(mine) - epoch < 0) { }                                                               // wait for peer threadIdx.x
    }
}

// All-reduce (sum) of the five statistics by push: slot [parity][rank] of every peer receives this rank's sums; after all flags
// of the epoch have arrived the slots are added in rank order (identical result on every rank).
__global__ void push_stats_kernel(char* const* __restrict__ peer, int world, int rank, int parity, double* __restrict__ sums5, int epoch) {
    const int t = threadIdx.x;
    if (t < world) {
        double* slot = reinterpret_cast<double*>(peer[t] + kStatsOff) + ((size_t)parity * kMaxRanks + rank) * 8;
        for (int j = 0; j < 5; ++j) slot[j] = sums5[j];
        __threadfence_system();
        st_flag(reinterpret_cast<int*>(peer[t] + kFlagsOff) + kMaxRanks + rank, epoch);
        const int* mine = reinterpret_cast<const int*>(peer[rank] + kFlagsOff) + kMaxRanks + t;
        while (ld_flag(mine) - epoch < 0) { }
    }
    __syncthreads();
    if (t < 5) {
        const volatile double* slots = reinterpret_cast<const volatile double*>(peer[rank] + kStatsOff) + (size_t)parity * kMaxRanks * 8;
        double acc = 0.0;
        for (int r = 0; r < world; ++r) acc += slots[r * 8 + t];
        sums5[t] = acc;
    }
}

// Dense shards (a rank's pairs alone fill a good part of the table): a push would deliver every pair to every peer (volume
// ~ n_local * world); instead the tables are MERGED like a two-shot all-reduce written over peer memory.  Every rank scatters
// its pairs into its OWN zeroed table and publishes flag kind 2 (scatter_signal_kernel).  Rank r then owns slice r of the key
// space: it ORs slice r of every peer's table into its own (P2P vector loads over NVLink) and stores the merged slice back
// into every peer's table (P2P stores), publishes flag kind 3 and waits for the kind-3 flags of its peers
// (merge_table_kernel).  Bitwise OR is the exact merge: an absent entry is all-zero bits, and copies of a key on several
// ranks carry the same amplitude by contract.  Volume per rank: 2 * (world - 1) / world tables whatever the rank count; two
// flag round trips instead of a collective call.
__global__ void scatter_signal_kernel(char* const* __restrict__ peer, const unsigned long long* __restrict__ table_off, int world, int rank, int parity,
                                      const uint64_t* __restrict__ keys, const float2* __restrict__ psi, int64_t n, int64_t entries, int epoch,
                                      int* __restrict__ done, int* __restrict__ err_flags) {
    float2* own = reinterpret_cast<float2*>(peer[rank] + table_off[2 * rank + parity]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        if (k >= (unsigned long long)entries) { if (err_flags) atomicOr(err_flags, 1); continue; }
        own[k] = psi[i];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = atomicAdd(done, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *done = 0;
    __threadfence_system();
    if ((int)threadIdx.x < world) st_flag(reinterpret_cast<int*>(peer[threadIdx.x] + kFlagsOff) + 2 * kMaxRanks + rank, epoch);
}

__device__ __forceinline__ uint4 ld_peer_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int WORLD>  // 0: run-time rank count
__global__ void __launch_bounds__(256) merge_table_kernel(char* const* __restrict__ peer, const unsigned long long* __restrict__ table_off, int world_rt, int rank,
                                                          int parity, int64_t n_vec, int epoch, int* __restrict__ done) {
    const int world = WORLD ? WORLD : world_rt;
    __shared__ uint4* tab[kMaxRanks];
    if ((int)threadIdx.x < world) {
        tab[threadIdx.x] = reinterpret_cast<uint4*>(peer[threadIdx.x] + table_off[2 * threadIdx.x + parity]);
        const int* mine = reinterpret_cast<const int*>(peer[rank] + kFlagsOff) + 2 * kMaxRanks + threadIdx.x;
        while (ld_flag(mine) - epoch < 0) { }   // peer threadIdx.x has scattered its pairs of this epoch
    }
    __syncthreads();
    const int64_t lo = n_vec * rank / world, hi = n_vec * (rank + 1) / world;
    uint4* const own = tab[rank];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (WORLD) {
        // BATCH vectors per thread and pass: every P2P load of the pass is in flight before the first one is consumed
        constexpr int BATCH = WORLD <= 4 ? 4 : 2;
        for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += BATCH * stride) {
            uint4 v[BATCH], in[BATCH][WORLD > 1 ? WORLD - 1 : 1];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const int64_t i = i0 + b * stride;
                if (i < hi) {
#pragma unroll
                    for (int r = 1; r < WORLD; ++r) in[b][r - 1] = ld_peer_v4(tab[(rank + r) % WORLD] + i);
                    v[b] = own[i];
                }
            }
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const int64_t i = i0 + b * stride;
                if (i < hi) {
#pragma unroll
                    for (int r = 1; r < WORLD; ++r) { v[b].x |= in[b][r - 1].x; v[b].y |= in[b][r - 1].y; v[b].z |= in[b][r - 1].z; v[b].w |= in[b][r - 1].w; }
                    own[i] = v[b];
#pragma unroll
                    for (int r = 1; r < WORLD; ++r) tab[(rank + r) % WORLD][i] = v[b];
                }
            }
        }
    } else {
        for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
            uint4 v = own[i];
            for (int r = 1; r < world; ++r) {
                const uint4 w = ld_peer_v4(tab[(rank + r) % world] + i);
                v.x |= w.x; v.y |= w.y; v.z |= w.z; v.w |= w.w;
            }
            own[i] = v;
            for (int r = 1; r < world; ++r) tab[(rank + r) % world][i] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = atomicAdd(done, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *done = 0;
    __threadfence_system();
    if ((int)threadIdx.x < world) {
        st_flag(reinterpret_cast<int*>(peer[threadIdx.x] + kFlagsOff) + 3 * kMaxRanks + rank, epoch);
        const int* mine = reinterpret_cast<const int*>(peer[rank] + kFlagsOff) + 3 * kMaxRanks + threadIdx.x;
        while (ld_flag(mine) - epoch < 0) { }   // every slice of the merged table has landed here
    }
}

// dense shards: the table is pre-filled with -0.0f (bit pattern INT32_MIN: "absent" for a MAX reduction on the int32 patterns
// and a numeric zero for the kernel), every rank scatters its own pairs, and the tables are all-reduced with MAX — copies of a
// key on several ranks carry the same amplitude by contract, so MAX simply keeps it.
__global__ void fill_absent_kernel(int4* __restrict__ table, int64_t n_vec) {
    const int v = (int)0x80000000u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) table[i] = make_int4(v, v, v, v);
}
__global__ void scatter_table_kernel(float2* __restrict__ table, const uint64_t* __restrict__ keys, const float2* __restrict__ psi, int64_t n, int64_t entries,
                                     int* __restrict__ err_flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    if (k >= (unsigned long long)entries) { if (err_flags) atomicOr(err_flags, 1); return; }
    table[k] = psi[i];
}

static int comm_sync_barrier(naqs_comm* c, cudaStream_t st) {  // host-visible barrier through NCCL (setup only)
    int* d = nullptr;
    NAQS_CUDA(cudaMalloc((void**)&d, sizeof(int)));
    NAQS_CUDA(cudaMemsetAsync(d, 0, sizeof(int), st));
    NAQS_NCCL(nccl_api().AllReduce(d, d, 1, ncclInt32, ncclSum, c->nccl, st));
    NAQS_CUDA(cudaStreamSynchronize(st));
    cudaFree(d);
    return NAQS_OK;
}

// One-time collective probe: can every rank map every peer's memory through CUDA IPC?  (Containers without a shared IPC namespace,
// GPUs without peer access.)  The answer is all-reduced (MIN), so all ranks take the same exchange forms afterwards.
static int ipc_probe(naqs_comm* c, cudaStream_t st) {
    if (c->ipc_state != 0) return NAQS_OK;
    int ok = getenv("NAQS_COMM_NO_IPC") ? 0 : 1;
    void* buf = nullptr;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (ok && (cudaMalloc(&buf, 1 << 16) != cudaSuccess || cudaIpcGetMemHandle(&mine, buf) != cudaSuccess)) { ok = 0; cudaGetLastError(); }
    cudaIpcMemHandle_t* d_h = nullptr;
    NAQS_CUDA(cudaMalloc((void**)&d_h, sizeof(mine) * c->world));
    NAQS_CUDA(cudaMemcpyAsync(d_h + c->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    NAQS_NCCL(nccl_api().AllGather(d_h + c->rank, d_h, sizeof(mine), ncclUint8, c->nccl, st));
    std::vector<cudaIpcMemHandle_t> all((size_t)c->world);
    NAQS_CUDA(cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * c->world, cudaMemcpyDeviceToHost, st));
    int* d_ok = nullptr;
    NAQS_CUDA(cudaMalloc((void**)&d_ok, sizeof(int)));
    NAQS_CUDA(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
    NAQS_NCCL(nccl_api().AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, c->nccl, st));   // did every rank export a handle?
    int all_exported = 0;
    NAQS_CUDA(cudaMemcpyAsync(&all_exported, d_ok, sizeof(int), cudaMemcpyDeviceToHost, st));
    NAQS_CUDA(cudaStreamSynchronize(st));
    ok = all_exported;
    std::vector<void*> opened;
    if (ok)
        for (int r = 0; r < c->world && ok; ++r) {
            if (r == c->rank) continue;
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
            else opened.push_back(p);
        }
    NAQS_CUDA(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
    NAQS_NCCL(nccl_api().AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, c->nccl, st));   // could every rank map every peer?
    NAQS_CUDA(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, st));
    NAQS_CUDA(cudaStreamSynchronize(st));
    for (void* p : opened) cudaIpcCloseMemHandle(p);
    int rc = comm_sync_barrier(c, st);   // every mapping is closed before the probe buffers are freed
    cudaFree(d_h); cudaFree(d_ok); cudaFree(buf);
    c->ipc_state = ok ? 1 : -1;
    return rc;
}

// (Re)create the peer-mapped region for direct-address tables of `entries` complex64 values.  Collective and synchronous:
// called the first time a table of this size is exchanged.
static int ensure_region(naqs_comm* c, int64_t entries, cudaStream_t st) {
    if (c->region && c->table_entries == entries) return NAQS_OK;
    NAQS_REQUIRE(!c->region, NAQS_ERR_STATE, "naqs_table_exchange: the communicator is already bound to a table of another size");
    const size_t tbytes = (size_t)entries * sizeof(float2);
    const size_t bytes = kTablesOff + 4 * tbytes;  // room to align both tables to their size
    NAQS_CUDA(cudaMalloc((void**)&c->region, bytes));
    NAQS_CUDA(cudaMemsetAsync(c->region, 0, bytes, st));
    c->region_bytes = bytes;
    c->table_entries = entries;
    PeerInfo mine;
    std::memset(&mine, 0, sizeof(mine));
    NAQS_CUDA(cudaIpcGetMemHandle(&mine.handle, c->region));
    const uintptr_t base = (uintptr_t)c->region;
    const uintptr_t t0 = (base + kTablesOff + tbytes - 1) & ~(uintptr_t)(tbytes - 1);
    mine.table_off[0] = (unsigned long long)(t0 - base);
    mine.table_off[1] = mine.table_off[0] + tbytes;
    // all-gather the descriptors through NCCL (bytes)
    PeerInfo* d_info = nullptr;
    NAQS_CUDA(cudaMalloc((void**)&d_info, sizeof(PeerInfo) * c->world));
    NAQS_CUDA(cudaMemcpyAsync(d_info + c->rank, &mine, sizeof(PeerInfo), cudaMemcpyHostToDevice, st));
    NAQS_NCCL(nccl_api().AllGather(d_info + c->rank, d_info, sizeof(PeerInfo), ncclUint8, c->nccl, st));
    c->info.resize((size_t)c->world);
    NAQS_CUDA(cudaMemcpyAsync(c->info.data(), d_info, sizeof(PeerInfo) * c->world, cudaMemcpyDeviceToHost, st));
    NAQS_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_info);
    c->peer.assign((size_t)c->world, nullptr);
    std::vector<unsigned long long> offs((size_t)c->world * 2);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) c->peer[(size_t)r] = c->region;
        else {
            void* p = nullptr;
            NAQS_CUDA(cudaIpcOpenMemHandle(&p, c->info[(size_t)r].handle, cudaIpcMemLazyEnablePeerAccess));
            c->peer[(size_t)r] = static_cast<char*>(p);
        }
        offs[2 * (size_t)r] = c->info[(size_t)r].table_off[0];
        offs[2 * (size_t)r + 1] = c->info[(size_t)r].table_off[1];
    }
    NAQS_CUDA(cudaMalloc((void**)&c->d_peer, sizeof(char*) * c->world));
    NAQS_CUDA(cudaMalloc((void**)&c->d_table_off, sizeof(unsigned long long) * 2 * c->world));
    NAQS_CUDA(cudaMalloc((void**)&c->d_done, sizeof(int)));
    NAQS_CUDA(cudaMemcpyAsync(c->d_peer, c->peer.data(), sizeof(char*) * c->world, cudaMemcpyHostToDevice, st));
    NAQS_CUDA(cudaMemcpyAsync(c->d_table_off, offs.data(), sizeof(unsigned long long) * 2 * c->world, cudaMemcpyHostToDevice, st));
    NAQS_CUDA(cudaMemsetAsync(c->d_done, 0, sizeof(int), st));
    return comm_sync_barrier(c, st);  // every region is zeroed and mapped before anyone pushes
}

// Large key spaces (hash lookup): every rank stores its (key, psi) pairs — padded to the slot capacity with all-ones keys, which
// the lookup build skips — straight into its slot of EVERY rank's gather buffer, publishes flag kind 0 and waits for its peers:
// the all-gather of the shards as one push kernel over peer memory (no NCCL call, no staging copies).
// Layout per parity: keys [world][cap][kw] u64 | psi [world][cap][pw] u64.
__global__ void push_pairs_kernel(char* const* __restrict__ peer, int world, int rank, unsigned long long keys_off, unsigned long long psi_off, int64_t cap,
                                  int kw, int pw, const uint64_t* __restrict__ keys, const uint64_t* __restrict__ psi, int64_t n_local, int epoch,
                                  int* __restrict__ done) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t k[2] = {~0ull, ~0ull}, v[2] = {0ull, 0ull};
        if (i < n_local) {
            for (int w = 0; w < kw; ++w) k[w] = keys[i * kw + w];
            for (int w = 0; w < pw; ++w) v[w] = psi[i * pw + w];
        }
        for (int r = 0; r < world; ++r) {
            const int rr = (rank + r) % world;
            uint64_t* dk = reinterpret_cast<uint64_t*>(peer[rr] + keys_off) + ((int64_t)rank * cap + i) * kw;
            uint64_t* dv = reinterpret_cast<uint64_t*>(peer[rr] + psi_off) + ((int64_t)rank * cap + i) * pw;
            for (int w = 0; w < kw; ++w) dk[w] = k[w];
            for (int w = 0; w < pw; ++w) dv[w] = v[w];
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = atomicAdd(done, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *done = 0;
    __threadfence_system();
    if ((int)threadIdx.x < world) {
        st_flag(reinterpret_cast<int*>(peer[threadIdx.x] + kFlagsOff) + rank, epoch);
        const int* mine = reinterpret_cast<const int*>(peer[rank] + kFlagsOff) + threadIdx.x;
        while (ld_flag(mine) - epoch < 0) { }
    }
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// (Re)create the gather region for slots of `cap` entries.  Collective and synchronous (first exchange, or the batch outgrew it).
static int ensure_gather_region(naqs_comm* c, int64_t max_local, size_t kb, size_t pb, cudaStream_t st) {
    if (c->gregion && c->g_cap >= max_local && c->g_kb == kb && c->g_pb == pb) return NAQS_OK;
    if (c->gregion) {  // every rank arrives here in the same call (max_local is the same everywhere): unmap, then free
        NAQS_CUDA(cudaStreamSynchronize(st));
        int rc = comm_sync_barrier(c, st);
        if (rc) return rc;
        for (int r = 0; r < (int)c->gpeer.size(); ++r)
            if (r != c->rank && c->gpeer[(size_t)r]) cudaIpcCloseMemHandle(c->gpeer[(size_t)r]);
        c->gpeer.clear();
        rc = comm_sync_barrier(c, st);
        if (rc) return rc;
        cudaFree(c->gregion); c->gregion = nullptr;
        cudaFree(c->d_gpeer); c->d_gpeer = nullptr;
    }
    const int64_t cap = std::max<int64_t>(1024, max_local + max_local / 4);
    const size_t per_parity = al256((size_t)c->world * cap * kb) + al256((size_t)c->world * cap * pb);
    const size_t bytes = kTablesOff + 2 * per_parity;
    NAQS_CUDA(cudaMalloc((void**)&c->gregion, bytes));
    NAQS_CUDA(cudaMemsetAsync(c->gregion, 0, kTablesOff, st));
    c->gregion_bytes = bytes; c->g_cap = cap; c->g_kb = kb; c->g_pb = pb;
    c->epoch_gather = 0;
    cudaIpcMemHandle_t mine;
    NAQS_CUDA(cudaIpcGetMemHandle(&mine, c->gregion));
    cudaIpcMemHandle_t* d_h = nullptr;
    NAQS_CUDA(cudaMalloc((void**)&d_h, sizeof(mine) * c->world));
    NAQS_CUDA(cudaMemcpyAsync(d_h + c->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    NAQS_NCCL(nccl_api().AllGather(d_h + c->rank, d_h, sizeof(mine), ncclUint8, c->nccl, st));
    std::vector<cudaIpcMemHandle_t> all((size_t)c->world);
    NAQS_CUDA(cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * c->world, cudaMemcpyDeviceToHost, st));
    NAQS_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_h);
    c->gpeer.assign((size_t)c->world, nullptr);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) { c->gpeer[(size_t)r] = c->gregion; continue; }
        void* p = nullptr;
        NAQS_CUDA(cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess));
        c->gpeer[(size_t)r] = static_cast<char*>(p);
    }
    NAQS_CUDA(cudaMalloc((void**)&c->d_gpeer, sizeof(char*) * c->world));
    NAQS_CUDA(cudaMemcpyAsync(c->d_gpeer, c->gpeer.data(), sizeof(char*) * c->world, cudaMemcpyHostToDevice, st));
    if (!c->d_gdone) {
        NAQS_CUDA(cudaMalloc((void**)&c->d_gdone, sizeof(int)));
        NAQS_CUDA(cudaMemsetAsync(c->d_gdone, 0, sizeof(int), st));
    }
    return comm_sync_barrier(c, st);  // every region is zeroed and mapped before anyone pushes
}

}  // namespace naqs

using namespace naqs;

extern "C" {

int naqs_comm_unique_id(void* id128) {
    NAQS_REQUIRE(id128, NAQS_ERR_ARG, "naqs_comm_unique_id: NULL buffer");
    NAQS_REQUIRE(nccl_api().ok, NAQS_ERR_STATE, "naqs_comm: libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == 128, "NAQS_COMM_ID_BYTES");
    ncclUniqueId id;
    NAQS_NCCL(nccl_api().GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return NAQS_OK;
}

int naqs_comm_init(naqs_comm_t** out, const void* id128, int world, int rank, int device) {
    NAQS_REQUIRE(out && id128 && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, NAQS_ERR_ARG, "naqs_comm_init: bad arguments (world <= 64)");
    NAQS_REQUIRE(nccl_api().ok, NAQS_ERR_STATE, "naqs_comm: libnccl.so.2 could not be loaded");
    DeviceGuard guard(device);
    NAQS_REQUIRE(guard.ok, NAQS_ERR_CUDA, "naqs_comm_init: cannot select the CUDA device (no CPU fallback)");
    naqs_comm* c = new naqs_comm;
    c->world = world; c->rank = rank; c->device = device;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclResult_t r = nccl_api().CommInitRank(&c->nccl, world, id, rank);
    if (r != ncclSuccess) {
        set_error(std::string("ncclCommInitRank: ") + nccl_api().GetErrorString(r));
        delete c;
        return NAQS_ERR_CUDA;
    }
    c->own_nccl = true;
    *out = c;
    return NAQS_OK;
}

int naqs_comm_from_nccl(naqs_comm_t** out, void* nccl_comm, int world, int rank, int device) {
    NAQS_REQUIRE(out && nccl_comm && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, NAQS_ERR_ARG, "naqs_comm_from_nccl: bad arguments");
    NAQS_REQUIRE(nccl_api().ok, NAQS_ERR_STATE, "naqs_comm: libnccl.so.2 could not be loaded");
    naqs_comm* c = new naqs_comm;
    c->world = world; c->rank = rank; c->device = device;
    c->nccl = static_cast<ncclComm_t>(nccl_comm);
    c->own_nccl = false;
    *out = c;
    return NAQS_OK;
}

int naqs_comm_destroy(naqs_comm_t* c) {
    if (!c) return NAQS_OK;
    DeviceGuard guard(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < (int)c->peer.size(); ++r)
        if (r != c->rank && c->peer[(size_t)r]) cudaIpcCloseMemHandle(c->peer[(size_t)r]);
    for (int r = 0; r < (int)c->gpeer.size(); ++r)
        if (r != c->rank && c->gpeer[(size_t)r]) cudaIpcCloseMemHandle(c->gpeer[(size_t)r]);
    cudaFree(c->region); cudaFree(c->d_peer); cudaFree(c->d_table_off); cudaFree(c->d_done); cudaFree(c->d_gather);
    cudaFree(c->gregion); cudaFree(c->d_gpeer); cudaFree(c->d_gdone); cudaFree(c->d_local_tbl_raw);
    if (c->own_nccl && c->nccl) nccl_api().CommDestroy(c->nccl);
    delete c;
    return NAQS_OK;
}

int naqs_comm_info(const naqs_comm_t* c, int* world, int* rank) {
    NAQS_REQUIRE(c, NAQS_ERR_ARG, "naqs_comm_info: NULL communicator");
    if (world) *world = c->world;
    if (rank) *rank = c->rank;
    return NAQS_OK;
}

int naqs_table_exchange(naqs_table_t* t, naqs_comm_t* c, const uint64_t* d_keys, const void* d_psi, int psi_dtype, int64_t n_local,
                        int64_t max_local, int flags, void* stream_) {
    NAQS_REQUIRE(t && c, NAQS_ERR_ARG, "naqs_table_exchange: NULL argument");
    NAQS_REQUIRE(n_local >= 0 && (n_local == 0 || (d_keys && d_psi)), NAQS_ERR_ARG, "naqs_table_exchange: NULL buffers");
    NAQS_REQUIRE(psi_dtype == NAQS_C128 || psi_dtype == NAQS_C64, NAQS_ERR_DTYPE, "naqs_table_exchange: psi must be complex64 or complex128");
    DeviceGuard guard(t->device);
    cudaStream_t st = (cudaStream_t)stream_;
    const bool dense_push = t->nw32 == 1 && t->n_qubits <= 22 && t->n_qubits >= 5 && psi_dtype == NAQS_C64 && t->algo == 0 && !(flags & NAQS_EXCHANGE_GATHER);
    if (dense_push) {
        const int64_t entries = 1ll << t->n_qubits;
        int rc = c->world > 1 ? ipc_probe(c, st) : NAQS_OK;
        if (rc) return rc;
        const bool no_ipc = c->world > 1 && c->ipc_state < 0;
        if (!no_ipc) {
            rc = ensure_region(c, entries, st);
            if (rc) return rc;
        }
        // dense shards -> merge of the tables over peer memory (NAQS_EXCHANGE_REDUCE: the NCCL all-reduce (MAX) of round 1 instead);
        // sparse shards -> push
        const bool dense_shards = c->world > 1 && !(flags & NAQS_EXCHANGE_PUSH) &&
                                  ((flags & (NAQS_EXCHANGE_REDUCE | NAQS_EXCHANGE_MERGE)) || (double)max_local * (c->world - 1) > 0.5 * (double)entries);
        const bool reduce = no_ipc || (dense_shards && (flags & NAQS_EXCHANGE_REDUCE) != 0);  // no peer mapping: always the NCCL form
        const bool merge = dense_shards && !reduce && entries >= 2 * (int64_t)c->world;
        if (reduce) {
            float2* tbl;
            if (no_ipc) {
                if (c->local_tbl_entries < entries) {
                    cudaFree(c->d_local_tbl_raw); c->d_local_tbl_raw = nullptr; c->d_local_tbl = nullptr; c->local_tbl_entries = 0;
                    const size_t tb = (size_t)entries * sizeof(float2);
                    NAQS_CUDA(cudaMalloc(&c->d_local_tbl_raw, 2 * tb));
                    c->d_local_tbl = reinterpret_cast<float2*>(((uintptr_t)c->d_local_tbl_raw + tb - 1) & ~(uintptr_t)(tb - 1));
                    c->local_tbl_entries = entries;
                }
                tbl = c->d_local_tbl;
            } else {
                tbl = reinterpret_cast<float2*>(c->region + c->info[(size_t)c->rank].table_off[0] + 2 * (size_t)entries * sizeof(float2));
            }
            fill_absent_kernel<<<4 * 148, 256, 0, st>>>(reinterpret_cast<int4*>(tbl), entries / 2);
            NAQS_LAUNCHED();
            if (n_local > 0) {
                scatter_table_kernel<<<(unsigned)((n_local + 255) / 256), 256, 0, st>>>(tbl, d_keys, reinterpret_cast<const float2*>(d_psi), n_local, entries, t->d_flags);
                NAQS_LAUNCHED();
            }
            NAQS_NCCL(nccl_api().AllReduce(tbl, tbl, (size_t)entries * 2, ncclInt32, ncclMax, c->nccl, st));
            return naqs_lookup_attach_dense32(t, reinterpret_cast<const float*>(tbl), entries);
        }
        const unsigned e = ++c->epoch_table;
        const int cur = (int)(e & 1u), nxt = cur ^ 1;
        // the table of the NEXT step is cleared now, before this rank signals epoch e: a peer only writes into it in step e + 1,
        // after it has seen that signal (see the header of this file)
        NAQS_CUDA(cudaMemsetAsync(c->region + c->info[(size_t)c->rank].table_off[nxt], 0, (size_t)entries * sizeof(float2), st));
        const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_local + 255) / 256, 4 * 148));
        const float* mine = reinterpret_cast<const float*>(c->region + c->info[(size_t)c->rank].table_off[cur]);
        if (merge) {
            scatter_signal_kernel<<<blocks, 256, 0, st>>>(c->d_peer, c->d_table_off, c->world, c->rank, cur, d_keys, reinterpret_cast<const float2*>(d_psi),
                                                          n_local, entries, (int)e, c->d_done, t->d_flags);
            NAQS_LAUNCHED();
            const int64_t n_vec = entries / 2;
            const unsigned mblocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_vec / c->world + 255) / 256, 2 * 148));
#define NAQS_MERGE(W) merge_table_kernel<W><<<mblocks, 256, 0, st>>>(c->d_peer, c->d_table_off, c->world, c->rank, cur, n_vec, (int)e, c->d_done)
            switch (c->world) {
                case 2: NAQS_MERGE(2); break;
                case 4: NAQS_MERGE(4); break;
                case 8: NAQS_MERGE(8); break;
                default: NAQS_MERGE(0); break;
            }
#undef NAQS_MERGE
            NAQS_LAUNCHED();
            return naqs_lookup_attach_dense32(t, mine, entries);
        }
        push_table_kernel<<<blocks, 256, 0, st>>>(c->d_peer, c->d_table_off, c->world, c->rank, cur, d_keys, reinterpret_cast<const float2*>(d_psi), n_local,
                                                  entries, (int)e, c->d_done, t->d_flags);
        NAQS_LAUNCHED();
        return naqs_lookup_attach_dense32(t, mine, entries);
    }
    // large key spaces: the shards are gathered by a push kernel over peer memory (push_pairs_kernel), then one lookup build over
    // the valid pairs; NAQS_EXCHANGE_GATHER selects the NCCL all-gather of round 1 instead
    NAQS_REQUIRE(max_local >= n_local, NAQS_ERR_ARG, "naqs_table_exchange: max_local must be the largest shard size of all ranks");
    if (c->world == 1)
        return naqs_lookup_build(t, d_keys, d_psi, psi_dtype, n_local, (flags & 0xff) | NAQS_LOOKUP_DUPLICATES_EQUAL, st);
    if (int rc_p = ipc_probe(c, st)) return rc_p;
    if (!(flags & NAQS_EXCHANGE_GATHER) && c->ipc_state > 0) {
        const size_t kb = (size_t)8 * t->words, pb = psi_dtype == NAQS_C64 ? 8 : 16;
        int rc = ensure_gather_region(c, max_local, kb, pb, st);
        if (rc) return rc;
        const unsigned e = ++c->epoch_gather;
        const size_t per_parity = al256((size_t)c->world * c->g_cap * kb) + al256((size_t)c->world * c->g_cap * pb);
        const size_t keys_off = kTablesOff + (size_t)(e & 1u) * per_parity, psi_off = keys_off + al256((size_t)c->world * c->g_cap * kb);
        const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((c->g_cap + 255) / 256, 2 * 148));
        push_pairs_kernel<<<blocks, 256, 0, st>>>(c->d_gpeer, c->world, c->rank, keys_off, psi_off, c->g_cap, (int)(kb / 8), (int)(pb / 8), d_keys,
                                                  reinterpret_cast<const uint64_t*>(d_psi), n_local, (int)e, c->d_gdone);
        NAQS_LAUNCHED();
        t->quiet_range_flag = true;  // padding keys are out of range on purpose
        rc = naqs_lookup_build(t, reinterpret_cast<const uint64_t*>(c->gregion + keys_off), c->gregion + psi_off, psi_dtype, (int64_t)c->world * c->g_cap,
                               (flags & 0xff) | NAQS_LOOKUP_DUPLICATES_EQUAL, st);
        t->quiet_range_flag = false;
        return rc;
    }
    // NCCL all-gather of equally sized (padded) shards, then one lookup build over the valid pairs.
    // max_local = the largest shard (every rank passes the same value); a shorter shard is padded with an out-of-range key,
    // which the build kernels skip (NAQS_EXCHANGE_PADDED tells naqs_table_check not to report it).
    NAQS_REQUIRE(max_local >= n_local, NAQS_ERR_ARG, "naqs_table_exchange: max_local must be the largest shard size of all ranks");
    const size_t kb = (size_t)8 * t->words, pb = psi_dtype == NAQS_C64 ? 8 : 16;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_k = 0, o_p = al((size_t)c->world * max_local * kb), total = o_p + al((size_t)c->world * max_local * pb);
    if (c->gather_bytes < total) {
        cudaFree(c->d_gather); c->d_gather = nullptr; c->gather_bytes = 0;
        NAQS_CUDA(cudaMalloc(&c->d_gather, total));
        c->gather_bytes = total;
    }
    char* g = static_cast<char*>(c->d_gather);
    char* my_k = g + o_k + (size_t)c->rank * max_local * kb;
    char* my_p = g + o_p + (size_t)c->rank * max_local * pb;
    if (n_local < max_local) {  // padding: all-ones keys (>= 2^n_qubits for every supported width) with zero amplitude
        NAQS_CUDA(cudaMemsetAsync(my_k + (size_t)n_local * kb, 0xff, (size_t)(max_local - n_local) * kb, st));
        NAQS_CUDA(cudaMemsetAsync(my_p + (size_t)n_local * pb, 0, (size_t)(max_local - n_local) * pb, st));
    }
    if (n_local > 0) {
        NAQS_CUDA(cudaMemcpyAsync(my_k, d_keys, (size_t)n_local * kb, cudaMemcpyDeviceToDevice, st));
        NAQS_CUDA(cudaMemcpyAsync(my_p, d_psi, (size_t)n_local * pb, cudaMemcpyDeviceToDevice, st));
    }
    if (c->world > 1) {
        NAQS_NCCL(nccl_api().AllGather(my_k, g + o_k, (size_t)max_local * kb, ncclUint8, c->nccl, st));
        NAQS_NCCL(nccl_api().AllGather(my_p, g + o_p, (size_t)max_local * pb, ncclUint8, c->nccl, st));
    }
    t->quiet_range_flag = true;  // padding keys are out of range on purpose
    int rc = naqs_lookup_build(t, reinterpret_cast<const uint64_t*>(g + o_k), g + o_p, psi_dtype, (int64_t)c->world * max_local,
                               (flags & 0xff) | NAQS_LOOKUP_DUPLICATES_EQUAL, st);
    t->quiet_range_flag = false;
    return rc;
}

int naqs_stats_allreduce(naqs_comm_t* c, double* d_sums5, void* stream_) {
    NAQS_REQUIRE(c && d_sums5, NAQS_ERR_ARG, "naqs_stats_allreduce: NULL argument");
    if (c->world == 1) return NAQS_OK;
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)stream_;
    if (c->region || c->gregion) {  // a peer-mapped region is available: one push kernel (flags / slots of that region)
        const unsigned e = ++c->epoch_stats;
        push_stats_kernel<<<1, 64, 0, st>>>(c->region ? c->d_peer : c->d_gpeer, c->world, c->rank, (int)(e & 1u), d_sums5, (int)e);
        NAQS_LAUNCHED();
        return NAQS_OK;
    }
    NAQS_NCCL(nccl_api().AllReduce(d_sums5, d_sums5, 5, ncclDouble, ncclSum, c->nccl, st));
    return NAQS_OK;
}

}  // extern "C"
