/* naqs_eloc.h — C ABI of the B200-native NAQS local-energy (E_loc) hot path.
 *
 * Drop-in boundary for the three compiled extension modules the reference imports by name
 * (src_cpp/setup.py:36-38: src.utils.hamiltonian_math / sparse_math / hilbert_math) and for the
 * numpy/scipy orchestration around them (src/optimizer/hamiltonian.py:272-370,
 * src/optimizer/energy.py:219-263).  All paths below are relative to the reference root.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / numpy / C++ types.  Every entry returns an int status
 *     (NAQS_OK = 0); naqs_last_error() gives the message of the calling thread's last failure.
 *     The Python shims map NAQS_ERR_DTYPE -> TypeError (hamiltonian_math.pyx:484) and the other
 *     codes -> RuntimeError / ValueError.
 *   - `d_` pointers are device memory of the table's device; `h_` pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous
 *     and stream-ordered unless stated otherwise; the caller owns every buffer it passes, the
 *     library owns only the table handle and its internal workspace.
 *   - a state / mask key is `words` consecutive uint64 (word 0 = qubits 0..63); bit q set <=> qubit q
 *     occupied (src/utils/hilbert.py:425,576-577).  words = 1 for n_qubits <= 63, 2 for <= 127.
 *   - complex numbers are interleaved (re, im) doubles (numpy complex128 / torch [...,2] float64).
 *   - there is NO CPU fallback anywhere in this library: without a CUDA device every compute entry
 *     fails with NAQS_ERR_CUDA.
 */
#ifndef NAQS_ELOC_H
#define NAQS_ELOC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAQS_ABI_VERSION 1

enum {
    NAQS_OK = 0,
    NAQS_ERR_ARG = 1,    /* bad argument (null pointer, size, unsupported width)            */
    NAQS_ERR_DTYPE = 2,  /* unsupported element type (-> TypeError in the Python shims)     */
    NAQS_ERR_CUDA = 3,   /* CUDA runtime / launch failure, or no device                    */
    NAQS_ERR_ALLOC = 4,  /* out of memory (host or device)                                 */
    NAQS_ERR_STATE = 5,  /* call order violated (e.g. E_loc before a lookup table was set) */
    NAQS_ERR_INDEX = 6   /* a state index outside [0, 2^n_qubits) (-> IndexError in the Python shims) */
};

/* psi element types of naqs_lookup_build / naqs_eloc */
enum { NAQS_C128 = 0, NAQS_C64 = 1 };

/* lookup-table organisations (naqs_lookup_build `kind`; NAQS_LOOKUP_AUTO picks by n_qubits) */
enum { NAQS_LOOKUP_AUTO = 0, NAQS_LOOKUP_DENSE = 1, NAQS_LOOKUP_HASH = 2 };
/* OR-ed into `kind`: the caller guarantees that the keys are unique — the contract of the reference's own call
 * update_H(states_idx, check_unseen=True, assume_unique=True) (src/optimizer/energy.py:245).  Lets the dense lookup keep
 * complex64 amplitudes as 8-byte entries (no duplicate summation needed), halving the lines a table read touches. */
#define NAQS_LOOKUP_ASSUME_UNIQUE 0x100
/* OR-ed into `kind`: keys may repeat, but every copy of a key carries the SAME amplitude (psi is a function of the
 * state; e.g. the all-gathered shards of several ranks that sampled the same configuration).  Copies are then dropped
 * instead of summed — one amplitude per key — which also permits the 8-byte-entry dense table. */
#define NAQS_LOOKUP_DUPLICATES_EQUAL 0x200

typedef struct naqs_table naqs_table_t; /* opaque, device resident */

const char* naqs_last_error(void);
int naqs_abi_version(void);
/* number of CUDA devices visible (0 without a driver); never fails */
int naqs_device_count(void);
/* kernels launched by this library in the calling process since load (bench.py's gpu_launches) */
int64_t naqs_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Term table  — replaces _PauliHamiltonianDynamic.__init__ / __calc_coupling_info's outputs
 * (src/optimizer/hamiltonian.py:241-252, 373-430): xy[k], yz[k], coeff[k] in REFERENCE TERM ORDER.
 * The table groups terms by XY mask (np.unique order, hamiltonian.py:248) keeping ascending k
 * inside a group, so that every H_ij is accumulated in exactly the order of
 * src_cpp/hamiltonian_math.pyx:31-34.
 * n_alpha / n_beta < 0  => no sector filter (the _HilbertFull case, src/utils/hilbert.py:377-378);
 * otherwise coupled states must carry n_alpha bits on even and n_beta bits on odd qubits
 * (src/utils/hilbert.py:446-449; filter of hamiltonian.py:321-328).
 * Host pointers; synchronous. */
int naqs_table_create(naqs_table_t** out, const uint64_t* h_xy, const uint64_t* h_yz, const double* h_coeff,
                      int64_t n_terms, int words, int n_qubits, int n_alpha, int n_beta, int device);
int naqs_table_destroy(naqs_table_t* t);
/* info[0..7] = K, Kxy (unique XY masks), Kyz (unique YZ masks), words, n_qubits, n_alpha, n_beta, device */
int naqs_table_info(const naqs_table_t* t, int64_t* info8);

/* ------------------------------------------------------------------------------------------
 * Amplitude lookup table — replaces the "is s' among the sampled states, and what is psi(s')" step
 * that the reference performs with scipy fancy indexing H[idx[:,None], idx]
 * (src/optimizer/hamiltonian.py:93-111 get_H; or the merge-join of src_cpp/sparse_math.pyx:316-342).
 * Builds, on `stream`, an internal device structure from T (key, psi) pairs.  Duplicate keys are
 * summed, as scipy's repeated column would be.  The table stays valid until the next build.
 * With a sector (n_alpha / n_beta >= 0) only keys inside it are stored: a coupled state outside the sector then never
 * matches, which is the reference's sector filter on coupled states (hamiltonian.py:321-328) applied once per table key;
 * a stray out-of-sector key in the batch is ignored exactly as the reference ignores it.  Hash lookups also get a
 * Bloom filter over the stored keys (shared memory up to 2^17 keys, L2-resident above). */
int naqs_lookup_build(naqs_table_t* t, const uint64_t* d_keys, const void* d_psi, int psi_dtype,
                      int64_t n_keys, int kind, void* stream);

/* Dense complex64 table owned by the CALLER (multi-GPU: the direct-address table itself is all-reduced over NVLink
 * instead of all-gathering the (key, psi) pairs — constant volume 8 * 2^n bytes, independent of the number of ranks):
 *   1. fill the table with -0.0f (bit pattern 0x80000000 = INT32_MIN; it acts as "absent" AND as a numeric zero),
 *   2. naqs_dense32_scatter: table[key] = psi for the rank's own pairs (plain stores),
 *   3. all-reduce MAX over the table viewed as int32 — any present bit pattern beats INT32_MIN, and copies of a key are
 *      equal by contract (psi is a function of the state),
 *   4. naqs_lookup_attach_dense32: naqs_eloc / naqs_apply_h then read this table (key-order walk).
 * The table must stay alive until the next naqs_lookup_build / attach, and its address must be a multiple of its size
 * (8 * 2^n bytes): the kernel forms entry addresses with XORs, (base ^ key * 8) ^ (flip * 8). */
int naqs_dense32_scatter(float* d_table, int64_t entries, const uint64_t* d_keys, const void* d_psi, int64_t n, void* stream);
int naqs_lookup_attach_dense32(naqs_table_t* t, const float* d_table, int64_t n_entries);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU exchange (the reference is single-process and has nothing to replace here): one process per GPU,
 * rows sharded, Pauli table replicated.  A communicator wraps an NCCL communicator — created from a unique id the host
 * distributes (naqs_comm_unique_id on rank 0 -> broadcast -> naqs_comm_init on every rank), or adopted from the caller
 * (naqs_comm_from_nccl) — and, for key spaces of <= 2^22 states, one CUDA-IPC-mapped region per rank through which the
 * exchanges run as push kernels over NVLink (csrc/comm.cu). */
typedef struct naqs_comm naqs_comm_t;
#define NAQS_COMM_ID_BYTES 128
#define NAQS_EXCHANGE_GATHER 0x1000 /* force the NCCL all-gather + lookup-build path */
#define NAQS_EXCHANGE_PUSH 0x2000   /* direct-address table: force the push kernel */
#define NAQS_EXCHANGE_REDUCE 0x4000 /* direct-address table: force the NCCL all-reduce (MAX) of the table */
#define NAQS_EXCHANGE_MERGE 0x8000  /* direct-address table: force the two-shot merge over peer memory (the default for dense shards) */
int naqs_comm_unique_id(void* id128);
int naqs_comm_init(naqs_comm_t** out, const void* id128, int world_size, int rank, int device);
int naqs_comm_from_nccl(naqs_comm_t** out, void* nccl_comm, int world_size, int rank, int device);
int naqs_comm_destroy(naqs_comm_t* c);
int naqs_comm_info(const naqs_comm_t* c, int* world_size, int* rank);
/* Make the (key, psi) pairs of ALL ranks the lookup table of `t` (collective; replaces naqs_lookup_build in a sharded step).
 *   n_qubits <= 22, complex64 psi: every rank stores its pairs straight into the direct-address complex64 table of every
 *     rank (peer memory), raises a flag there and waits for its peers' flags — one kernel, no reduction (copies of a key on
 *     several ranks carry the same amplitude by contract), no host synchronisation; the table is then attached like
 *     naqs_lookup_attach_dense32.  The first call maps the peer regions (synchronous).  When the shards are DENSE
 *     (max_local * (world - 1) > 2^n_qubits / 2: a push would deliver more than the table holds) the tables are MERGED
 *     instead, as a two-shot all-reduce written over peer memory: every rank scatters its pairs into its own zeroed table,
 *     rank r ORs slice r of every peer's table into its own (P2P loads) and stores the merged slice into every table (P2P
 *     stores); two flag round trips, volume 2 (world - 1) / world tables per rank whatever the rank count.
 *     NAQS_EXCHANGE_REDUCE selects round 1's form (fill with -0.0f, scatter, ncclAllReduce MAX on the int32 bit patterns).
 *   otherwise (hash lookups): every rank stores its shard, padded to the slot capacity with out-of-range keys, into its slot of
 *     every rank's gather buffer (push kernel over peer memory, one flag round trip), then one naqs_lookup_build with
 *     NAQS_LOOKUP_DUPLICATES_EQUAL over all slots; `flags` may carry NAQS_LOOKUP_DENSE / _HASH.  NAQS_EXCHANGE_GATHER selects
 *     round 1's form (NCCL all-gather of the padded shards).
 * max_local must be the SAME on every rank (the largest n_local): the choice between merge and push and the slot size of the
 * gather buffers are derived from it, and ranks that decided differently would wait for each other forever. */
int naqs_table_exchange(naqs_table_t* t, naqs_comm_t* c, const uint64_t* d_keys, const void* d_psi, int psi_dtype, int64_t n_local,
                        int64_t max_local, int flags, void* stream);
/* In-place all-reduce (sum) of the five statistics sums of naqs_eloc_stats (energy.py:328,372-375).  With a mapped peer
 * region: one push kernel (each rank writes its sums into a slot of every peer, then adds the slots in rank order — the
 * result is bitwise identical on every rank); otherwise ncclAllReduce. */
int naqs_stats_allreduce(naqs_comm_t* c, double* d_sums5, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused local energy — replaces OptimizerBase.calculate_local_energy's body
 * (src/optimizer/energy.py:245-248): update_H + get_H + sparse_dense_mv + "/ psi" + conj.
 *   E_loc[m] = conj( sum_{u} H[s_m, s_m^u] * psi_table(s_m^u) / psi[m] ),
 *   H[s, s^u] = sum_{k: xy_k = u} coeff_k * (-1)^popcount(s & yz_k)        (k ascending)
 * over the unique XY masks u whose coupled state passes the sector filter and whose H is not
 * exactly 0.0 (hamiltonian.py:328,363); coupled states absent from the lookup table contribute 0
 * (energy.py:247, set_unsampled_states_to_zero=True).  complex128 arithmetic
 * (src_cpp/sparse_math.pyx:33-37).  d_eloc: M interleaved complex128.
 * Requires a prior naqs_lookup_build on the same table. */
int naqs_eloc(naqs_table_t* t, const uint64_t* d_states, const void* d_psi, int psi_dtype, int64_t n_states,
              double* d_eloc, void* stream);

/* Matrix-free H.v — the same kernels without the division: out[m] = sum_u H[s_m, s_m^u] * v(s_m^u), complex128, where
 * the vector v is what the last naqs_lookup_build stored (keys = basis states, "psi" = v).  This is the mat-vec that
 * OptimizerBase.solve_H / calculate_energy obtain from the cached scipy CSR (src/optimizer/energy.py:199-217,762-786)
 * — here no matrix is ever formed. */
int naqs_apply_h(naqs_table_t* t, const uint64_t* d_states, int64_t n_states, double* d_out, void* stream);

/* Kernel formulation used by naqs_eloc: 0 = nibble-sliced parity + group LUT (default), 1 = direct
 * AND/POPC walk (the formulation of hamiltonian_math.pyx:449-451, 31-34 transcribed; kept for A/B checks).
 * Both give bit-identical H_ij.  The environment variable NAQS_ELOC_ALGO=direct sets the default. */
int naqs_table_set_algo(naqs_table_t* t, int algo);

/* Accumulation type of H_ij: 64 (default) = float64, the type the reference's experiments run with
 * (experiments/_base.py:234); 32 = float32, the constructor default of PauliHamiltonian.get (src/optimizer/hamiltonian.py:48):
 * the coefficients must already be float32 values (couplings.astype(np.float32), hamiltonian.py:424) and every partial
 * sum is rounded to float32 exactly as in __inner_int64_float (src_cpp/hamiltonian_math.pyx:60-100 family), so H_ij is
 * bit-identical to the reference's float32 matrix.  E_loc itself is still accumulated in complex128 (the reference uses
 * complex64 there, sparse_math.pyx:13-41) — within float32 rounding of it.  A float32 table runs the direct formulation.
 * bits other than 32 / 64 (np.float128) -> NAQS_ERR_DTYPE. */
int naqs_table_set_precision(naqs_table_t* t, int bits);

/* Key range errors.  A state index with bits at or above n_qubits is not an index at all: the reference fails with IndexError
 * when it reaches hilbert.py:607-640 (full2restricted_idx indexes a 2^N lookup table) or scipy's H[idx[:,None], idx]
 * (hamiltonian.py:94).  The device path never uses such a key as an address: it is left out of the lookup table, its own row is
 * NaN, and a flag is raised that this call reports (NAQS_ERR_INDEX) and clears.  Synchronises `stream`.  naqs_eloc_host
 * performs the check itself. */
int naqs_table_check(naqs_table_t* t, void* stream);

/* Same through HOST buffers (what a caller holding numpy arrays / CPU tensors uses; bench.py's e2e leg): uploads
 * states + psi, builds the lookup table, runs naqs_eloc and downloads E_loc.  Synchronous.
 *   key_itemsize : 8 = `words` uint64 per key; 4 / 2 = the reference's int32 / int16 state indices
 *                  (src/utils/hilbert.py:405-410), single-word keys only
 *   h_table_keys : NULL => the batch is its own lookup table (the reference's only mode), else n_table (key, psi) pairs
 *   lookup_kind  : NAQS_LOOKUP_* [| NAQS_LOOKUP_ASSUME_UNIQUE]
 *   eloc_dtype   : NAQS_C128, or NAQS_C64 = the float32 pairs the reference returns to torch (src/utils/complex.py:139-140);
 *                  the arithmetic is complex128 either way.
 * Page-locked host buffers make the copies run at full PCIe rate; pageable ones work too. */
int naqs_eloc_host(naqs_table_t* t, const void* h_states, int key_itemsize, const void* h_psi, int psi_dtype, int64_t n_states,
                   const void* h_table_keys, const void* h_table_psi, int64_t n_table, int lookup_kind,
                   void* h_eloc, int eloc_dtype);
/* The same call split in two, for callers that keep several batches in flight (one table handle per batch): _begin enqueues
 * upload, lookup build, kernel and download on the table's own stream and returns; _end waits for them and reports a key
 * range error.  With page-locked buffers the copies of one handle overlap the kernel of another (bench.py's pipelined e2e leg).
 * The host buffers must stay untouched between the two calls. */
int naqs_eloc_host_begin(naqs_table_t* t, const void* h_states, int key_itemsize, const void* h_psi, int psi_dtype, int64_t n_states,
                         const void* h_table_keys, const void* h_table_psi, int64_t n_table, int lookup_kind, void* h_eloc, int eloc_dtype);
int naqs_eloc_host_end(naqs_table_t* t);

/* ------------------------------------------------------------------------------------------
 * Stored Hamiltonian rows (CSR / coupled-set mode) — replaces update_H's row construction
 * (src/optimizer/hamiltonian.py:301-363) and get_coupled_state_idxs (:122-132).
 * Two passes: count -> (caller or naqs_exclusive_scan) -> fill.  A stored entry is a coupled state
 * that passes the sector filter with H != 0.0.  Columns of a row come in ascending unique-XY order.
 * d_col_ridx (optional, may be NULL) receives the restricted index of each column
 * (src/utils/hilbert.py:607-640 full2restricted_idx; = low key word without a sector).
 * Small batches are walked with the term table cut into up to 16 chunks on XY-group boundaries (grid.y), so that a VMC
 * sector of 10^4 states fills the GPU; every group is still summed by one thread in reference term order (bit-exact H).
 * d_indptr of naqs_rows_fill must be the exclusive scan of naqs_rows_count's output for the SAME d_states contents; a fill
 * that directly follows that count on this table (same d_states, n_states, stream) reuses its per-chunk counts. */
int naqs_rows_count(naqs_table_t* t, const uint64_t* d_states, int64_t n_states, int64_t* d_counts, void* stream);
int naqs_exclusive_scan(naqs_table_t* t, const int64_t* d_counts, int64_t n, int64_t* d_indptr /* n+1 */, void* stream);
int naqs_rows_fill(naqs_table_t* t, const uint64_t* d_states, int64_t n_states, const int64_t* d_indptr,
                   uint64_t* d_col_keys, int64_t* d_col_ridx, double* d_vals, void* stream);
/* dense H_ij[M*Kxy] exactly as get_Hij_cy returns it (src_cpp/hamiltonian_math.pyx:200-288),
 * i.e. before the sector mask and without dropping zeros */
int naqs_hij_dense(naqs_table_t* t, const uint64_t* d_states, int64_t n_states, double* d_hij, void* stream);
/* sorted unique union of keys (np.unique of hamiltonian.py:131): LSD radix sort + adjacent-unique.
 * d_out holds up to n keys; *h_n_unique is written after an internal stream sync. */
int naqs_unique_keys(naqs_table_t* t, const uint64_t* d_keys, int64_t n, uint64_t* d_out, int64_t* h_n_unique, void* stream);

/* ------------------------------------------------------------------------------------------
 * Level-0 shims: function-for-function twins of the Cython entry points, on device buffers. */
/* popcount_parity (src_cpp/hamiltonian_math.pyx:455-484): out[i] = 1 - 2*(popcount(in[i]) & 1);
 * itemsize in {1,2,4,8} (signed or unsigned: the bit pattern decides), else NAQS_ERR_DTYPE. */
int naqs_popcount_parity(const void* d_in, int itemsize, int64_t n, int8_t* d_out, void* stream);
/* get_Hij_cy (src_cpp/hamiltonian_math.pyx:200-288): H_ij[m*Kxy + u2a_xy[k]] += P[m, u2a_yz[k]] * c[k],
 * k ascending.  coeff_itemsize 8 (float64) or 4 (float32), output of the same type, else NAQS_ERR_DTYPE. */
int naqs_get_hij(int64_t n_states, int64_t n_xy, int64_t n_yz, int64_t n_terms, const int64_t* d_u2a_xy,
                 const int8_t* d_parity, const int64_t* d_u2a_yz, const void* d_coeff, int coeff_itemsize,
                 void* d_hij, void* stream);
/* sparse_dense_mv (src_cpp/sparse_math.pyx:49-243): out[r] = sum_e data[e] * v[indices[e]] over CSR row r,
 * accumulated in storage order.  data float64/float32 (data_itemsize 8/4), v and out complex of twice that
 * width, indices/indptr int32 or int64 (idx_itemsize 4/8). */
int naqs_sparse_dense_mv(const void* d_data, int data_itemsize, const void* d_indices, const void* d_indptr,
                         int idx_itemsize, int64_t n_rows, const void* d_v, void* d_out, void* stream);
/* sparse_sparse_mv (src_cpp/sparse_math.pyx:251-402): out[k] = sum over row v_idxs[k] of data * v[pos(col)]
 * for the columns present in the SORTED v_idxs list (binary search replaces the merge-join). */
int naqs_sparse_sparse_mv(const void* d_data, int data_itemsize, const void* d_indices, const void* d_indptr,
                          int idx_itemsize, const void* d_v, const void* d_v_idxs_sorted, int64_t n_v,
                          void* d_out, void* stream);
/* make_basis_idxs_cy (src_cpp/hilbert_math.pyx:12-44): out[i*N + j] = i & (1 << j), int32 [2^N, N] */
int naqs_make_basis_idxs(int n_qubits, int32_t* d_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Hilbert-space encodings (src/utils/hilbert.py) */
/* state2idx (hilbert.py:573-581): int8 rows [M, N] (occupied = value > 0) -> keys [M, words] */
int naqs_state2idx(const int8_t* d_states, int64_t n_states, int n_qubits, int words, uint64_t* d_keys, void* stream);
/* full2restricted_idx (hilbert.py:607-640) without the 2^N LUT: combinatorial rank in the
 * reference's sector order (hilbert.py:446-469), -1 outside the sector */
int naqs_restricted_index(naqs_table_t* t, const uint64_t* d_keys, int64_t n, int64_t* d_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * E_loc statistics (src/optimizer/energy.py:328,372-375), fp64:
 * out5 = [sum w, sum w*Re E, sum w*Im E, sum w*(Re E)^2, n].  d_w may be NULL (w = 1).
 * These five numbers are what the multi-GPU path all-reduces. */
int naqs_eloc_stats(naqs_table_t* t, const double* d_eloc, const double* d_w, int64_t n, double* d_out5, void* stream);
/* Loss terms of one VMC step (src/optimizer/energy.py:316-329, 367-375) from E_loc and the five sums above — the sums of
 * THIS rank, or the all-reduced ones when the batch is sharded — in fp64, results as float32 pairs like the tensors the
 * reference's loss sees (src/utils/complex.py:139-140):
 *   d_eloc_f32      [n][2]  E_loc truncated to float32 (what calculate_local_energy returns)             or NULL
 *   d_eloc_corr_f32 [n][2]  E_loc - (sum w E)/(sum w)           (e_loc_corr, energy.py:328)                or NULL
 *   d_grad_w_f32    [n][2]  2 w_i/(sum w) * conj(e_loc_corr_i): exp_op = sum_i <log_psi_i, grad_w_i> — the detached
 *                            weight vector autograd needs (energy.py:329)                                  or NULL
 *   d_energy_var    [3]     Re mean, Im mean, variance of Re E   (energy.py:372-375)                       or NULL
 * d_w: the (unnormalised) sample weights, NULL = 1. */
int naqs_loss_terms(const double* d_eloc, const double* d_w, int64_t n, const double* d_sums5, float* d_eloc_f32, float* d_eloc_corr_f32,
                    float* d_grad_w_f32, double* d_energy_var, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NAQS_ELOC_H */
