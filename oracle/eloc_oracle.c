/* CPU oracle (C, OpenMP) for the NAQS local-energy hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load this library.  It is the fast twin of
 * oracle/eloc_oracle.py (same arithmetic, same summation orders) so that parity can be
 * checked at sizes the numpy restatement would take minutes for.
 *
 * Parity status: PINNED (tests/test_oracle_golden.py compares it with fixtures produced by
 * the reference's own code, tests/golden/make_golden.py).
 *
 * Reference lines restated (relative to the reference root):
 *   term grouping        src/optimizer/hamiltonian.py:248-252   np.unique(XY, return_inverse)
 *   sign                 src_cpp/hamiltonian_math.pyx:449-451   1 - 2*(popcount(x) & 1)
 *   H_ij accumulation    src_cpp/hamiltonian_math.pyx:31-34     serial in k (ascending)
 *   sector filter        src/optimizer/hamiltonian.py:321-328 + src/utils/hilbert.py:446-469
 *   exact-zero drop      src/optimizer/hamiltonian.py:363       (scipy sparse add)
 *   E_loc                src/optimizer/energy.py:247-248 + src_cpp/sparse_math.pyx:87-100
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC oracle/eloc_oracle.c -o oracle/liboracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t w[2]; } key_t2;

static inline int key_cmp(const uint64_t* a, const uint64_t* b, int W) {
    for (int w = W - 1; w >= 0; --w) {
        if (a[w] < b[w]) return -1;
        if (a[w] > b[w]) return 1;
    }
    return 0;
}

/* ---------------------------------------------------------------- grouping */
typedef struct { key_t2 xy; int64_t k; } term_ref;
static int g_W = 1;
static int term_ref_cmp(const void* pa, const void* pb) {
    const term_ref* a = (const term_ref*)pa; const term_ref* b = (const term_ref*)pb;
    int c = key_cmp(a->xy.w, b->xy.w, g_W);
    if (c) return c;
    return (a->k > b->k) - (a->k < b->k);
}

typedef struct {
    int W, n_qubits, n_alpha, n_beta; /* n_alpha < 0: no sector filter */
    int64_t K, G;
    uint64_t* g_xy;      /* [G*W]   unique XY masks, ascending                */
    int64_t* g_start;    /* [G+1]   term ranges (group-major, k ascending)    */
    uint64_t* t_yz;      /* [K*W]   YZ masks, group-major                     */
    double* t_c;         /* [K]                                               */
    uint64_t even[2], odd[2];
    double binom[130][66];
} otable;

void* oracle_table_create(const uint64_t* xy, const uint64_t* yz, const double* c, int64_t K, int W,
                          int n_qubits, int n_alpha, int n_beta) {
    otable* t = (otable*)calloc(1, sizeof(otable));
    t->W = W; t->n_qubits = n_qubits; t->n_alpha = n_alpha; t->n_beta = n_beta; t->K = K;
    term_ref* r = (term_ref*)malloc(sizeof(term_ref) * (K > 0 ? K : 1));
    for (int64_t k = 0; k < K; ++k) {
        r[k].xy.w[0] = xy[k * W]; r[k].xy.w[1] = W > 1 ? xy[k * W + 1] : 0; r[k].k = k;
    }
    g_W = W;
    qsort(r, K, sizeof(term_ref), term_ref_cmp);
    t->g_xy = (uint64_t*)malloc(sizeof(uint64_t) * W * (K + 1));
    t->g_start = (int64_t*)malloc(sizeof(int64_t) * (K + 2));
    t->t_yz = (uint64_t*)malloc(sizeof(uint64_t) * W * (K + 1));
    t->t_c = (double*)malloc(sizeof(double) * (K + 1));
    int64_t G = 0;
    for (int64_t i = 0; i < K; ++i) {
        if (i == 0 || key_cmp(r[i].xy.w, r[i - 1].xy.w, W) != 0) {
            for (int w = 0; w < W; ++w) t->g_xy[G * W + w] = r[i].xy.w[w];
            t->g_start[G++] = i;
        }
        for (int w = 0; w < W; ++w) t->t_yz[i * W + w] = yz[r[i].k * W + w];
        t->t_c[i] = c[r[i].k];
    }
    t->g_start[G] = K; t->G = G;
    free(r);
    for (int q = 0; q < n_qubits; ++q) {
        if (q % 2 == 0) t->even[q / 64] |= 1ull << (q % 64); else t->odd[q / 64] |= 1ull << (q % 64);
    }
    for (int n = 0; n < 130; ++n) for (int k = 0; k < 66; ++k)
        t->binom[n][k] = (k == 0) ? 1.0 : (n == 0 ? 0.0 : t->binom[n - 1][k - 1] + t->binom[n - 1][k]);
    return t;
}

void oracle_table_destroy(void* h) {
    otable* t = (otable*)h; if (!t) return;
    free(t->g_xy); free(t->g_start); free(t->t_yz); free(t->t_c); free(t);
}
int64_t oracle_table_groups(void* h) { return ((otable*)h)->G; }

/* ---------------------------------------------------------------- sector */
static inline int in_sector(const otable* t, const uint64_t* s) {
    if (t->n_alpha < 0) return 1;
    int na = 0, nb = 0;
    for (int w = 0; w < t->W; ++w) {
        na += __builtin_popcountll(s[w] & t->even[w]);
        nb += __builtin_popcountll(s[w] & t->odd[w]);
    }
    return na == t->n_alpha && nb == t->n_beta;
}

/* rank in itertools.combinations(range(n), k) order of the set bits of `bits` read at
 * positions start, start+2, ... (hilbert.py:446-447) */
static double lex_rank(const otable* t, const uint64_t* s, int start, int n, int k) {
    double r = 0; int seen = 0;
    for (int p = 0; p < n && seen < k; ++p) {
        int q = start + 2 * p;
        if ((s[q / 64] >> (q % 64)) & 1) { seen++; }
        else { r += t->binom[n - 1 - p][k - 1 - seen]; }
    }
    return r;
}

/* full2restricted_idx: -1 outside the sector; identity (low word) without a sector */
double oracle_restricted_index_one(const otable* t, const uint64_t* s) {
    if (t->n_alpha < 0) return (double)s[0] + (t->W > 1 ? 18446744073709551616.0 * (double)s[1] : 0.0);
    if (!in_sector(t, s)) return -1.0;
    int n_even = (t->n_qubits + 1) / 2, n_odd = t->n_qubits / 2;
    return lex_rank(t, s, 0, n_even, t->n_alpha) * t->binom[n_odd][t->n_beta] + lex_rank(t, s, 1, n_odd, t->n_beta);
}

void oracle_restricted_index(void* h, const uint64_t* keys, int64_t n, int64_t* out) {
    otable* t = (otable*)h;
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = (int64_t)oracle_restricted_index_one(t, keys + i * t->W);
}

/* ---------------------------------------------------------------- rows */
static inline double group_element(const otable* t, const uint64_t* s, int64_t g) {
    double acc = 0.0;
    for (int64_t i = t->g_start[g]; i < t->g_start[g + 1]; ++i) {
        uint64_t f = s[0] & t->t_yz[i * t->W];
        if (t->W > 1) f ^= s[1] & t->t_yz[i * t->W + 1];
        /* parity * c_k: the product by +-1 is exact (hamiltonian_math.pyx:34) */
        acc += (__builtin_popcountll(f) & 1) ? -t->t_c[i] : t->t_c[i];
    }
    return acc;
}

/* dense H_ij [M*G] exactly as get_Hij_cy returns it (group index = rank of the XY mask) */
void oracle_hij_dense(void* h, const uint64_t* states, int64_t M, double* out) {
    otable* t = (otable*)h;
    #pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m)
        for (int64_t g = 0; g < t->G; ++g) out[m * t->G + g] = group_element(t, states + m * t->W, g);
}

/* per-state number of stored couplings (sector filter + exact-zero drop) */
void oracle_rows_count(void* h, const uint64_t* states, int64_t M, int64_t* counts) {
    otable* t = (otable*)h;
    #pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
        const uint64_t* s = states + m * t->W; int64_t n = 0; uint64_t j[2];
        for (int64_t g = 0; g < t->G; ++g) {
            for (int w = 0; w < t->W; ++w) j[w] = s[w] ^ t->g_xy[g * t->W + w];
            if (!in_sector(t, j)) continue;
            if (group_element(t, s, g) != 0.0) n++;
        }
        counts[m] = n;
    }
}

/* fill (col_keys[nnz*W], vals[nnz]) given indptr[M+1]; columns in ascending unique-XY order */
void oracle_rows_fill(void* h, const uint64_t* states, int64_t M, const int64_t* indptr,
                      uint64_t* col_keys, double* vals) {
    otable* t = (otable*)h;
    #pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
        const uint64_t* s = states + m * t->W; int64_t e = indptr[m]; uint64_t j[2];
        for (int64_t g = 0; g < t->G; ++g) {
            for (int w = 0; w < t->W; ++w) j[w] = s[w] ^ t->g_xy[g * t->W + w];
            if (!in_sector(t, j)) continue;
            double v = group_element(t, s, g);
            if (v != 0.0) { for (int w = 0; w < t->W; ++w) col_keys[e * t->W + w] = j[w]; vals[e++] = v; }
        }
    }
}

/* ---------------------------------------------------------------- E_loc */
typedef struct { key_t2 key; int64_t idx; } tab_ref;
static int tab_ref_cmp(const void* pa, const void* pb) {
    const tab_ref* a = (const tab_ref*)pa; const tab_ref* b = (const tab_ref*)pb;
    int c = key_cmp(a->key.w, b->key.w, g_W);
    if (c) return c;
    return (a->idx > b->idx) - (a->idx < b->idx);
}
typedef struct { double rank; double val; key_t2 key; } coupl;
static int coupl_cmp(const void* pa, const void* pb) {
    const coupl* a = (const coupl*)pa; const coupl* b = (const coupl*)pb;
    if (a->rank != b->rank) return (a->rank > b->rank) - (a->rank < b->rank);
    return key_cmp(a->key.w, b->key.w, 2);
}

/* E_loc[m] = conj( sum_{s' in table} H[s_m, s'] psi(s') / psi(s_m) ), complex128.
 * Coupled states of a row are summed in ascending restricted index (the canonical CSR order
 * the reference's sub-matrix keeps); duplicate table keys are all added (in input order).
 * order_mode: 0 = reference order (ascending restricted index), 1 = ascending unique-XY order. */
int oracle_eloc(void* h, const uint64_t* states, const double* psi, int64_t M,
                const uint64_t* tkeys, const double* tpsi, int64_t T, double* eloc, int order_mode) {
    otable* t = (otable*)h; int W = t->W;
    tab_ref* tab = (tab_ref*)malloc(sizeof(tab_ref) * (T > 0 ? T : 1));
    for (int64_t i = 0; i < T; ++i) {
        tab[i].key.w[0] = tkeys[i * W]; tab[i].key.w[1] = W > 1 ? tkeys[i * W + 1] : 0; tab[i].idx = i;
    }
    g_W = W;
    qsort(tab, T, sizeof(tab_ref), tab_ref_cmp);
    #pragma omp parallel
    {
        coupl* buf = (coupl*)malloc(sizeof(coupl) * (t->G > 0 ? t->G : 1));
        #pragma omp for schedule(dynamic, 64)
        for (int64_t m = 0; m < M; ++m) {
            const uint64_t* s = states + m * W; int64_t n = 0; uint64_t j[2] = {0, 0};
            for (int64_t g = 0; g < t->G; ++g) {
                for (int w = 0; w < W; ++w) j[w] = s[w] ^ t->g_xy[g * W + w];
                if (!in_sector(t, j)) continue;
                double v = group_element(t, s, g);
                if (v == 0.0) continue;
                buf[n].val = v; buf[n].key.w[0] = j[0]; buf[n].key.w[1] = j[1];
                /* no sector: restricted index == key, compare the key words themselves */
                buf[n].rank = (order_mode == 0 && t->n_alpha >= 0) ? oracle_restricted_index_one(t, j) : 0.0;
                n++;
            }
            if (order_mode == 0) qsort(buf, n, sizeof(coupl), coupl_cmp);
            double re = 0.0, im = 0.0;
            for (int64_t e = 0; e < n; ++e) {
                int64_t lo = 0, hi = T;
                while (lo < hi) { int64_t mid = (lo + hi) / 2; if (key_cmp(tab[mid].key.w, buf[e].key.w, W) < 0) lo = mid + 1; else hi = mid; }
                for (; lo < T && key_cmp(tab[lo].key.w, buf[e].key.w, W) == 0; ++lo) {
                    re += buf[e].val * tpsi[2 * tab[lo].idx]; im += buf[e].val * tpsi[2 * tab[lo].idx + 1];
                }
            }
            /* (re + i im) / psi, then conj — numpy complex division (Smith's algorithm) */
            double a = psi[2 * m], b = psi[2 * m + 1], qr, qi;
            if (__builtin_fabs(a) >= __builtin_fabs(b)) {
                if (a == 0.0 && b == 0.0) { qr = re / __builtin_fabs(a); qi = im / __builtin_fabs(a); }
                else { double rat = b / a, scl = 1.0 / (a + b * rat); qr = (re + im * rat) * scl; qi = (im - re * rat) * scl; }
            } else { double rat = a / b, scl = 1.0 / (a * rat + b); qr = (re * rat + im) * scl; qi = (im * rat - re) * scl; }
            eloc[2 * m] = qr; eloc[2 * m + 1] = -qi;
        }
        free(buf);
    }
    free(tab);
    return 0;
}
