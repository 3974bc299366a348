/*
 * grb_oracle.c -- CPU restatement of the GraphBLAS subset on FalkorDB's traversal path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (falkordb_b200/, the C-ABI
 * library libb200grb.so) links, imports or calls this file.  Only tests/, the smoke
 * check in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs use it,
 * and there only as the checker / the reported CPU baseline.
 *
 * The arithmetic of the reference's path lives in SuiteSparse:GraphBLAS v10.5.0
 * (third-party, NOT vendored under /root/reference: graphblas.sh:72,153-156) and LAGraph
 * v1.3.x (graphblas.sh:73,218-219).  This file restates the published GraphBLAS C API v2.1
 * semantics exactly as the reference's call sites use them:
 *
 *   GrB_mxm over GxB_ANY_PAIR_BOOL            graph/src/graph/graphblas/matrix.rs:935,956,1346,1366,1386
 *   masks: structural / complement / replace  matrix.rs:1383-1394 (GrB_DESC_RSC), versioned_matrix.rs:914-924
 *   GrB_Matrix_eWiseAdd_BinaryOp              matrix.rs:862 (GxB_ANY_BOOL / GrB_SECOND_UINT64, matrix.rs:269-281)
 *   GrB_Matrix_eWiseMult_Semiring             matrix.rs:749,884
 *   GrB_transpose (+ RCT0 masked copy)        matrix.rs:658,829,841
 *   GrB_Matrix_apply(ANY_BOOL accum, ONE)     matrix.rs:913
 *   GxB_Matrix_build_Scalar / build_UINT64    matrix.rs:1199,1297 (duplicates collapse, test matrix.rs:1686-1695)
 *   LAGr_BreadthFirstSearch_Extended          runtime/functions/algo_procedures.rs:1079-1148
 *
 * Parity pinning: the reference has no direct known-answer test for GrB_mxm;
 * this oracle is pinned against (i) every known answer the reference's own tests hold at
 * this boundary -- transcribed as data in tests/golden/reference_known_answers.json and checked
 * by tests/test_golden.py and tests/test_oracle.py (matrix.rs:1617-1775, versioned_matrix.rs:
 * 1278-1330, tests/flow/test_bfs.py, test_variable_length_traversals.py, README.md:85-110) --
 * (ii) scipy.sparse as an independent implementation and (iii) algebraic identities.  For raw mxm on large inputs parity is therefore
 * "pinned to scipy + identities, unpinned by reference goldens" -- see DESIGN.md.
 *
 * Layout: CSR, rowptr int64[nrows+1], col uint32[nnz] ascending within a row,
 * val uint64[nnz] or NULL (NULL = iso/pattern-only: every stored value is `true`/1).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t nrows, ncols, nnz;
    int64_t *p;
    uint32_t *j;
    uint64_t *x; /* NULL => pattern only */
} orc_csr;

void orc_free(void *ptr) { free(ptr); }
void orc_csr_free(orc_csr *c) {
    if (!c) return;
    free(c->p); free(c->j); free(c->x);
    c->p = NULL; c->j = NULL; c->x = NULL; c->nnz = 0;
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static void *xmalloc(size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p) { fprintf(stderr, "oracle: out of memory (%zu)\n", n); abort(); }
    return p;
}

/* ------------------------------------------------------------------ sort */
static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}
typedef struct { uint64_t r, c, v; int64_t ord; } tuple_t;
static int cmp_tuple(const void *a, const void *b) {
    const tuple_t *x = a, *y = b;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    return (x->ord > y->ord) - (x->ord < y->ord);
}

/* ------------------------------------------------------------------ build
 * GxB_Matrix_build_Scalar (X == NULL) / GrB_Matrix_build_UINT64 with dup = GxB_ANY_UINT64.
 * Duplicates collapse to one entry (matrix.rs:1686-1695).  ANY may keep any duplicate's
 * value; this restatement (and the CUDA path) keeps the FIRST in input order so the two
 * are comparable bit for bit.  Returns -1 on an out-of-range index (GrB_INDEX_OUT_OF_BOUNDS).
 */
int orc_build(int64_t nrows, int64_t ncols, int64_t n, const uint64_t *I, const uint64_t *J,
              const uint64_t *X, orc_csr *out) {
    tuple_t *t = xmalloc(sizeof(tuple_t) * (size_t)n);
    for (int64_t k = 0; k < n; k++) {
        if (I[k] >= (uint64_t)nrows || J[k] >= (uint64_t)ncols) { free(t); return -1; }
        t[k].r = I[k]; t[k].c = J[k]; t[k].v = X ? X[k] : 1; t[k].ord = k;
    }
    qsort(t, (size_t)n, sizeof(tuple_t), cmp_tuple);
    out->nrows = nrows; out->ncols = ncols;
    out->p = xmalloc(sizeof(int64_t) * (size_t)(nrows + 1));
    out->j = xmalloc(sizeof(uint32_t) * (size_t)n);
    out->x = X ? xmalloc(sizeof(uint64_t) * (size_t)n) : NULL;
    memset(out->p, 0, sizeof(int64_t) * (size_t)(nrows + 1));
    int64_t m = 0;
    for (int64_t k = 0; k < n; k++) {
        if (k > 0 && t[k].r == t[k - 1].r && t[k].c == t[k - 1].c) continue;
        out->j[m] = (uint32_t)t[k].c;
        if (X) out->x[m] = t[k].v;
        out->p[t[k].r + 1]++;
        m++;
    }
    for (int64_t r = 0; r < nrows; r++) out->p[r + 1] += out->p[r];
    out->nnz = m;
    free(t);
    return 0;
}

/* ------------------------------------------------------------------ mask helpers */
/* is (row i, col c) "true" in mask M?  structural: entry present.  valued: present and != 0 */
static inline int mask_has(const orc_csr *M, int64_t i, uint32_t c, int structural) {
    int64_t lo = M->p[i], hi = M->p[i + 1];
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (M->j[mid] < c) lo = mid + 1; else hi = mid;
    }
    if (lo < M->p[i + 1] && M->j[lo] == c) {
        if (structural || !M->x) return 1;
        return M->x[lo] != 0;
    }
    return 0;
}

/*
 * GraphBLAS write-back, accum == NULL (every call site on the path) or accum == ANY
 * (GrB_Matrix_apply in set_pattern, matrix.rs:913-920):
 *     Z = accum ? (C_old (+) T) : T
 *     C(i,j) = mask(i,j) ? Z(i,j) : (replace ? none : C_old(i,j))
 * `M == NULL` means no mask: with complement set that is an empty mask (nothing written;
 * C_old kept unless replace), as the spec says.  T_x/C_old x may be NULL (pattern).
 * When accum merges two entries the result value is T's (ANY may pick either; bool is iso).
 */
int orc_mask_assign(const orc_csr *Cold, const orc_csr *T, const orc_csr *M, int comp,
                    int structural, int replace, int accum, orc_csr *out) {
    int64_t nrows = T->nrows;
    int has_x = (T->x != NULL) || (Cold && Cold->x != NULL);
    int64_t cap = T->nnz + (Cold ? Cold->nnz : 0);
    out->nrows = nrows; out->ncols = T->ncols;
    out->p = xmalloc(sizeof(int64_t) * (size_t)(nrows + 1));
    out->j = xmalloc(sizeof(uint32_t) * (size_t)cap);
    out->x = has_x ? xmalloc(sizeof(uint64_t) * (size_t)cap) : NULL;
    int64_t m = 0;
    out->p[0] = 0;
    for (int64_t i = 0; i < nrows; i++) {
        int64_t a = T->p[i], ae = T->p[i + 1];
        int64_t b = Cold ? Cold->p[i] : 0, be = Cold ? Cold->p[i + 1] : 0;
        while (a < ae || b < be) {
            uint32_t c; int inT = 0, inC = 0;
            if (b >= be || (a < ae && T->j[a] <= Cold->j[b])) { c = T->j[a]; inT = 1; }
            else c = Cold->j[b];
            if (b < be && Cold->j[b] == c) inC = 1;
            int mk = M ? mask_has(M, i, c, structural) : 1;
            if (comp) mk = !mk;
            int keep; uint64_t v = 1;
            if (mk) {
                if (inT) { keep = 1; v = T->x ? T->x[a] : 1; }
                else if (accum && inC) { keep = 1; v = Cold->x ? Cold->x[b] : 1; }
                else keep = 0;
            } else {
                keep = (!replace) && inC;
                if (keep) v = Cold->x ? Cold->x[b] : 1;
            }
            if (keep) { out->j[m] = c; if (has_x) out->x[m] = v; m++; }
            if (inT) a++;
            if (inC) b++;
        }
        out->p[i + 1] = m;
    }
    out->nnz = m;
    return 0;
}

/* ------------------------------------------------------------------ mxm, ANY_PAIR
 * T = pattern(A*B): T(i,j) present iff exists k with A(i,k) and B(k,j) present; values are
 * never read (ANY_PAIR, matrix.rs:926-947).  Gustavson row-wise with a per-thread stamp
 * workspace, the algorithm family SuiteSparse's saxpy3 uses.  A.ncols == B.nrows.
 * Fused structural mask (mode 0 none, 1 C<M>, 2 C<!M>) drops entries at emission, which with
 * replace is the full semantics of matrix.rs:1386-1394; other combinations go through
 * orc_mask_assign.  flops_out (optional) receives sum_{(i,k) in A} deg_B(k).
 */
int orc_mxm_anypair(const orc_csr *A, const orc_csr *B, const orc_csr *M, int mask_mode,
                    orc_csr *out, int64_t *flops_out) {
    int64_t nrows = A->nrows, ncols = B->ncols;
    if (A->ncols != B->nrows) return -6; /* GrB_DIMENSION_MISMATCH */
    int64_t *cnt = xmalloc(sizeof(int64_t) * (size_t)(nrows + 1));
    uint32_t **rows = xmalloc(sizeof(uint32_t *) * (size_t)(nrows ? nrows : 1));
    int64_t flops_total = 0;
#pragma omp parallel
    {
        uint32_t *stamp = xmalloc(sizeof(uint32_t) * (size_t)(ncols ? ncols : 1));
        memset(stamp, 0, sizeof(uint32_t) * (size_t)ncols);
        uint32_t cur = 0;
        int64_t lcap = 1024;
        uint32_t *list = xmalloc(sizeof(uint32_t) * (size_t)lcap);
        int64_t my_flops = 0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < nrows; i++) {
            if (++cur == 0) { memset(stamp, 0, sizeof(uint32_t) * (size_t)ncols); cur = 1; }
            int64_t n = 0;
            uint32_t lo = UINT32_MAX, hi = 0;
            for (int64_t a = A->p[i]; a < A->p[i + 1]; a++) {
                uint32_t k = A->j[a];
                int64_t s = B->p[k], e = B->p[k + 1];
                my_flops += e - s;
                if (n + (e - s) > lcap) {
                    while (n + (e - s) > lcap) lcap *= 2;
                    list = realloc(list, sizeof(uint32_t) * (size_t)lcap);
                    if (!list) abort();
                }
                for (int64_t q = s; q < e; q++) {
                    uint32_t c = B->j[q];
                    if (stamp[c] != cur) {
                        stamp[c] = cur; list[n++] = c;
                        if (c < lo) lo = c;
                        if (c > hi) hi = c;
                    }
                }
            }
            /* sorted emission: scan the stamp range when dense enough, else sort the list */
            uint32_t *row = NULL;
            int64_t m = 0;
            if (n > 0) {
                row = xmalloc(sizeof(uint32_t) * (size_t)n);
                if ((int64_t)(hi - lo) < 16 * n) {
                    for (uint32_t c = lo;; c++) { if (stamp[c] == cur) row[m++] = c; if (c == hi) break; }
                } else {
                    memcpy(row, list, sizeof(uint32_t) * (size_t)n);
                    qsort(row, (size_t)n, sizeof(uint32_t), cmp_u32);
                    m = n;
                }
                if (M && mask_mode) {
                    int64_t w = 0;
                    for (int64_t q = 0; q < m; q++) {
                        int mk = mask_has(M, i, row[q], 1);
                        if (mask_mode == 2) mk = !mk;
                        if (mk) row[w++] = row[q];
                    }
                    m = w;
                }
            }
            rows[i] = row; cnt[i + 1] = m;
        }
#pragma omp atomic
        flops_total += my_flops;
        free(stamp); free(list);
    }
    cnt[0] = 0;
    for (int64_t i = 0; i < nrows; i++) cnt[i + 1] += cnt[i];
    out->nrows = nrows; out->ncols = ncols; out->nnz = cnt[nrows];
    out->p = cnt;
    out->j = xmalloc(sizeof(uint32_t) * (size_t)out->nnz);
    out->x = NULL;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < nrows; i++) {
        int64_t m = cnt[i + 1] - cnt[i];
        if (m) memcpy(out->j + cnt[i], rows[i], sizeof(uint32_t) * (size_t)m);
        free(rows[i]);
    }
    free(rows);
    if (flops_out) *flops_out = flops_total;
    return 0;
}

/* ------------------------------------------------------------------ eWiseAdd
 * T = A (+) B, set union.  On overlap op(a,b): GxB_ANY_BOOL -> true; GrB_SECOND_UINT64 -> b
 * (matrix.rs:257-281).  Single-side entries are copied.  keep_values: emit a value array
 * (u64 matrices); a NULL input x reads as 1.
 */
int orc_ewise_add(const orc_csr *A, const orc_csr *B, int keep_values, orc_csr *out) {
    if (A->nrows != B->nrows || A->ncols != B->ncols) return -6;
    int64_t nrows = A->nrows, cap = A->nnz + B->nnz;
    out->nrows = nrows; out->ncols = A->ncols;
    out->p = xmalloc(sizeof(int64_t) * (size_t)(nrows + 1));
    out->j = xmalloc(sizeof(uint32_t) * (size_t)cap);
    out->x = keep_values ? xmalloc(sizeof(uint64_t) * (size_t)cap) : NULL;
    int64_t m = 0; out->p[0] = 0;
    for (int64_t i = 0; i < nrows; i++) {
        int64_t a = A->p[i], ae = A->p[i + 1], b = B->p[i], be = B->p[i + 1];
        while (a < ae || b < be) {
            if (b >= be || (a < ae && A->j[a] < B->j[b])) {
                out->j[m] = A->j[a]; if (keep_values) out->x[m] = A->x ? A->x[a] : 1; a++;
            } else if (a >= ae || B->j[b] < A->j[a]) {
                out->j[m] = B->j[b]; if (keep_values) out->x[m] = B->x ? B->x[b] : 1; b++;
            } else {
                out->j[m] = B->j[b]; if (keep_values) out->x[m] = B->x ? B->x[b] : 1; a++; b++;
            }
            m++;
        }
        out->p[i + 1] = m;
    }
    out->nnz = m;
    return 0;
}

/* ------------------------------------------------------------------ eWiseMult, ANY_PAIR
 * T = pattern(A) intersect pattern(B), all true (matrix.rs:743-761, 876-896). */
int orc_ewise_mult(const orc_csr *A, const orc_csr *B, orc_csr *out) {
    if (A->nrows != B->nrows || A->ncols != B->ncols) return -6;
    int64_t nrows = A->nrows, cap = A->nnz < B->nnz ? A->nnz : B->nnz;
    out->nrows = nrows; out->ncols = A->ncols;
    out->p = xmalloc(sizeof(int64_t) * (size_t)(nrows + 1));
    out->j = xmalloc(sizeof(uint32_t) * (size_t)cap);
    out->x = NULL;
    int64_t m = 0; out->p[0] = 0;
    for (int64_t i = 0; i < nrows; i++) {
        int64_t a = A->p[i], ae = A->p[i + 1], b = B->p[i], be = B->p[i + 1];
        while (a < ae && b < be) {
            if (A->j[a] < B->j[b]) a++;
            else if (B->j[b] < A->j[a]) b++;
            else { out->j[m++] = A->j[a]; a++; b++; }
        }
        out->p[i + 1] = m;
    }
    out->nnz = m;
    return 0;
}

/* ------------------------------------------------------------------ transpose
 * T = A', values carried (matrix.rs:633-662).  Counting sort by column; rows stay ascending. */
int orc_transpose(const orc_csr *A, orc_csr *out) {
    int64_t nr = A->ncols, nnz = A->nnz;
    out->nrows = nr; out->ncols = A->nrows; out->nnz = nnz;
    out->p = xmalloc(sizeof(int64_t) * (size_t)(nr + 1));
    out->j = xmalloc(sizeof(uint32_t) * (size_t)nnz);
    out->x = A->x ? xmalloc(sizeof(uint64_t) * (size_t)nnz) : NULL;
    memset(out->p, 0, sizeof(int64_t) * (size_t)(nr + 1));
    for (int64_t q = 0; q < nnz; q++) out->p[A->j[q] + 1]++;
    for (int64_t r = 0; r < nr; r++) out->p[r + 1] += out->p[r];
    int64_t *cur = xmalloc(sizeof(int64_t) * (size_t)(nr ? nr : 1));
    memcpy(cur, out->p, sizeof(int64_t) * (size_t)nr);
    for (int64_t i = 0; i < A->nrows; i++)
        for (int64_t q = A->p[i]; q < A->p[i + 1]; q++) {
            int64_t d = cur[A->j[q]]++;
            out->j[d] = (uint32_t)i;
            if (A->x) out->x[d] = A->x[q];
        }
    free(cur);
    return 0;
}

/* ------------------------------------------------------------------ BFS
 * LAGr_BreadthFirstSearch_Extended as the reference consumes it
 * (algo_procedures.rs:1079-1148): level(src)=0, parent(src)=src, unreached = -1 (absent in
 * the sparse INT64 vectors), max_level < 0 means unbounded, else stop after that many hops.
 * LAGraph's parent comes from ANY_SECONDI (any valid parent); the deterministic tie-break
 * chosen here and in the CUDA path is the MINIMUM parent id in the previous level.
 */
int orc_bfs(const orc_csr *A, int64_t src, int64_t max_level, int64_t *level, int64_t *parent) {
    int64_t n = A->nrows;
    if (src < 0 || src >= n) return -4;
    for (int64_t i = 0; i < n; i++) { level[i] = -1; if (parent) parent[i] = -1; }
    int64_t *front = xmalloc(sizeof(int64_t) * (size_t)n), *next = xmalloc(sizeof(int64_t) * (size_t)n);
    int64_t nf = 1, lvl = 0;
    front[0] = src; level[src] = 0; if (parent) parent[src] = src;
    while (nf > 0 && (max_level < 0 || lvl < max_level)) {
        int64_t nn = 0;
        for (int64_t f = 0; f < nf; f++) {
            int64_t u = front[f];
            for (int64_t q = A->p[u]; q < A->p[u + 1]; q++) {
                int64_t v = A->j[q];
                if (level[v] < 0) { level[v] = lvl + 1; if (parent) parent[v] = u; next[nn++] = v; }
                else if (parent && level[v] == lvl + 1 && u < parent[v]) parent[v] = u;
            }
        }
        int64_t *t = front; front = next; next = t; nf = nn; lvl++;
    }
    free(front); free(next);
    return 0;
}

/* ------------------------------------------------------------------ RMAT generator
 * Graph500-style Kronecker edges (a,b,c,d = .57,.19,.19,.05), counter-based so the CUDA
 * generator (falkordb_b200/csrc/rmat.cuh) produces the identical edge list: edge e, level l
 * draws a 32-bit uniform from splitmix64(seed, e, l/2); a bijective scramble of the vertex
 * ids follows.  Self loops / duplicates are removed by the build step, as SURVEY 8(d) says.
 */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static inline uint64_t scramble(uint64_t v, int scale, uint64_t seed) {
    uint64_t mask = (scale >= 64) ? ~0ULL : ((1ULL << scale) - 1);
    uint64_t k1 = (splitmix64(seed ^ 0xA5A5A5A5ULL) | 1ULL), k2 = (splitmix64(seed ^ 0x5A5A5A5AULL) | 1ULL);
    int sh = scale / 2 > 0 ? scale / 2 : 1;
    v = (v * k1) & mask;
    v ^= v >> sh;
    v = (v * k2) & mask;
    v ^= v >> sh;
    return v & mask;
}
void orc_rmat_edges(int scale, int64_t nedges, uint64_t seed, uint64_t *I, uint64_t *J) {
    const uint32_t TA = (uint32_t)(0.57 * 4294967296.0), TB = (uint32_t)((0.57 + 0.19) * 4294967296.0),
                   TC = (uint32_t)((0.57 + 0.19 + 0.19) * 4294967296.0);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < nedges; e++) {
        uint64_t r = 0, c = 0, h = 0;
        for (int l = 0; l < scale; l++) {
            if ((l & 1) == 0) h = splitmix64(seed * 0xD1342543DE82EF95ULL + (uint64_t)e * 64 + (uint64_t)(l >> 1));
            uint32_t u = (l & 1) ? (uint32_t)(h >> 32) : (uint32_t)h;
            int rb, cb;
            if (u < TA) { rb = 0; cb = 0; } else if (u < TB) { rb = 0; cb = 1; }
            else if (u < TC) { rb = 1; cb = 0; } else { rb = 1; cb = 1; }
            r = (r << 1) | (uint64_t)rb; c = (c << 1) | (uint64_t)cb;
        }
        I[e] = scramble(r, scale, seed); J[e] = scramble(c, scale, seed);
    }
}

/* Build a deduplicated, self-loop-free pattern CSR from an edge list (parallel-friendly
 * counting approach for the large cpu_baseline inputs; orc_build's qsort is for tests). */
int orc_csr_from_edges(int64_t n, int64_t ne, const uint64_t *I, const uint64_t *J, orc_csr *out) {
    int64_t *p = xmalloc(sizeof(int64_t) * (size_t)(n + 1));
    memset(p, 0, sizeof(int64_t) * (size_t)(n + 1));
    for (int64_t e = 0; e < ne; e++) if (I[e] != J[e]) p[I[e] + 1]++;
    for (int64_t r = 0; r < n; r++) p[r + 1] += p[r];
    uint32_t *jj = xmalloc(sizeof(uint32_t) * (size_t)p[n]);
    int64_t *cur = xmalloc(sizeof(int64_t) * (size_t)(n ? n : 1));
    memcpy(cur, p, sizeof(int64_t) * (size_t)n);
    for (int64_t e = 0; e < ne; e++) if (I[e] != J[e]) jj[cur[I[e]]++] = (uint32_t)J[e];
    free(cur);
    int64_t *np = xmalloc(sizeof(int64_t) * (size_t)(n + 1));
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t r = 0; r < n; r++) {
        int64_t s = p[r], e = p[r + 1];
        qsort(jj + s, (size_t)(e - s), sizeof(uint32_t), cmp_u32);
        int64_t w = s;
        for (int64_t q = s; q < e; q++) if (q == s || jj[q] != jj[q - 1]) jj[w++] = jj[q];
        np[r + 1] = w - s;
    }
    np[0] = 0;
    for (int64_t r = 0; r < n; r++) np[r + 1] += np[r];
    uint32_t *oj = xmalloc(sizeof(uint32_t) * (size_t)np[n]);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t r = 0; r < n; r++) memcpy(oj + np[r], jj + p[r], sizeof(uint32_t) * (size_t)(np[r + 1] - np[r]));
    free(jj); free(p);
    out->nrows = n; out->ncols = n; out->nnz = np[n]; out->p = np; out->j = oj; out->x = NULL;
    return 0;
}
