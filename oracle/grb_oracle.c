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
 * never read (ANY_PAIR, matrix.rs:926-947).  Gustavson row-wise, the algorithm family
 * SuiteSparse's saxpy3 uses, one row of A per task (schedule(dynamic,1): a 512-row frontier
 * keeps every host thread busy).  Each thread owns a PERSISTENT workspace that survives across
 * calls: an ncols-bit "seen" bitmap (2 MB at 2^24 columns -- cache resident, unlike a 4-byte
 * stamp per column) and a short list of first-seen columns.  Sparse rows are emitted by
 * sorting the list and cleared through it; rows that outgrow the list are emitted by a ctz
 * scan of the bitmap range [lo, hi] (already sorted) and cleared with one memset of that range.
 * A.ncols == B.nrows.  Fused structural mask (mode 0 none, 1 C<M>, 2 C<!M>) drops entries at
 * emission, which with replace is the full semantics of matrix.rs:1386-1394; other combinations
 * go through orc_mask_assign.  flops_out (optional) receives sum_{(i,k) in A} deg_B(k).
 */
/* rows are staged in thread-local bump arenas that persist across calls: a 512-row frontier produces multi-MB rows, and a malloc /
 * free pair per row means an mmap, a page-fault storm and an munmap per row, serialised on the process's mm lock with 128 threads */
#define ORC_ARENA_CHUNK ((size_t)64 << 20)
typedef struct { char **chunk; size_t *cap; int nchunks, cur; size_t used; } orc_arena;
typedef struct { uint64_t *bits; int64_t nwords; uint32_t *list; orc_arena ar; } orc_ws;
static void arena_reset(orc_arena *a) { a->cur = 0; a->used = 0; }
static void *arena_alloc(orc_arena *a, size_t bytes) {
    bytes = (bytes + 63) & ~(size_t)63;
    while (1) {
        if (a->cur < a->nchunks && a->used + bytes <= a->cap[a->cur]) { void *p = a->chunk[a->cur] + a->used; a->used += bytes; return p; }
        if (a->cur + 1 < a->nchunks && a->used > 0) { a->cur++; a->used = 0; continue; }           /* next retained chunk */
        if (a->cur < a->nchunks && a->used == 0 && bytes > a->cap[a->cur]) {                       /* retained chunk too small: replace */
            free(a->chunk[a->cur]);
            a->cap[a->cur] = bytes;
            a->chunk[a->cur] = malloc(bytes);
            if (!a->chunk[a->cur]) abort();
            continue;
        }
        if (a->cur >= a->nchunks || a->used > 0) {                                                 /* grow the chunk table */
            const size_t cap = bytes > ORC_ARENA_CHUNK ? bytes : ORC_ARENA_CHUNK;
            a->chunk = realloc(a->chunk, sizeof(char *) * (size_t)(a->nchunks + 1));
            a->cap = realloc(a->cap, sizeof(size_t) * (size_t)(a->nchunks + 1));
            if (!a->chunk || !a->cap) abort();
            a->chunk[a->nchunks] = malloc(cap);
            if (!a->chunk[a->nchunks]) abort();
            a->cap[a->nchunks] = cap;
            a->cur = a->nchunks;
            a->nchunks++;
            a->used = 0;
        }
    }
}
typedef struct { int64_t flops, row; } rowwork_t;
static int cmp_rowwork(const void *a, const void *b) {
    const rowwork_t *x = a, *y = b;
    if (x->flops != y->flops) return x->flops > y->flops ? -1 : 1;
    return (x->row > y->row) - (x->row < y->row);
}
#define ORC_LIST_CAP 8192
static orc_ws *g_ws = NULL;
static int g_nws = 0;
static double g_busy_s = 0.0, g_wall_s = 0.0;      /* of the last orc_mxm_anypair / orc_chain */
static int g_busy_threads = 0, g_team = 1;
static double now_s(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}
static void ws_reserve(int nthreads) {              /* called outside parallel regions */
    if (nthreads <= g_nws) return;
    g_ws = realloc(g_ws, sizeof(orc_ws) * (size_t)nthreads);
    if (!g_ws) abort();
    for (int t = g_nws; t < nthreads; t++) { g_ws[t].bits = NULL; g_ws[t].nwords = 0; g_ws[t].list = NULL; memset(&g_ws[t].ar, 0, sizeof(orc_arena)); }
    g_nws = nthreads;
}
static orc_ws *ws_get(int t, int64_t ncols) {       /* thread t's workspace, grown (zeroed) on demand */
    orc_ws *w = &g_ws[t];
    int64_t nw = (ncols + 63) / 64 + 1;
    if (w->nwords < nw) {
        free(w->bits);
        w->bits = xmalloc(sizeof(uint64_t) * (size_t)nw);
        memset(w->bits, 0, sizeof(uint64_t) * (size_t)nw);
        w->nwords = nw;
    }
    if (!w->list) w->list = xmalloc(sizeof(uint32_t) * ORC_LIST_CAP);
    return w;
}
/* dst <- src with the destination's pages FIRST TOUCHED by the whole team, 2 MB pieces dealt round-robin: on a multi-socket host
 * the copy ends up interleaved over the memory controllers instead of sitting on the node of the thread that built the graph, where
 * every thread's random row fetches would queue on one socket's DRAM.  dst must be freshly allocated (untouched) memory. */
void orc_parallel_copy(void *dst, const void *src, int64_t bytes) {
    const int64_t piece = (int64_t)2 << 20, npieces = (bytes + piece - 1) / piece;
    const int team = orc_num_threads();
    (void)team;
#pragma omp parallel for schedule(static, 1) num_threads(team)
    for (int64_t c = 0; c < npieces; c++) {
        const int64_t off = c * piece, len = bytes - off < piece ? bytes - off : piece;
        memcpy((char *)dst + off, (const char *)src + off, (size_t)len);
    }
}

/* how well the last product used the host: busy thread-seconds / (team size * wall seconds) */
double orc_last_busy_fraction(void) { return (g_wall_s > 0 && g_team > 0) ? g_busy_s / (g_wall_s * g_team) : 0.0; }
int orc_last_busy_threads(void) { return g_busy_threads; }

int orc_transpose(const orc_csr *A, orc_csr *out);
/* C<M, struct> = A*B, dot formulation: C(i,j) for (i,j) in M exists iff A(i,:) and B(:,j) share a column -- what SuiteSparse itself
 * picks for a masked product (dot3).  Nothing outside the mask is ever formed: the saxpy form below would materialise whole unmasked
 * rows first (for the triangle pattern on RMAT-24 that is ~1e11 entries, hundreds of GB -- it took the test box down).  Row i's
 * columns are marked in the thread's bitmap; for every mask entry the SHORTER of A(i,:) and B'(j,:) drives: scan B'(j,:) against the
 * bitmap, or binary-search A(i,:)'s entries in the sorted B'(j,:). */
static int mxm_masked_dot(const orc_csr *A, const orc_csr *B, const orc_csr *M, orc_csr *out, int64_t *flops_out) {
    const int64_t nrows = A->nrows;
    orc_csr BT;
    orc_csr Bp = *B;
    Bp.x = NULL;
    orc_transpose(&Bp, &BT);
    unsigned char *keep = xmalloc((size_t)(M->nnz ? M->nnz : 1));
    int64_t *cnt = xmalloc(sizeof(int64_t) * (size_t)(nrows + 1));
    int64_t flops_total = 0;
    int team = orc_num_threads();
    ws_reserve(team);
    double t_wall = now_s();
#pragma omp parallel num_threads(team) reduction(+ : flops_total)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        orc_ws *w = ws_get(tid, A->ncols);
        uint64_t *bits = w->bits;
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < nrows; i++) {
            const int64_t as = A->p[i], ae = A->p[i + 1];
            int64_t c = 0;
            for (int64_t a = as; a < ae; a++) { const uint32_t k = A->j[a]; flops_total += B->p[k + 1] - B->p[k]; bits[k >> 6] |= 1ULL << (k & 63); }
            for (int64_t q = M->p[i]; q < M->p[i + 1]; q++) {
                const uint32_t j = M->j[q];
                const int64_t bs = BT.p[j], be = BT.p[j + 1];
                int hit = 0;
                if (be - bs <= (ae - as) * 4) {
                    for (int64_t t = bs; t < be && !hit; t++) { const uint32_t k = BT.j[t]; hit = (int)((bits[k >> 6] >> (k & 63)) & 1ULL); }
                } else {
                    for (int64_t a = as; a < ae && !hit; a++) {
                        const uint32_t k = A->j[a];
                        int64_t lo = bs, hi = be;
                        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (BT.j[mid] < k) lo = mid + 1; else hi = mid; }
                        hit = (lo < be && BT.j[lo] == k);
                    }
                }
                keep[q] = (unsigned char)hit;
                c += hit;
            }
            for (int64_t a = as; a < ae; a++) bits[A->j[a] >> 6] = 0;
            cnt[i + 1] = c;
        }
    }
    cnt[0] = 0;
    for (int64_t i = 0; i < nrows; i++) cnt[i + 1] += cnt[i];
    out->nrows = nrows; out->ncols = B->ncols; out->nnz = cnt[nrows];
    out->p = cnt;
    out->j = xmalloc(sizeof(uint32_t) * (size_t)(out->nnz ? out->nnz : 1));
    out->x = NULL;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(team)
    for (int64_t i = 0; i < nrows; i++) {
        int64_t o = cnt[i];
        for (int64_t q = M->p[i]; q < M->p[i + 1]; q++) if (keep[q]) out->j[o++] = M->j[q];
    }
    free(keep);
    orc_csr_free(&BT);
    g_busy_s = 0; g_wall_s = now_s() - t_wall; g_busy_threads = team; g_team = team;
    if (flops_out) *flops_out = flops_total;
    return 0;
}

int orc_mxm_anypair(const orc_csr *A, const orc_csr *B, const orc_csr *M, int mask_mode,
                    orc_csr *out, int64_t *flops_out) {
    int64_t nrows = A->nrows, ncols = B->ncols;
    if (A->ncols != B->nrows) return -6; /* GrB_DIMENSION_MISMATCH */
    if (M && mask_mode == 1) {
        if (M->nrows != nrows || M->ncols != ncols) return -6;
        return mxm_masked_dot(A, B, M, out, flops_out);
    }
    int64_t *cnt = xmalloc(sizeof(int64_t) * (size_t)(nrows + 1));
    uint32_t **rows = xmalloc(sizeof(uint32_t *) * (size_t)(nrows ? nrows : 1));
    int64_t flops_total = 0;
    int team = orc_num_threads();
    ws_reserve(team);
    double busy_total = 0.0, t_wall = now_s();
    int busy_threads = 0;
    /* longest-processing-time-first: rows are handed out in descending flops order so the heavy
     * rows do not end up as a serial tail */
    rowwork_t *order = xmalloc(sizeof(rowwork_t) * (size_t)(nrows ? nrows : 1));
#pragma omp parallel for schedule(static) num_threads(team)
    for (int64_t i = 0; i < nrows; i++) {
        int64_t f = 0;
        for (int64_t a = A->p[i]; a < A->p[i + 1]; a++) f += B->p[A->j[a] + 1] - B->p[A->j[a]];
        order[i].flops = f; order[i].row = i;
    }
    qsort(order, (size_t)nrows, sizeof(rowwork_t), cmp_rowwork);
#pragma omp parallel num_threads(team)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        orc_ws *w = ws_get(tid, ncols);
        uint64_t *bits = w->bits;
        uint32_t *list = w->list;
        arena_reset(&w->ar);
        int64_t my_flops = 0, my_rows = 0;
        double my_busy = 0.0;
#pragma omp for schedule(dynamic, 1) nowait
        for (int64_t oi = 0; oi < nrows; oi++) {
            const int64_t i = order[oi].row;
            const double t0 = now_s();
            int64_t n = 0;               /* distinct columns seen */
            uint32_t lo = UINT32_MAX, hi = 0;
            const int64_t a_end = A->p[i + 1];
            for (int64_t a = A->p[i]; a < a_end; a++) {
                /* the rows of B a frontier row touches are short and scattered: fetch the row pointers 8 entries ahead and
                 * the first line of the column list 4 ahead, so the misses overlap instead of serialising */
                if (a + 8 < a_end) __builtin_prefetch(&B->p[A->j[a + 8]], 0, 1);
                if (a + 4 < a_end) __builtin_prefetch(&B->j[B->p[A->j[a + 4]]], 0, 1);
                const uint32_t k = A->j[a];
                const int64_t s = B->p[k], e = B->p[k + 1];
                my_flops += e - s;
                for (int64_t q = s; q < e; q++) {
                    const uint32_t c = B->j[q];
                    const uint64_t bit = 1ULL << (c & 63);
                    uint64_t *wd = &bits[c >> 6];
                    if (!(*wd & bit)) {
                        *wd |= bit;
                        if (n < ORC_LIST_CAP) list[n] = c;
                        n++;
                        if (c < lo) lo = c;
                        if (c > hi) hi = c;
                    }
                }
            }
            uint32_t *row = NULL;
            int64_t m = 0;
            if (n > 0) {
                row = arena_alloc(&w->ar, sizeof(uint32_t) * (size_t)n);
                if (n <= ORC_LIST_CAP) {             /* sparse row: sort the list, clear through it */
                    memcpy(row, list, sizeof(uint32_t) * (size_t)n);
                    qsort(row, (size_t)n, sizeof(uint32_t), cmp_u32);
                    for (int64_t q = 0; q < n; q++) bits[row[q] >> 6] = 0;
                    m = n;
                } else {                             /* dense row: the bitmap range is the sorted row */
                    const int64_t w0 = lo >> 6, w1 = hi >> 6;
                    for (int64_t x = w0; x <= w1; x++) {
                        uint64_t v = bits[x];
                        while (v) { row[m++] = (uint32_t)((x << 6) + __builtin_ctzll(v)); v &= v - 1; }
                    }
                    memset(bits + w0, 0, sizeof(uint64_t) * (size_t)(w1 - w0 + 1));
                }
                if (M && mask_mode) {
                    int64_t wq = 0;
                    for (int64_t q = 0; q < m; q++) {
                        int mk = mask_has(M, i, row[q], 1);
                        if (mask_mode == 2) mk = !mk;
                        if (mk) row[wq++] = row[q];
                    }
                    m = wq;
                }
            }
            rows[i] = row; cnt[i + 1] = m;
            my_busy += now_s() - t0;
            my_rows++;
        }
#pragma omp critical
        { flops_total += my_flops; busy_total += my_busy; busy_threads += my_rows > 0; }
    }
    free(order);
    cnt[0] = 0;
    for (int64_t i = 0; i < nrows; i++) cnt[i + 1] += cnt[i];
    out->nrows = nrows; out->ncols = ncols; out->nnz = cnt[nrows];
    out->p = cnt;
    out->j = xmalloc(sizeof(uint32_t) * (size_t)out->nnz);
    out->x = NULL;
#pragma omp parallel for schedule(dynamic, 1) num_threads(team)
    for (int64_t i = 0; i < nrows; i++) {
        int64_t m = cnt[i + 1] - cnt[i];
        if (m) memcpy(out->j + cnt[i], rows[i], sizeof(uint32_t) * (size_t)m);
    }
    free(rows);                          /* the rows themselves live in the thread arenas (kept for the next call) */
    g_busy_s = busy_total; g_wall_s = now_s() - t_wall; g_busy_threads = busy_threads; g_team = team;
    if (flops_out) *flops_out = flops_total;
    return 0;
}

/* ------------------------------------------------------------------ k-hop chain, all in C
 * CondTraverse's expand_batch core (cond_traverse.rs:600-608): F(i, src_i) = 1, then `hops` times
 * F <- F*A over ANY_PAIR (Matrix::lmxm, matrix.rs:930-947), result materialised as sorted CSR.
 * Same code path as orc_mxm_anypair; exists so that bench.py can time the CPU arm without the
 * Python-side copies of multi-GB intermediates.  out may be NULL (the result is digested and freed).
 * digest[0] = nvals, [1] = sum mix(key), [2] = sum mix(key + GOLD * position): see orc_digest. */
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDULL; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ULL; x ^= x >> 33;
    return x;
}
/* Order-sensitive digest of a CSR pattern: key = row << 32 | col, position = index in CSR order.
 * A result with the same entry set in a different order, or with one entry changed, differs. */
void orc_digest(const orc_csr *A, uint64_t *digest) {
    uint64_t s1 = 0, s2 = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : s1, s2)
    for (int64_t i = 0; i < A->nrows; i++)
        for (int64_t q = A->p[i]; q < A->p[i + 1]; q++) {
            const uint64_t key = ((uint64_t)i << 32) | A->j[q];
            s1 += mix64(key);
            s2 += mix64(key + 0x9E3779B97F4A7C15ULL * (uint64_t)(q + 1));
        }
    digest[0] = (uint64_t)A->nnz; digest[1] = s1; digest[2] = s2;
}
int orc_chain(const orc_csr *A, const uint64_t *sources, int64_t nsrc, int hops, orc_csr *out, int64_t *flops_out,
              uint64_t *digest, double *busy_fraction) {
    orc_csr F;
    F.nrows = nsrc; F.ncols = A->nrows; F.nnz = nsrc; F.x = NULL;
    F.p = xmalloc(sizeof(int64_t) * (size_t)(nsrc + 1));
    F.j = xmalloc(sizeof(uint32_t) * (size_t)(nsrc ? nsrc : 1));
    for (int64_t i = 0; i < nsrc; i++) {
        if (sources[i] >= (uint64_t)A->nrows) { free(F.p); free(F.j); return -4; }
        F.p[i] = i; F.j[i] = (uint32_t)sources[i];
    }
    F.p[nsrc] = nsrc;
    int64_t flops = 0;
    double busy = 0.0, wall = 0.0;
    for (int h = 0; h < hops; h++) {
        orc_csr T;
        int64_t fl = 0;
        int rc = orc_mxm_anypair(&F, A, NULL, 0, &T, &fl);
        orc_csr_free(&F);
        if (rc) return rc;
        F = T; flops += fl;
        busy += g_busy_s; wall += g_wall_s * g_team;
    }
    if (flops_out) *flops_out = flops;
    if (busy_fraction) *busy_fraction = wall > 0 ? busy / wall : 0.0;
    if (digest) orc_digest(&F, digest);
    if (out) *out = F; else orc_csr_free(&F);
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
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) { level[i] = -1; if (parent) parent[i] = -1; }
    int64_t *front = xmalloc(sizeof(int64_t) * (size_t)n), *next = xmalloc(sizeof(int64_t) * (size_t)n);
    int64_t nf = 1, lvl = 0;
    front[0] = src; level[src] = 0; if (parent) parent[src] = src;
    /* level-synchronous, parallel over the frontier: a vertex is claimed by one compare-and-swap on its level (so it
     * enters `next` once); its parent is the MINIMUM frontier vertex with an edge to it, kept by an atomic-min loop --
     * the result does not depend on the thread schedule. */
    while (nf > 0 && (max_level < 0 || lvl < max_level)) {
        int64_t nn = 0;
#pragma omp parallel
        {
            int64_t lcap = 4096, ln = 0;
            int64_t *loc = xmalloc(sizeof(int64_t) * (size_t)lcap);
#pragma omp for schedule(dynamic, 256) nowait
            for (int64_t f = 0; f < nf; f++) {
                const int64_t u = front[f];
                for (int64_t q = A->p[u]; q < A->p[u + 1]; q++) {
                    const int64_t v = A->j[q];
                    int64_t lv = __atomic_load_n(&level[v], __ATOMIC_RELAXED);
                    if (lv < 0) {
                        int64_t expect = -1;
                        if (__atomic_compare_exchange_n(&level[v], &expect, lvl + 1, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
                            if (ln == lcap) { lcap *= 2; loc = realloc(loc, sizeof(int64_t) * (size_t)lcap); if (!loc) abort(); }
                            loc[ln++] = v;
                            lv = lvl + 1;
                        } else lv = expect;
                    }
                    if (parent && lv == lvl + 1) {
                        int64_t cur = __atomic_load_n(&parent[v], __ATOMIC_RELAXED);
                        while ((cur < 0 || u < cur) &&
                               !__atomic_compare_exchange_n(&parent[v], &cur, u, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                    }
                }
            }
            int64_t at = __atomic_fetch_add(&nn, ln, __ATOMIC_RELAXED);
            memcpy(next + at, loc, sizeof(int64_t) * (size_t)ln);
            free(loc);
        }
        int64_t *t = front; front = next; next = t; nf = nn; lvl++;
    }
    free(front); free(next);
    return 0;
}

/* ------------------------------------------------------------------ PLUS_TIMES / PLUS_SECOND, FP64
 * y = A*x over GrB_PLUS_TIMES_SEMIRING_FP64 (A valued: x_bits holds IEEE doubles) or GxB_PLUS_SECOND_FP64 (A's values are not
 * read: "second" returns x(k)) -- the mxv LAGraph's PageRank is built on (LAGr_PageRank.c: r += AT*w over plus_second; PreJIT
 * evidence build/graphblas/PreJIT/GB_jit__AxB_saxpy3__e3f4410b0b2b0b65.c; call site algo_procedures.rs:744).  Dense x and y;
 * `present` (optional, one byte per entry of x) marks which entries of x exist, ypresent (optional) receives which entries of y
 * do: y(i) exists iff some A(i,k) meets an existing x(k).  Rows are summed in ascending column order (sequential, one
 * rounding per term); the CUDA kernel sums the same terms in a fixed tree order, so the two agree to ~1e-15 relative per term --
 * SURVEY 8(d) states rel 1e-12. */
void orc_mxv_fp64(const orc_csr *A, int use_values, const double *x, const unsigned char *present, double *y,
                  unsigned char *ypresent) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < A->nrows; i++) {
        double acc = 0.0;
        int any = 0;
        for (int64_t q = A->p[i]; q < A->p[i + 1]; q++) {
            const uint32_t k = A->j[q];
            if (present && !present[k]) continue;
            double a = 1.0;
            if (use_values) memcpy(&a, &A->x[q], sizeof(double));
            acc += use_values ? a * x[k] : x[k];
            any = 1;
        }
        y[i] = acc;
        if (ypresent) ypresent[i] = (unsigned char)any;
    }
}

/* ------------------------------------------------------------------ PageRank
 * LAGr_PageRank as the reference calls it (algo_procedures.rs:744-752: damping 0.85, tol 1e-4, itermax 100), restated from
 * LAGraph v1.x src/algorithm/LAGr_PageRank.c (third-party, not vendored: graphblas.sh:73):
 *     d = max(out_degree / damping, 1 / damping);  r = 1/n;  sink = vertices without out-edges
 *     repeat (while iters < itermax and rdiff > tol):
 *         teleport = (1 - damping) / n  +  (damping / n) * sum of r over the sinks
 *         t <- r;  w = t ./ d;  r = teleport + AT * w  (plus_second);  rdiff = sum |t - r|
 * LAGraph runs this in FP32; this restatement and the CUDA path run it in FP64 (the reference reads the result through
 * extract_vector_f64, and its own tests only need sum = 1 within 1e-4: tests/flow/test_pagerank.py:94-96).
 * AT = transpose of the adjacency pattern (rows = in-neighbours).  Returns the iteration count. */
int orc_pagerank(const orc_csr *AT, const int64_t *out_degree, double damping, double tol, int itermax, double *r) {
    const int64_t n = AT->nrows;
    if (n == 0) return 0;
    double *t = xmalloc(sizeof(double) * (size_t)n), *w = xmalloc(sizeof(double) * (size_t)n), *d = xmalloc(sizeof(double) * (size_t)n);
    const double dmin = 1.0 / damping, scaled = (1.0 - damping) / (double)n, damping_over_n = damping / (double)n;
    for (int64_t i = 0; i < n; i++) {
        const double di = (double)out_degree[i] / damping;
        d[i] = di > dmin ? di : dmin;
        r[i] = 1.0 / (double)n;
    }
    double rdiff = 1.0;
    int iters = 0;
    for (; iters < itermax && rdiff > tol; iters++) {
        double sink = 0.0;
        for (int64_t i = 0; i < n; i++) if (out_degree[i] == 0) sink += r[i];
        const double teleport = scaled + damping_over_n * sink;
        for (int64_t i = 0; i < n; i++) { t[i] = r[i]; w[i] = t[i] / d[i]; }
#pragma omp parallel for schedule(dynamic, 1024)
        for (int64_t i = 0; i < n; i++) {
            double acc = 0.0;
            for (int64_t q = AT->p[i]; q < AT->p[i + 1]; q++) acc += w[AT->j[q]];
            r[i] = teleport + acc;
        }
        rdiff = 0.0;
        for (int64_t i = 0; i < n; i++) { const double e = t[i] - r[i]; rdiff += e < 0 ? -e : e; }
    }
    free(t); free(w); free(d);
    return iters;
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
