/*
 * b200grb.h -- C ABI of libb200grb.so, the Blackwell-native GraphBLAS traversal backend.
 *
 * Drop-in boundary: FalkorDB links SuiteSparse:GraphBLAS through the bindgen
 * declarations in graph/src/graph/graphblas/mod.rs and the link line graph/build.rs:52-55.
 * Every declaration below mirrors the C signature mod.rs binds (file:line cited per entry), so
 * swapping `static=graphblas` for `dylib=b200grb` re-points the graph store's hot calls
 * (graph/src/graph/graphblas/matrix.rs) at this library.  All entry points return GrB_Info
 * (mod.rs:274-296); handles are opaque and owned by the library until *_free.
 *
 * Scope: the subset on the traversal path -- GrB_mxm over GxB_ANY_PAIR_BOOL with structural /
 * complemented / replace masks, eWiseAdd / eWiseMult / transpose (incl. the RCT0 masked copy) /
 * apply(ONE) for delta-matrix sync, build, element access, row iterators, wait, and the BFS entry
 * LAGr_BreadthFirstSearch_Extended.  Anything else returns GrB_NOT_IMPLEMENTED.  There is no CPU
 * fallback: bulk operations run on the GPU or fail with GxB_GPU_ERROR.
 *
 * B200_* entry points are extensions (zero-copy CSR hand-off, statistics, the synthetic RMAT input
 * generator) with no SuiteSparse counterpart; they play the role GxB_Container load/unload plays in
 * the reference (matrix.rs:428-546).
 */
#ifndef B200GRB_H
#define B200GRB_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t GrB_Index; /* mod.rs:271 */

typedef enum { /* mod.rs:274-296 */
    GrB_SUCCESS = 0,
    GrB_NO_VALUE = 1,
    GxB_EXHAUSTED = 7089,
    GrB_UNINITIALIZED_OBJECT = -1,
    GrB_NULL_POINTER = -2,
    GrB_INVALID_VALUE = -3,
    GrB_INVALID_INDEX = -4,
    GrB_DOMAIN_MISMATCH = -5,
    GrB_DIMENSION_MISMATCH = -6,
    GrB_OUTPUT_NOT_EMPTY = -7,
    GrB_NOT_IMPLEMENTED = -8,
    GrB_ALREADY_SET = -9,
    GrB_PANIC = -101,
    GrB_OUT_OF_MEMORY = -102,
    GrB_INSUFFICIENT_SPACE = -103,
    GrB_INVALID_OBJECT = -104,
    GrB_INDEX_OUT_OF_BOUNDS = -105,
    GrB_EMPTY_OBJECT = -106,
    GxB_JIT_ERROR = -7001,
    GxB_GPU_ERROR = -7002,
    GxB_OUTPUT_IS_READONLY = -7003
} GrB_Info;

typedef enum { GrB_NONBLOCKING = 0, GrB_BLOCKING = 1 } GrB_Mode;
typedef enum { GrB_COMPLETE = 0, GrB_MATERIALIZE = 1 } GrB_WaitMode;   /* mod.rs:3032-3033 */
typedef enum { GrB_ROWMAJOR = 0, GrB_COLMAJOR = 1 } GrB_Orientation;   /* mod.rs:3006-3007 */

/* option fields used by the reference (mod.rs:2887-2963, 152) */
#define GrB_STORAGE_ORIENTATION_HINT 100
#define GxB_BURBLE 7019
#define GxB_JIT_C_CONTROL 7029
#define GxB_SPARSITY_STATUS 7034
#define GxB_SPARSITY_CONTROL 7036
#define GxB_HYPER_HASH 7048
#define GxB_WILL_WAIT 7076
#define GxB_NTHREADS 7086
/* sparsity bits (mod.rs:159-162) */
#define GxB_HYPERSPARSE 1
#define GxB_SPARSE 2
#define GxB_BITMAP 4
#define GxB_FULL 8

typedef struct GB_Type_opaque *GrB_Type;             /* mod.rs:316 */
typedef struct GB_UnaryOp_opaque *GrB_UnaryOp;       /* mod.rs:322 */
typedef struct GB_BinaryOp_opaque *GrB_BinaryOp;     /* mod.rs:328 */
typedef struct GB_Semiring_opaque *GrB_Semiring;     /* mod.rs:346 */
typedef struct GB_Descriptor_opaque *GrB_Descriptor; /* mod.rs:310 */
typedef struct GB_Scalar_opaque *GrB_Scalar;         /* mod.rs:352 */
typedef struct GB_Vector_opaque *GrB_Vector;         /* mod.rs:358 */
typedef struct GB_Matrix_opaque *GrB_Matrix;         /* mod.rs:364 */
typedef struct GB_Global_opaque *GrB_Global;         /* mod.rs:370 */
typedef struct GB_Iterator_opaque *GxB_Iterator;     /* mod.rs:383 */

/* exported data symbols (mod.rs:430-544, 721, 1301, 1547, 1643, 3001, 6852) */
extern GrB_Type GrB_BOOL, GrB_UINT64, GrB_INT64, GrB_FP64;
extern GrB_Semiring GxB_ANY_PAIR_BOOL;
/* the weighted / ranking corner of the path (north_star: "a stated fp tolerance for PLUS_TIMES weighted paths"): FP64 mxv / vxm */
extern GrB_Semiring GrB_PLUS_TIMES_SEMIRING_FP64, GxB_PLUS_SECOND_FP64;
extern GrB_BinaryOp GxB_ANY_BOOL, GrB_SECOND_UINT64, GxB_ANY_UINT64, GrB_PLUS_FP64;
extern GrB_UnaryOp GxB_ONE_BOOL;
extern const GrB_Global GrB_GLOBAL;
/* the 31 predefined descriptors: T0/T1 = transpose input 0/1, C = complement mask,
 * S = structural mask, R = replace output (matrix.rs:80-85, 313-351) */
extern GrB_Descriptor GrB_DESC_T1, GrB_DESC_T0, GrB_DESC_T0T1, GrB_DESC_C, GrB_DESC_CT1, GrB_DESC_CT0, GrB_DESC_CT0T1,
    GrB_DESC_S, GrB_DESC_ST1, GrB_DESC_ST0, GrB_DESC_ST0T1, GrB_DESC_SC, GrB_DESC_SCT1, GrB_DESC_SCT0, GrB_DESC_SCT0T1,
    GrB_DESC_R, GrB_DESC_RT1, GrB_DESC_RT0, GrB_DESC_RT0T1, GrB_DESC_RC, GrB_DESC_RCT1, GrB_DESC_RCT0, GrB_DESC_RCT0T1,
    GrB_DESC_RS, GrB_DESC_RST1, GrB_DESC_RST0, GrB_DESC_RST0T1, GrB_DESC_RSC, GrB_DESC_RSCT1, GrB_DESC_RSCT0,
    GrB_DESC_RSCT0T1;

/* ---- lifecycle (matrix.rs:116-221) ---- */
GrB_Info GxB_init(int mode, void *(*user_malloc)(size_t), void *(*user_calloc)(size_t, size_t),
                  void *(*user_realloc)(void *, size_t), void (*user_free)(void *)); /* mod.rs:7970 */
GrB_Info GrB_init(int mode);
GrB_Info GrB_finalize(void);                                                          /* mod.rs:7967 */
GrB_Info GrB_Global_set_INT32(GrB_Global g, int32_t value, int field);                /* mod.rs:10974 */
GrB_Info GxB_Global_Option_set_INT32(int field, int32_t value);                       /* mod.rs:15577 */

/* ---- matrix objects ---- */
GrB_Info GrB_Matrix_new(GrB_Matrix *A, GrB_Type type, GrB_Index nrows, GrB_Index ncols); /* mod.rs:9444 */
GrB_Info GrB_Matrix_dup(GrB_Matrix *C, GrB_Matrix A);                                    /* mod.rs:9452 */
GrB_Info GrB_Matrix_free(GrB_Matrix *A);                                                 /* mod.rs:15072 */
GrB_Info GrB_Matrix_clear(GrB_Matrix A);                                                 /* mod.rs:9476 */
GrB_Info GrB_Matrix_resize(GrB_Matrix C, GrB_Index nrows_new, GrB_Index ncols_new);      /* mod.rs:14055 */
GrB_Info GrB_Matrix_nrows(GrB_Index *nrows, GrB_Matrix A);                               /* mod.rs:9479 */
GrB_Info GrB_Matrix_ncols(GrB_Index *ncols, GrB_Matrix A);                               /* mod.rs:9485 */
GrB_Info GrB_Matrix_nvals(GrB_Index *nvals, GrB_Matrix A);                               /* mod.rs:9491 */
GrB_Info GrB_Matrix_set_INT32(GrB_Matrix A, int32_t value, int field);                   /* mod.rs:10713 */
GrB_Info GrB_Matrix_get_INT32(GrB_Matrix A, int32_t *value, int field);                  /* mod.rs:10230 */
GrB_Info GxB_Matrix_type(GrB_Type *type, GrB_Matrix A);                                  /* mod.rs:9503 */
GrB_Info GxB_Matrix_iso(bool *iso, GrB_Matrix A);                                        /* mod.rs:15103 */
GrB_Info GxB_Matrix_memoryUsage(size_t *size, GrB_Matrix A);                             /* mod.rs:9497 */
GrB_Info GxB_Matrix_fprint(GrB_Matrix A, const char *name, int pr, FILE *f);             /* mod.rs:14132 */
GrB_Info GrB_Matrix_wait(GrB_Matrix A, int waitmode);                                    /* mod.rs:11078 */

/* ---- element access (matrix.rs:1143-1172, 1248-1275, 731-737, 1035-1045) ---- */
GrB_Info GrB_Matrix_setElement_BOOL(GrB_Matrix C, bool x, GrB_Index i, GrB_Index j);       /* mod.rs:9685 */
GrB_Info GrB_Matrix_setElement_UINT64(GrB_Matrix C, uint64_t x, GrB_Index i, GrB_Index j); /* mod.rs:9749 */
GrB_Info GrB_Matrix_extractElement_BOOL(bool *x, GrB_Matrix A, GrB_Index i, GrB_Index j);  /* mod.rs:9797 */
GrB_Info GrB_Matrix_extractElement_UINT64(uint64_t *x, GrB_Matrix A, GrB_Index i, GrB_Index j); /* mod.rs:9861 */
GrB_Info GrB_Matrix_removeElement(GrB_Matrix C, GrB_Index i, GrB_Index j);                 /* mod.rs:9924 */
GrB_Info GxB_Matrix_isStoredElement(GrB_Matrix A, GrB_Index i, GrB_Index j);               /* mod.rs:9917 */
GrB_Info GrB_Matrix_extractTuples_BOOL(GrB_Index *I, GrB_Index *J, bool *X, GrB_Index *nvals, GrB_Matrix A);     /* mod.rs:9931 */
GrB_Info GrB_Matrix_extractTuples_UINT64(GrB_Index *I, GrB_Index *J, uint64_t *X, GrB_Index *nvals, GrB_Matrix A); /* mod.rs:10003 */

/* ---- build (matrix.rs:1186-1210, 1281-1303) ---- */
GrB_Info GrB_Scalar_new(GrB_Scalar *s, GrB_Type type);      /* mod.rs:8677 */
GrB_Info GrB_Scalar_setElement_BOOL(GrB_Scalar s, bool x);  /* mod.rs:8726 */
GrB_Info GrB_Scalar_free(GrB_Scalar *s);                    /* mod.rs:15066 */
GrB_Info GxB_Matrix_build_Scalar(GrB_Matrix C, const GrB_Index *I, const GrB_Index *J, GrB_Scalar scalar,
                                 GrB_Index nvals);          /* mod.rs:9659 */
GrB_Info GrB_Matrix_build_UINT64(GrB_Matrix C, const GrB_Index *I, const GrB_Index *J, const uint64_t *X,
                                 GrB_Index nvals, GrB_BinaryOp dup); /* mod.rs:9589 */
GrB_Info GrB_Matrix_build_BOOL(GrB_Matrix C, const GrB_Index *I, const GrB_Index *J, const bool *X, GrB_Index nvals,
                               GrB_BinaryOp dup);           /* mod.rs:9509 */

/* ---- bulk algebra: the hot calls ---- */
GrB_Info GrB_mxm(GrB_Matrix C, GrB_Matrix Mask, GrB_BinaryOp accum, GrB_Semiring semiring, GrB_Matrix A, GrB_Matrix B,
                 GrB_Descriptor desc);                      /* mod.rs:11162; matrix.rs:935,956,1346,1366,1386 */
GrB_Info GrB_Matrix_eWiseAdd_BinaryOp(GrB_Matrix C, GrB_Matrix Mask, GrB_BinaryOp accum, GrB_BinaryOp add, GrB_Matrix A,
                                      GrB_Matrix B, GrB_Descriptor desc); /* mod.rs:11316; matrix.rs:862 */
GrB_Info GrB_Matrix_eWiseMult_Semiring(GrB_Matrix C, GrB_Matrix Mask, GrB_BinaryOp accum, GrB_Semiring semiring,
                                       GrB_Matrix A, GrB_Matrix B, GrB_Descriptor desc); /* mod.rs:11228; matrix.rs:749,884 */
GrB_Info GrB_transpose(GrB_Matrix C, GrB_Matrix Mask, GrB_BinaryOp accum, GrB_Matrix A,
                       GrB_Descriptor desc);                /* mod.rs:14013; matrix.rs:658,829,841 */
GrB_Info GrB_Matrix_apply(GrB_Matrix C, GrB_Matrix Mask, GrB_BinaryOp accum, GrB_UnaryOp op, GrB_Matrix A,
                          GrB_Descriptor desc);             /* mod.rs:12375; matrix.rs:913 */

/* ---- vectors (frontier / BFS outputs; graph/src/graph/graphblas/vector.rs:44-55) ---- */
GrB_Info GrB_Vector_new(GrB_Vector *v, GrB_Type type, GrB_Index n);
GrB_Info GrB_Vector_free(GrB_Vector *v);
GrB_Info GrB_Vector_size(GrB_Index *n, GrB_Vector v);
GrB_Info GrB_Vector_nvals(GrB_Index *nvals, GrB_Vector v);
GrB_Info GrB_Vector_setElement_BOOL(GrB_Vector v, bool x, GrB_Index i);     /* vector.rs:127, 507 */
GrB_Info GrB_Vector_setElement_UINT64(GrB_Vector w, uint64_t x, GrB_Index i); /* mod.rs:9158; vector.rs:442 */
GrB_Info GrB_Vector_removeElement(GrB_Vector v, GrB_Index i);                 /* mod.rs:9318; vector.rs:519 */
GrB_Info GrB_Vector_clear(GrB_Vector v);                                      /* mod.rs:8924; vector.rs:98 */
GrB_Info GrB_Vector_wait(GrB_Vector object, int waitmode);                    /* mod.rs:11072; vector.rs:134 */
GrB_Info GrB_Vector_resize(GrB_Vector w, GrB_Index nrows_new);                /* mod.rs:14062; vector.rs:494 */
/* vector iterator: positions 0 .. nvals-1 in ascending index order; seek / next return GxB_EXHAUSTED past the end
 * (mod.rs:14972-14997; vector.rs:553-594).  GxB_Iterator_get_UINT64 reads the current value. */
GrB_Info GxB_Vector_Iterator_attach(GxB_Iterator iterator, GrB_Vector v, GrB_Descriptor desc);
GrB_Index GxB_Vector_Iterator_getpmax(GxB_Iterator iterator);
GrB_Info GxB_Vector_Iterator_seek(GxB_Iterator iterator, GrB_Index p);
GrB_Info GxB_Vector_Iterator_next(GxB_Iterator iterator);
GrB_Index GxB_Vector_Iterator_getp(GxB_Iterator iterator);
GrB_Index GxB_Vector_Iterator_getIndex(GxB_Iterator iterator);
GrB_Info GrB_Vector_extractElement_INT64(int64_t *x, GrB_Vector v, GrB_Index i);
GrB_Info GrB_Vector_extractElement_BOOL(bool *x, GrB_Vector v, GrB_Index i);
GrB_Info GrB_Vector_extractTuples_INT64(GrB_Index *I, int64_t *X, GrB_Index *nvals, GrB_Vector v);
GrB_Info GrB_Vector_extractTuples_BOOL(GrB_Index *I, bool *X, GrB_Index *nvals, GrB_Vector v);
GrB_Info GrB_Vector_setElement_FP64(GrB_Vector w, double x, GrB_Index i);
GrB_Info GrB_Vector_extractElement_FP64(double *x, GrB_Vector v, GrB_Index i);
GrB_Info GrB_Vector_extractTuples_FP64(GrB_Index *I, double *X, GrB_Index *nvals, GrB_Vector v);  /* algo_procedures.rs: extract_vector_f64 */
GrB_Info GrB_Matrix_build_FP64(GrB_Matrix C, const GrB_Index *I, const GrB_Index *J, const double *X, GrB_Index nvals, GrB_BinaryOp dup);
GrB_Info GrB_Matrix_extractTuples_FP64(GrB_Index *I, GrB_Index *J, double *X, GrB_Index *nvals, GrB_Matrix A);
/* w<mask> = u*A / A*u: one frontier step over ANY_PAIR (mod.rs:11173, 11184); over GrB_PLUS_TIMES_SEMIRING_FP64 / GxB_PLUS_SECOND_FP64
 * the FP64 mxv (no mask; accum NULL or GrB_PLUS_FP64; fixed summation order, rel 1e-12 against the sequential oracle) */
GrB_Info GrB_vxm(GrB_Vector w, GrB_Vector mask, GrB_BinaryOp accum, GrB_Semiring semiring, GrB_Vector u, GrB_Matrix A,
                 GrB_Descriptor desc);
GrB_Info GrB_mxv(GrB_Vector w, GrB_Vector mask, GrB_BinaryOp accum, GrB_Semiring semiring, GrB_Matrix A, GrB_Vector u,
                 GrB_Descriptor desc);

/* ---- row iterator (matrix.rs:1471-1605) -- real symbols, as bindgen declares them ---- */
GrB_Info GxB_Iterator_new(GxB_Iterator *it);                                      /* mod.rs:14848 */
GrB_Info GxB_Iterator_free(GxB_Iterator *it);                                     /* mod.rs:15084 */
GrB_Info GxB_rowIterator_attach(GxB_Iterator it, GrB_Matrix A, GrB_Descriptor d); /* mod.rs:14875 */
GrB_Index GxB_rowIterator_kount(GxB_Iterator it);                                 /* mod.rs:14882 */
GrB_Info GxB_rowIterator_seekRow(GxB_Iterator it, GrB_Index row);                 /* mod.rs:14885 */
GrB_Info GxB_rowIterator_nextRow(GxB_Iterator it);                                /* mod.rs:14897 */
GrB_Info GxB_rowIterator_nextCol(GxB_Iterator it);                                /* mod.rs:14900 */
GrB_Index GxB_rowIterator_getRowIndex(GxB_Iterator it);                           /* mod.rs:14903 */
GrB_Index GxB_rowIterator_getColIndex(GxB_Iterator it);                           /* mod.rs:14906 */
uint64_t GxB_Iterator_get_UINT64(GxB_Iterator it);                                /* mod.rs:15024 */
bool GxB_Iterator_get_BOOL(GxB_Iterator it);

/* ---- serialization (matrix.rs:428-546 Encode/Decode via GxB_Container; vector.rs:150-420) ---- */
#define GrB_NAME 10                 /* mod.rs:2879 */
#define GxB_JIT_C_NAME 7041         /* mod.rs:2897 */
#define GxB_MAX_NAME_LEN 128        /* mod.rs:158 */
extern GrB_Type GrB_UINT32;         /* mod.rs:541 (type of the `i` payload vector when ncols <= 2^32) */
/* 608 bytes, field offsets as asserted in mod.rs:14165-14237; the reference copies the struct bytes into its RDB stream
 * (pointer fields are nulled and re-created on decode, matrix.rs:455-470). */
struct GxB_Container_struct {
    uint64_t nrows, ncols;
    int64_t nrows_nonempty, ncols_nonempty;
    uint64_t nvals;
    uint64_t u64_future[11];
    int32_t format, orientation, header_arena;
    uint32_t u32_future[13];
    GrB_Vector p, h, b, i, x;
    GrB_Vector vector_future[11];
    GrB_Matrix Y;
    GrB_Matrix matrix_future[15];
    bool iso, jumbled;
    bool bool_future[30];
    void *void_future[16];
};
typedef struct GxB_Container_struct *GxB_Container;                                            /* mod.rs:14238 */
GrB_Info GxB_Container_new(GxB_Container *Container);                                          /* mod.rs:14240 */
GrB_Info GxB_Container_free(GxB_Container *Container);                                         /* mod.rs:15081 */
/* A's content moves into the container's p/h/b/i/x vectors (row-major sparse or hypersparse, see grb_api.cu); A is left empty */
GrB_Info GxB_unload_Matrix_into_Container(GrB_Matrix A, GxB_Container Container, GrB_Descriptor desc);  /* mod.rs:14264 */
/* A takes type, dimensions and content from the container (validated: the payload may come from GRAPH.RESTORE) */
GrB_Info GxB_load_Matrix_from_Container(GrB_Matrix A, GxB_Container Container, GrB_Descriptor desc);    /* mod.rs:14250 */
/* full (dense) payload vectors <-> raw arrays; arrays handed out are allocated with GxB_init's malloc */
GrB_Info GxB_Vector_load(GrB_Vector V, void **X, GrB_Type type, uint64_t n, uint64_t X_memsize, int handling,
                         GrB_Descriptor desc);                                                  /* mod.rs:14278 */
GrB_Info GxB_Vector_unload(GrB_Vector V, void **X, GrB_Type *type, uint64_t *n, uint64_t *X_memsize, int *handling,
                           GrB_Descriptor desc);                                                /* mod.rs:14289 */
/* opaque blob of a (sparse) vector: Tensor's multi-edge id lists (vector.rs:150-195); round-trips through this library only */
GrB_Info GxB_Vector_serialize(void **blob_handle, GrB_Index *blob_size, GrB_Vector u, GrB_Descriptor desc);   /* mod.rs:14717 */
GrB_Info GxB_Vector_deserialize(GrB_Vector *w, GrB_Type type, const void *blob, GrB_Index blob_size,
                                GrB_Descriptor desc);                                           /* mod.rs:14768 */
GrB_Info GrB_Type_get_String(GrB_Type type, char *value, int field);                            /* mod.rs:10503 */
GrB_Info GxB_Type_from_name(GrB_Type *type, const char *type_name);                             /* mod.rs:8075 */

/* ---- LAGraph subset (lagraph_bindings.rs:160-188, lagraphx_bindings.rs:585-594) ---- */
typedef enum { LAGraph_ADJACENCY_UNDIRECTED = 0, LAGraph_ADJACENCY_DIRECTED = 1, LAGraph_KIND_UNKNOWN = -1 } LAGraph_Kind;
typedef struct LAGraph_Graph_struct {
    GrB_Matrix A;
    LAGraph_Kind kind;
    GrB_Matrix AT;
    GrB_Vector out_degree, in_degree;
    int is_symmetric_structure;
    int64_t nself_edges;
    GrB_Scalar emin;
    int emin_state;
    GrB_Scalar emax;
    int emax_state;
} *LAGraph_Graph;
int LAGraph_Init(char *msg);
int LAGraph_Finalize(char *msg);
int LAGraph_New(LAGraph_Graph *G, GrB_Matrix *A, LAGraph_Kind kind, char *msg);
int LAGraph_Delete(LAGraph_Graph *G, char *msg);
int LAGr_BreadthFirstSearch_Extended(GrB_Vector *level, GrB_Vector *parent, LAGraph_Graph G, GrB_Index src,
                                     int64_t max_level, int64_t dest, bool many_expected, char *msg);
/* algo.pageRank (algo_procedures.rs:744-752; lagraph_bindings.rs:549-558): FP64 on the device, centrality = full GrB_FP64 vector */
int LAGraph_Cached_AT(LAGraph_Graph G, char *msg);
int LAGraph_Cached_OutDegree(LAGraph_Graph G, char *msg);
int LAGr_PageRank(GrB_Vector *centrality, int *iters, LAGraph_Graph G, float damping, float tol, int itermax, char *msg);
/* algo.WCC (algo_procedures.rs:838-846; lagraph_bindings.rs:521-526): component(i) = smallest vertex id of i's component (dense) */
int LAGr_ConnectedComponents(GrB_Vector *component, LAGraph_Graph G, char *msg);
/* algo.labelPropagation (algo_procedures.rs:1232-1237; lagraphx_bindings.rs:218-223): label(i) after <= itermax synchronous rounds
   of "most frequent neighbour label, smallest on ties" from label(i) = i (LDBC Graphalytics CDLP); full GrB_UINT64 vector */
int LAGraph_cdlp(GrB_Vector *CDLP_handle, LAGraph_Graph G, int itermax, char *msg);

/* ---- B200 extensions ---- */
#define B200_LOC_HOST 0
#define B200_LOC_DEVICE 1
/* Import a CSR (rowptr u64[nrows+1], col u32[nnz] ascending per row, val u64[nnz] or NULL) from host
 * or device memory; arrays are copied.  Plays the role of GxB_load_Matrix_from_Container. */
GrB_Info B200_Matrix_import_CSR(GrB_Matrix *A, GrB_Type type, GrB_Index nrows, GrB_Index ncols, const uint64_t *Ap,
                                const uint32_t *Aj, const uint64_t *Ax, int location);
/* Export into caller buffers sized by GrB_Matrix_nrows / nvals (Ax may be NULL). */
GrB_Info B200_Matrix_export_CSR(GrB_Matrix A, uint64_t *Ap, uint32_t *Aj, uint64_t *Ax, int location);
/* Export as a row-major packed bitmap: word `bits_out[i * words_per_row + (j >> 6)]` has bit (j & 63) set iff A(i,j) is an
 * entry; words_per_row must be ceil(ncols / 64) and every word of the nrows x words_per_row array is written.  The
 * interchange format for dense results (SuiteSparse keeps such matrices in GxB_BITMAP form and exports them with
 * GxB_Matrix_export_BitmapR; this is the 1-bit-per-slot equivalent): a frontier chain result that is still in device
 * bit-matrix form is exported without ever building its CSR.  Cheaper than CSR when nvals * 32 > nrows * ncols. */
GrB_Info B200_Matrix_export_bitmap(GrB_Matrix A, uint64_t *bits_out, uint64_t words_per_row, uint64_t *nvals_out, int location);
/* Same hand-off without blocking: the device-to-host copy runs on a second stream, so the caller can submit the next
 * batch's GrB_mxm calls while it is in flight (bits_out should be pinned).  A may be modified or freed right away; bits_out
 * is complete once B200_Ticket_wait returns.  Every ticket must be waited on exactly once. */
typedef struct B200_Ticket_opaque *B200_Ticket;
GrB_Info B200_Matrix_export_bitmap_async(GrB_Matrix A, uint64_t *bits_out, uint64_t words_per_row, B200_Ticket *ticket);
GrB_Info B200_Ticket_wait(B200_Ticket *ticket);
/* Borrow the device-resident CSR (valid until A is next modified or freed). */
/* digest3 = { nvals, sum mix(row << 32 | col), sum mix(key + GOLD * (CSR position + 1)) }: a multi-GB result is compared with the
 * oracle's through three numbers; sensitive to any changed entry and to the order inside a row */
GrB_Info B200_Matrix_digest(GrB_Matrix A, uint64_t *digest3);
GrB_Info B200_Matrix_device_view(GrB_Matrix A, const uint64_t **Ap, const uint32_t **Aj, const uint64_t **Ax);
/* Pre-build the cached transpose mirror used by the pull direction (done lazily otherwise). */
GrB_Info B200_Matrix_prepare(GrB_Matrix A, int want_transpose);
/* Synthetic Graph500-style RMAT adjacency (a,b,c,d=.57,.19,.19,.05), dedupe + no self loops, built on
 * the device.  Benchmark / test input only. */
GrB_Info B200_Matrix_rmat(GrB_Matrix *A, int scale, uint64_t edge_factor, uint64_t seed);
/* Row block [lo,hi) of B200_Matrix_rmat's matrix ((hi-lo) x 2^scale, rows local-indexed); by_col == 1 gives the same
 * block of the TRANSPOSE, by_col == 2 of the symmetrised strictly-lower-triangular matrix L = tril(A u A') (config 4).  Every rank regenerates the counter-based edge stream and keeps what it owns. */
GrB_Info B200_Matrix_rmat_block(GrB_Matrix *A, int scale, uint64_t edge_factor, uint64_t seed, uint64_t lo, uint64_t hi,
                                int by_col);
/* 1-D row-partitioned BFS building blocks: one level = expand the owned part of the frontier into an
 * n-bit `disc` bitmap; the caller all-gathers the bitmaps (NCCL); merge ORs them, updates `visited`, assigns levels to
 * owned vertices and emits the next owned frontier.  counters2[0] = next owned frontier size, [1] = new vertices
 * globally (0 => done).  Every pointer is a DEVICE pointer except counters2 / edges_out (host). */
GrB_Info B200_bfs_dist_expand(GrB_Matrix Alocal, uint64_t row_lo, const uint32_t *frontier, uint64_t nf, const uint64_t *visited,
                              uint64_t *disc, uint64_t nwords, uint64_t *edges_out);
GrB_Info B200_bfs_dist_merge(const uint64_t *gathered, int nranks, uint64_t nwords, uint64_t *visited, uint64_t row_lo,
                             uint64_t row_hi, int32_t *level_local, int32_t lvl, uint32_t *next_frontier, uint64_t *counters2,
                             uint64_t *frontier_bits /* optional: receives the new frontier as an n-bit bitmap */);
/* bottom-up step: owned unvisited vertices scan their in-neighbours (row block of A') for a frontier member */
GrB_Info B200_bfs_dist_pull(GrB_Matrix ATlocal, uint64_t row_lo, const uint64_t *frontier_bits, const uint64_t *visited,
                            uint64_t *disc, uint64_t nwords, uint64_t *scanned_out);
GrB_Info B200_bfs_dist_parents(GrB_Matrix ATlocal, uint64_t row_lo, const int32_t *level_full, int64_t *parent_local);

/* GRAPH.BULK's edge load into an EMPTY relationship tensor (src/commands/bulk_insert.rs:497 -> graph.rs:2062 ->
 * Tensor::set_all_from_slices, tensor.rs:333-447) as one device-side build: *fwd = the forward UINT64 matrix (value = the pair's
 * edge id, or UINT64_MAX = tensor.rs:206's MULTI_EDGE when the pair has several edges); *multi_keys / *multi_ids = (src << 32 | dst,
 * edge id) for every edge of every multi-edge pair, sorted by (key, id) -- the entries of `me`.  The two arrays come from GxB_init's
 * malloc (NULL when *nmulti == 0) and are the caller's to free. */
GrB_Info B200_Tensor_bulk_build(GrB_Matrix *fwd, GrB_Index **multi_keys, GrB_Index **multi_ids, GrB_Index *nmulti, GrB_Index nrows,
                                GrB_Index ncols, const GrB_Index *srcs, const GrB_Index *dsts, const GrB_Index *ids, GrB_Index n);
/* Batched point lookup: found[t] = 1 (and values[t] = A(I[t],J[t]) when `values` is non-NULL) iff the entry is stored.
 * ExpandInto's per-row Tensor::get (graph/src/runtime/ops/expand_into.rs:195-249 -> GrB_Matrix_extractElement_UINT64)
 * for a whole 1024-row batch in one device call; host arrays in, host arrays out. */
GrB_Info B200_Matrix_extract_pairs(GrB_Matrix A, const GrB_Index *I, const GrB_Index *J, GrB_Index n, uint8_t *found,
                                   uint64_t *values);
/* The coalesced traversal of CondTraverseOp::expand_batch as one call (cond_traverse.rs:600-608, 1264-1285): F(i, sources[i]) = 1,
 * F <- F * hops[0] * ... * hops[nhops-1] over GxB_ANY_PAIR_BOOL (square, equally sized operands), result rows to host memory as a
 * packed row-major bitmap (128-row slices, copies overlapped with the next slice's hops), as CSR, or whichever moves fewer bytes. */
#define B200_OUT_AUTO 0
#define B200_OUT_BITMAP 1
#define B200_OUT_CSR 2
GrB_Info B200_traverse_batch(const GrB_Index *sources, GrB_Index nsrc, const GrB_Matrix *hops, int nhops, int format,
                             uint64_t *out_bits, uint64_t words_per_row, uint64_t *out_p, uint32_t *out_j, uint64_t out_j_capacity,
                             uint64_t *nvals_out, uint64_t *flops_out, int *format_out);
/* Multi-source reachability in one call (CondVarLenTraverse with emit_path = false, cond_var_len_traverse.rs:196;
 * AllShortestPaths' BFS phase, all_shortest_paths.rs:7-25): row i of *reached (new nsrc x n BOOL matrix) = the vertices reachable
 * from sources[i] by 1..max_hops edges of the square matrix A (max_hops < 0: to the fixed point); include_sources adds the
 * zero-length walk.  Levels are F<!R,replace,struct> = F*A then R = R u F, in device frontier form; *levels_out = levels run. */
GrB_Info B200_reach_batch(GrB_Matrix *reached, const GrB_Index *sources, GrB_Index nsrc, GrB_Matrix A, int64_t max_hops,
                          int include_sources, int64_t *levels_out);
GrB_Info B200_sync(void);
GrB_Info B200_pool_trim(void); /* hand the caching allocator's free device blocks back to the driver */
void *B200_stream(void); /* the cudaStream_t every kernel of this library is launched on */
/* stats: "launches", "lib_launches", "last_flops", "total_flops", "last_path", "h2d_bytes", "d2h_bytes" */
uint64_t B200_get_stat(const char *name);
void B200_reset_stats(void);
/* with option "timing"=1 the library brackets its main kernels with CUDA events on its stream; this returns the
 * accumulated device time (ms), launch count and algorithmic bytes of one kernel family since the last reset:
 * "bits_pull", "bits_pull_long", "bits_push", "heavy_accumulate", "bits_fill", "bits_count".  0 on success. */
int B200_kernel_stats(const char *name, double *ms, uint64_t *launches, uint64_t *bytes);
/* options: "bits_mode" (-1 auto,0 off,1 on), "pull_mode" (-1 auto,0 push,1 pull), "small_cap",
 * "bitmap_budget", "bits_min_flops", "timing" (0/1), "sync_after_op" (0/1) */
GrB_Info B200_set_option(const char *name, int64_t value);
const char *B200_last_error(void);
/* single-source BFS straight into caller buffers (int64 level/parent per vertex, -1 = unreached) */
GrB_Info B200_bfs(GrB_Matrix A, GrB_Index src, int64_t max_level, int64_t *level, int64_t *parent, int location,
                  uint64_t *edges_traversed);
/* what one BFS did: levels by direction, exchange volume / time of the partitioned form (device-event timed) */
typedef struct {
    uint64_t depth, edges, td_levels, bu_levels, sparse_levels, exchanges, exchanged_bytes;
    double device_ms, exchange_ms;
} B200_BfsInfo;
/* BFS with the `dest` early exit of LAGr_BreadthFirstSearch_Extended (lagraphx_bindings.rs:585-594; -1 = none).  Runs the
 * direction-optimising engine when the transpose mirror exists (B200_Matrix_prepare(A, 1)), else the top-down kernel. */
GrB_Info B200_bfs_ex(GrB_Matrix A, GrB_Index src, int64_t max_level, int64_t dest, int64_t *level, int64_t *parent, int location,
                     B200_BfsInfo *info);
/* 1-D row-block partitioned BFS over NCCL (BASELINE config 5).  Rank 0 obtains the 128-byte id, the caller ships it
 * to the other ranks, every rank calls B200_comm_init; world == 1 needs no id.  NCCL is resolved with dlopen at run time. */
typedef struct B200_Comm_opaque *B200_Comm;
GrB_Info B200_comm_unique_id(uint8_t *id128);
GrB_Info B200_comm_init(B200_Comm *comm, int rank, int world, const uint8_t *id128);
GrB_Info B200_comm_free(B200_Comm *comm);
/* Alocal / ATlocal: rows [row_lo, row_lo + nloc) of the n x n adjacency matrix and of its transpose (global column ids); blocks
 * are ceil(n / world) rounded up to a multiple of 64.  level_local / parent_local: int64[nloc].  Collective over `comm`. */
GrB_Info B200_bfs_partitioned(GrB_Matrix Alocal, GrB_Matrix ATlocal, uint64_t n, uint64_t row_lo, B200_Comm comm, GrB_Index src,
                              int64_t max_level, int64_t dest, int64_t *level_local, int64_t *parent_local, int location,
                              B200_BfsInfo *info);

#ifdef __cplusplus
}
#endif
#endif /* B200GRB_H */
