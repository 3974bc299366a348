// ops.cuh -- host-callable entry points of the kernel layer (all enqueue on b200::stream()).
#pragma once
#include "common.cuh"

namespace b200 {

// build.cu
void build_from_device_coo(const u64 *dI, const u64 *dJ, const u64 *dX, u64 n, u64 nrows, u64 ncols, DevCSR &out,
                           bool *index_error);
// GRAPH.BULK into an empty tensor: forward u64 CSR (edge id or the MULTI_EDGE sentinel) + the (pair key, id) list of multi-edge pairs
void tensor_bulk_build(const u64 *dI, const u64 *dJ, const u64 *dID, u64 n, u64 nrows, u64 ncols, DevCSR &fwd, DevBuf<u64> &mkeys,
                       DevBuf<u64> &mids, u64 *nmulti, bool *index_error);
void transpose_csr(const DevCSR &A, DevCSR &out, bool keep_values);
void rmat_csr(int scale, u64 edge_factor, u64 seed, DevCSR &out);
void rmat_block_csr(int scale, u64 edge_factor, u64 seed, u64 lo, u64 hi, int by_col, DevCSR &out);

// spgemm.cu : C = pattern(A*B) over ANY_PAIR, row-wise push (Gustavson family)
void spgemm_anypair(const DevCSR &A, const DevCSR &B, DevCSR &C, u64 *flops_out);
u64 spgemm_flops(const DevCSR &A, const DevCSR &B);

struct LongRows;
// bits.cu : frontier bit-matrix path for short-fat left operands (<= 1024 rows)
u32 bits_words_for(u64 nrows);
void bits_from_csr(const DevCSR &F, DevBits &X);
void bits_to_csr(const DevBits &X, DevCSR &C);
u64 bits_nvals(const DevBits &X);
// Y = F * A expanded straight from F's CSR (tiny frontiers); false = too much work for this path, nothing was done
// lr (optional): A's pull tables; when they carry a vertex order the result is written in it (permuted form)
bool bits_push_from_csr(const DevCSR &F, const DevCSR &A, DevBits &Y, u64 *flops_out, const LongRows *lr = nullptr);
void bits_naturalise(DevBits &X);                                // permuted form -> natural vertex order (no-op when natural)
bool csr_is_diagonal(const DevCSR &A);                          // square, every entry (i,i)
void bits_diag(const DevBits &X, const DevCSR &A, DevBits &Y, u64 *flops_out);   // Y = X * A for diagonal A
void bits_to_rowmajor(const DevBits &X, u64 *out, u64 wpr);   // out[row * wpr + (col >> 6)], every word written
void csr_to_rowmajor(const DevCSR &A, u64 *out, u64 wpr);     // out must be zero on entry
struct LongRows {   // per-matrix auxiliary data of the pull direction, cached with the transpose mirror
    bool built = false;
    // legacy 8-lane / merge-path kernels
    DevBuf<u32> rows; u64 n = 0; u64 maxdeg = 0;      // rows of A' longer than LONG_ROW
    DevBuf<u64> choff; u64 nchunks = 0;               // chunk table of those rows (prefix of ceil(len / LONG_CHUNK))
    DevBuf<u64> mp_r;                                  // merge-path coordinates of every 256th diagonal
    DevBuf<u32> m_row, m_len; DevBuf<u64> m_start; u64 nm = 0;   // mid rows (SMALL_ROW < len <= LONG_ROW), longest first
    DevBuf<u32> jp;                                    // relabelled col_idx of A' (all rows)
    // lane-split pull (pull_kernel = 5): mid rows and LONG_ROW-entry segments of long rows, longest first; bit 31 of s_len marks a
    // segment of a long row (merged with RED.OR)
    DevBuf<u32> s_row, s_len; DevBuf<u64> s_start; u64 ns = 0; bool seg_built = false;
    // hot-set packing: vertices with out-edges, by out-degree descending
    DevBuf<u32> vert, slot, hdeg; u64 n1 = 0; bool packed = false;   // hdeg[s] = out-degree of vert[s]
    // the same order extended to ALL vertices (sinks follow the n1 slots, by id): a frontier stored in it needs no packing pass
    std::shared_ptr<DevBuf<u32>> pvert, pperm; u64 perm_tag = 0;    // pvert[position] = vertex, pperm[vertex] = position
    // CSR-stream form for the pull kernel: short rows of A' (cols relabelled to slots), window -> first row,
    // and the long rows kept apart; built per frontier word count W
    u32 sW = 0; u64 swin = 0; bool s_packed = false;
    DevBuf<u64> rp_s; DevBuf<u32> jp_s, wstart; u64 nwin = 0, nnz_s = 0;
    DevBuf<u32> lrows, jp_l; DevBuf<u64> lrp; u64 nlong = 0, maxlong = 0;
    void clear() {
        s_row.release(); s_len.release(); s_start.release(); ns = 0; seg_built = false;
        built = false; rows.release(); n = 0; maxdeg = 0; choff.release(); nchunks = 0; mp_r.release(); jp.release(); m_row.release(); m_len.release(); m_start.release(); nm = 0;
        vert.release(); slot.release(); hdeg.release(); n1 = 0; packed = false; pvert.reset(); pperm.reset(); perm_tag = 0;
        sW = 0; swin = 0; s_packed = false; rp_s.release(); jp_s.release(); wstart.release(); nwin = 0; nnz_s = 0;
        lrows.release(); jp_l.release(); lrp.release(); nlong = 0; maxlong = 0;
    }
};
void build_hot_pack(const DevCSR &A, const DevCSR &AT, LongRows &lr);
void build_stream(const DevCSR &AT, LongRows &lr, u32 W);
void build_long_rows(const DevCSR &AT, LongRows &lr);
// Y = X * A.  AT (= A') + its long-row list enable the pull direction; may be null (push only).
void bits_hop(const DevBits &X, const DevCSR &A, const DevCSR *AT, LongRows *lr, DevBits &Y, u64 *flops_out,
              int *path_out);
void bits_prepare_pull(const DevCSR &A, const DevCSR &AT, LongRows &lr);
void bits_andnot(DevBits &Y, const DevBits &M); // Y &= ~M
void bits_or(DevBits &Y, const DevBits &Z);     // Y |= Z
void bits_copy(const DevBits &X, DevBits &Y);

// ewise.cu : search-based set algebra on sorted CSR rows
void csr_copy(const DevCSR &A, DevCSR &out, bool keep_values);
void ewise_union(const DevCSR &A, const DevCSR &B, bool keep_values, DevCSR &out); // overlap: B's value (SECOND)
void ewise_intersect(const DevCSR &A, const DevCSR &B, DevCSR &out);               // pattern only
// keep t in T iff (t in M [and M's value != 0 unless structural]) XOR comp
void filter_by_mask(const DevCSR &T, const DevCSR &M, bool comp, bool structural, DevCSR &out);
void csr_resize(const DevCSR &A, u64 nrows, u64 ncols, DevCSR &out); // grow/shrink dims (drops out-of-range)
// Z = pattern(A*B) restricted to M's structure (valued mask: entries of M with value 0 excluded unless structural)
void spgemm_masked(const DevCSR &A, const DevCSR &B, const DevCSR &M, bool structural, DevCSR &out, u64 flops = 0);

// bfs.cu
void bfs_run(const DevCSR &A, u64 src, i64 max_level, i64 *d_level, i64 *d_parent, u64 *edges_traversed);
void bfs_dist_expand(const DevCSR &Aloc, u64 row_lo, const u32 *frontier, u64 nf, const u64 *visited, u64 *disc, u64 nwords,
                     u64 *edges_out);
void bfs_dist_pull(const DevCSR &ATloc, u64 row_lo, const u64 *frontier, const u64 *visited, u64 *disc, u64 nwords,
                   u64 *scanned_out);
void bfs_dist_merge(const u64 *gathered, int P, u64 nwords, u64 *visited, u64 row_lo, u64 row_hi, int *level_local, int lvl,
                    u32 *next, u64 *host_counters, u64 *frontier_bits);
void bfs_dist_parents(const DevCSR &ATloc, u64 row_lo, const int *level_full, i64 *parent_local);

void csr_digest(const DevCSR &A, u64 *host_out3);   // {nnz, sum mix(key), sum mix(key + GOLD * (pos + 1))}: see oracle orc_digest
// bfs_do.cu : direction-optimising BFS, single GPU or 1-D row-block partitioned over NCCL
struct BfsComm;
struct BfsInfo {          // mirrored by B200_BfsInfo in include/b200grb.h
    u64 depth, edges, td_levels, bu_levels, sparse_levels, exchanges, exchanged_bytes;
    double device_ms, exchange_ms;
};
void comm_unique_id(unsigned char *id128);
BfsComm *comm_init(int rank, int world, const unsigned char *id128);
void comm_free(BfsComm *c);
int comm_rank(const BfsComm *c);
int comm_world(const BfsComm *c);
void comm_allgather(BfsComm *c, const void *send, void *recv, size_t bytes_per_rank);
void bfs_build_degrees(const DevCSR &Aloc, u64 n, u64 lo, u64 hi, BfsComm *comm, DevBuf<u32> &deg_all, u64 *total_edges);
void bfs_do(const DevCSR &Aloc, const DevCSR &ATloc, u64 n, u64 lo, u64 hi, const u32 *deg_all, u64 total_edges, BfsComm *comm,
            u64 src, i64 max_level, i64 dest, i64 *d_level, i64 *d_parent, BfsInfo *info);

// algo.cu : FP64 mxv (PLUS_TIMES / PLUS_SECOND) and PageRank
void mxv_fp64(const DevCSR &A, bool use_values, const double *x, const unsigned char *present, double *y, unsigned char *ypresent,
              double init, bool accum);
int pagerank(const DevCSR &A, const DevCSR &AT, double damping, double tol, int itermax, double *r);
int connected_components(const DevCSR &A, u64 *d_comp);     // symmetric pattern; d_comp[v] = smallest vertex id of v's component
int cdlp(const DevCSR &A, int itermax, u64 *d_label);         // synchronous label propagation, min label on ties; returns the rounds run

void probe_pairs(const DevCSR &A, const u64 *dI, const u64 *dJ, u64 n, unsigned char *d_found, u64 *d_val);

// hypersparse host form <-> dense device rowptr (ewise.cu)
void rowptr_from_hyper(const u64 *d_hrow, const u64 *d_hptr, u64 nvec, u64 nrows, u64 *d_p);
u64 hyper_from_rowptr(const u64 *d_p, u64 nrows, u64 nnz, DevBuf<u64> &d_hrow, DevBuf<u64> &d_hptr);
void widen_u32(const u32 *in, u64 *out, u64 n);
void narrow_u64(const u64 *in, u32 *out, u64 n);

// generic tiny helpers (ewise.cu)
void fill_u64(u64 *p, u64 v, u64 n);
void fill_i64(i64 *p, i64 v, u64 n);

} // namespace b200
