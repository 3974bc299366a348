// common.cuh -- device-memory, stream, launch-accounting and error plumbing shared by every
// translation unit of libb200grb.so (the B200-native GraphBLAS traversal backend).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <stdexcept>
#include <string>
#include <vector>
#include <memory>
#include <cstdio>

namespace b200 {

typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

struct CudaError : std::runtime_error {
    cudaError_t code;
    CudaError(cudaError_t c, const char *file, int line)
        : std::runtime_error(std::string(cudaGetErrorName(c)) + ": " + cudaGetErrorString(c) + " at " + file + ":" +
                             std::to_string(line)),
          code(c) {}
};
struct GrbError : std::runtime_error {
    int info;
    GrbError(int i, const std::string &m) : std::runtime_error(m), info(i) {}
};

#define CUDA_TRY(expr)                                                    \
    do {                                                                  \
        cudaError_t e__ = (expr);                                         \
        if (e__ != cudaSuccess) throw ::b200::CudaError(e__, __FILE__, __LINE__); \
    } while (0)

// ---- global context (one process per GPU; one library stream) -------------------------------
struct Context {
    bool ready = false;
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    std::atomic<u64> launches{0};     // kernels of THIS library launched (bench.py gpu_launches)
    std::atomic<u64> lib_launches{0}; // CUB primitives launched on our behalf (sort / scan)
    // last-op statistics (read through B200_get_stat)
    std::atomic<u64> last_flops{0}, total_flops{0}, last_path{0};
    std::atomic<u64> h2d_bytes{0}, d2h_bytes{0};
    // tunables (B200_set_option)
    i64 opt_bits_mode = -1;        // -1 auto, 0 never use the bit-frontier path, 1 always when legal
    i64 opt_pull_mode = -1;        // -1 auto, 0 push only, 1 pull only
    i64 opt_small_cap = 4096;      // rows with <= this many flops use the shared-memory sort path
    i64 opt_bitmap_budget = (i64)2 << 30; // bytes of global bitmap scratch per heavy-row wave
    i64 opt_bits_min_flops = 1 << 22;     // auto mode: use bit-frontier when flops >= this
    i64 opt_sync_after_op = 0;
    i64 opt_timing = 0;
    i64 opt_pull_kernel = 5;       // 5 = lane-split degree-binned (default); 4 = degree-binned, one lane per vertex record; 0..3 = earlier kernels
    i64 opt_perm_push = 1;         // a CSR frontier pushed through a prepared matrix lands in that matrix's hot-set order (no packing pass before the next pull)
    i64 opt_small_split = 0;       // small-row pull kernel: 1 = split the vertex record across lanes like the segment kernel, 0 = one lane per record
    i64 opt_l2_window = 0;         // bytes of the packed frontier's hot prefix kept L2-resident through a persisting access-policy window (0 = off)
    i64 opt_l2_reset = 0;          // cudaCtxResetPersistingL2Cache after each windowed pull
    i64 opt_count_kernel = 1;      // materialise count pass: 1 = vertical (carry-save) counters, 0 = transpose + popcount
    i64 opt_bfs_direction = 0;     // 0 = direction-optimising, 1 = top-down only, 2 = bottom-up only (test hooks)
    i64 opt_bfs_sparse_exchange = 1;   // partitioned BFS: ship discovered-vertex lists instead of bitmaps when the frontier is sparse
    u64 l2_persist_max = 0, l2_window_max = 0;   // device limits (bytes), read at bring-up
    i64 opt_early_exit = 1;        // stop a pull row once it holds the OR monoid's terminal value (exact)
    i64 opt_hints = -1;            // L2 createpolicy hints in the pull kernels: -1 = auto (on for W <= 2, where the gathers take a cache-hint operand; measured slower at W = 8: 2.27 vs 2.17 ms), 0 / 1 = off / on
    i64 opt_hot_bytes = 64 << 20;  // size of that hot prefix
    i64 opt_hot_pack = 1;          // gather through the degree-sorted, sink-free relabelling of the frontier
    i64 opt_fill_cap = 0;          // 0 = auto; >0 forces the materialise staging capacity (test hook)
    i64 opt_fill_kernel = 3;       // materialise: 3 = row-per-warp over 2048-vertex tiles, paired scans (default); 1 = row-per-warp, 1024-vertex tiles; 0 = block-staged lists; 2 = from kept masks
    i64 opt_fused_prep = 1;        // pull hops: pack the frontier, count flops / edges and OR it in one pass
    i64 opt_csr_push = 1;          // tiny CSR frontiers push straight from their entries (no O(n*W) bit-matrix passes)
    i64 opt_diag_filter = 1;       // frontier-form mxm by a diagonal (label) matrix runs as an elementwise column filter
    i64 opt_pull_grid = 0;         // CTAs per SM of the grid-stride pull kernels (0 = occupancy: one resident wave)
    i64 opt_unroll = 4;            // gathers in flight per lane in the 8-lane pull kernel
};
Context &ctx();
void ensure_init();

inline cudaStream_t stream() { return ctx().stream; }

// ---- optional per-kernel timing (CUDA events on the launch stream; B200_set_option("timing",1)) ----
enum TimedId { TK_BITS_PULL = 0, TK_BITS_PULL_LONG, TK_BITS_PUSH, TK_HEAVY_ACC, TK_SMALL_ROWS, TK_BITS_FILL, TK_BITS_COUNT,
               TK_BITMAP_EXPAND, TK_BFS_EXPAND, TK_UNION, TK_FILTER, TK_MXV, TK_MASKED, TK_TRANSPOSE, TK_COUNT_ };
const char *timed_name(int id);
void timed_begin(int id);
void timed_end(int id, u64 algorithmic_bytes) noexcept;
struct TimedScope {
    int id; u64 bytes;
    TimedScope(int i, u64 b) : id(i), bytes(b) { timed_begin(id); }
    ~TimedScope() { timed_end(id, bytes); }
};
// drains recorded events: total ms / launches / bytes per id
void timed_collect(double *ms, u64 *launches, u64 *bytes);
void timed_reset();

// ---- caching device allocator (prims.cu) ----------------------------------------------------------
// Size-class free lists over cudaMalloc.  Every buffer is only ever touched by work enqueued on the one library
// stream, so a freed block can be handed out again immediately (stream order guarantees the previous user is done
// before the next one starts): steady state has no driver allocation calls and no pool growth / fragmentation.
void *pool_alloc(size_t bytes);
void pool_free(void *p);
void pool_trim();
size_t pool_bytes_cached();

// ---- device buffers -------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T *ptr = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : ptr(o.ptr), n(o.n) { o.ptr = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); ptr = o.ptr; n = o.n; o.ptr = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count == 0) { ptr = nullptr; return; }
        ptr = (T *)pool_alloc(count * sizeof(T));
    }
    void release() {
        if (ptr) { pool_free(ptr); ptr = nullptr; }
        n = 0;
    }
    void zero() { if (n) CUDA_TRY(cudaMemsetAsync(ptr, 0, n * sizeof(T), stream())); }
    T *release_ownership() { T *p = ptr; ptr = nullptr; n = 0; return p; }
    size_t bytes() const { return n * sizeof(T); }
};

// Small device-to-host reads (scalars the host steers by: flops, nnz, bin counts) do not use the copy engine: a one-warp
// kernel stores them into mapped pinned memory and sync_stream() hands them to their destinations.  A DMA read-back would
// queue behind whatever the D2H engine is busy with -- e.g. the 100+ MB result bitmap of the previous batch slice travelling
// on the copy stream -- and stall the compute stream for milliseconds (prims.cu).
bool small_read(void *dst, const void *src, size_t bytes);   // false: not taken (too large / no slot), use a memcpy
void flush_small_reads();                                    // after the stream has been synchronised
void drop_small_reads();                                     // error paths: pending deliveries are abandoned, not copied
inline void sync_stream() { CUDA_TRY(cudaStreamSynchronize(stream())); flush_small_reads(); }

template <typename T>
inline T read_scalar(const T *dptr) {
    T v;
    if (!small_read(&v, dptr, sizeof(T))) CUDA_TRY(cudaMemcpyAsync(&v, dptr, sizeof(T), cudaMemcpyDeviceToHost, stream()));
    sync_stream();
    return v;
}
template <typename T>
inline void h2d(T *dst, const T *src, size_t n) {
    if (n) { CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, stream())); ctx().h2d_bytes += n * sizeof(T); }
}
template <typename T>
inline void d2h(T *dst, const T *src, size_t n) {
    if (n) {
        if (!small_read(dst, src, n * sizeof(T))) CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, stream()));
        ctx().d2h_bytes += n * sizeof(T);
    }
}
template <typename T>
inline void d2d(T *dst, const T *src, size_t n) {
    if (n) CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToDevice, stream()));
}

// Kernel launch with accounting.  Usage: LAUNCH(kernel, grid, block, smem, args...)
#define LAUNCH(kern, grid, block, smem, ...)                                   \
    do {                                                                       \
        ::b200::ctx().launches.fetch_add(1, std::memory_order_relaxed);        \
        kern<<<(grid), (block), (smem), ::b200::stream()>>>(__VA_ARGS__);      \
        CUDA_TRY(cudaGetLastError());                                          \
    } while (0)

inline u32 grid_for(u64 items, u32 per_block, u64 cap = 0x7fffffffULL) {
    u64 g = (items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (u32)g;
}

// ---- device CSR -----------------------------------------------------------------------------
// rowptr u64[nrows+1], col u32[nnz] ascending within a row, val u64[nnz] or null (iso/pattern).
struct DevCSR {
    u64 nrows = 0, ncols = 0, nnz = 0;
    DevBuf<u64> p;
    DevBuf<u32> j;
    DevBuf<u64> x; // empty => pattern-only (every stored value is true / 1)
    bool has_values() const { return x.ptr != nullptr; }
    void clear() { p.release(); j.release(); x.release(); nnz = 0; }
};

// frontier bit-matrix: an r x n boolean matrix (r <= 64*W) stored vertex-major:
// word[v*W + w] bit b  <=>  entry (row 64*w+b, col v).
struct DevBits {
    u64 nrows = 0, ncols = 0;
    u32 W = 0;
    DevBuf<u64> w;
    // Permuted form (bits.cu, "hot-set order"): the words of vertex v live at position pos(v) instead of v, pos = the degree-sorted
    // gather order of the matrix this frontier is about to be multiplied by (LongRows::pperm), so the pull of the next hop gathers
    // straight from it with no packing pass.  pvert[position] = vertex undoes it; perm_tag names the permutation (0 = natural).
    std::shared_ptr<DevBuf<u32>> pvert;
    u64 perm_tag = 0;
    bool valid() const { return w.ptr != nullptr; }
    bool permuted() const { return perm_tag != 0; }
    void clear() { w.release(); W = 0; pvert.reset(); perm_tag = 0; }
};

// ---- L2 residency hints (createpolicy descriptors; sm_80+) ---------------------------------------
// keep  : data that is gathered at random and must stay L2-resident (the frontier bit-matrix X)
// stream: data read or written exactly once (col_idx / rowptr streams, outputs)
#ifdef __CUDACC__
__device__ __forceinline__ u64 policy_keep() { u64 p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 policy_stream() { u64 p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
// [base, base+primary) evict_last (kept resident), the rest of [base, base+total) evict_first: the hot prefix of the
// degree-sorted packed frontier stays in L2 while its cold tail streams through
__device__ __forceinline__ u64 policy_range(const void *base, u32 primary_bytes, u32 total_bytes) {
    u64 p;
    asm volatile("createpolicy.range.L2::evict_last.L2::evict_first.b64 %0, [%1], %2, %3;" : "=l"(p) : "l"(base), "r"(primary_bytes), "r"(total_bytes));
    return p;
}
__device__ __forceinline__ ulonglong2 ld_v2_hint(const ulonglong2 *p, u64 pol) {
    ulonglong2 v;
    asm volatile("ld.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(pol));
    return v;
}
// 32-byte gathers (LDG.E.256, sm_100 + PTX 8.8): one L1TEX request per 4-word row instead of two
struct u64x4 { u64 a, b, c, d; };
__device__ __forceinline__ u64x4 ld_v4(const u64 *p) {
    u64x4 v;
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    return v;
}
__device__ __forceinline__ u64 ld_u64_hint(const u64 *p, u64 pol) {
    u64 v; asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ u64 ld_u64_stream(const u64 *p, u64 pol) {
    u64 v; asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ u32 ld_u32_stream(const u32 *p, u64 pol) {
    u32 v; asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ void st_u64_stream(u64 *p, u64 v, u64 pol) {
    asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_u32_stream(u32 *p, u32 v, u64 pol) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
#endif

// ---- primitive wrappers implemented in prims.cu (CUB scan / sort) ----------------------------
void exclusive_scan_u64(const u64 *in, u64 *out, size_t n);           // out[i] = sum in[0..i)
void exclusive_scan_u32_to_u64(const u32 *in, u64 *out, size_t n);
void sort_keys_u64(u64 *keys_in_out, size_t n, int end_bit);            // ascending, in place (uses temp)
void sort_pairs_u64(u64 *keys_in_out, u64 *vals_in_out, size_t n, int end_bit); // stable
u64 reduce_sum_u64(const u64 *in, size_t n);

} // namespace b200
