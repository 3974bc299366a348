// cond_traverse.hpp -- C++ mirror of the batched F*A path of the reference's CondTraverse operator
// (graph/src/runtime/ops/cond_traverse.rs:452-751), reduced to what sits on the GraphBLAS boundary:
//   collect (row_i, src) + source-label pre-filter   :554-589
//   F = Matrix<bool>(|batch|, n); F.build(rows, srcs) :600-601
//   for hop: delta_lmxm_into(F)                        :602-605   (matrix.rs:1317-1402)
//   F.wait(); for (row_i, dest) in F.iter()            :608, 644
//   destination-label post-filter                      :646-652
// Batch/column plumbing (gather, set_column, null padding) is runtime code outside the hot path.
#pragma once
#include "tensor.hpp"

namespace fdb {

// Graph::node_has_label_id (graph.rs:1057-1065): label matrices are n x n diagonal (graph.rs:1191)
inline bool node_has_label(const VersionedMatrix &label, uint64_t id) { return label.get(id, id); }

struct ExpandResult {
    std::vector<uint64_t> row_idx; // index into the input batch (active_subset position)
    std::vector<uint64_t> dest;    // destination node id
};

static const size_t BATCH_SIZE = 1024; // graph/src/runtime/batch.rs:81

// TraversalMatrix (cond_traverse.rs:64-90): the adjacency VersionedMatrix<bool>, or one relationship type's Tensor
struct TraversalMatrix {
    const VersionedMatrix *vm = nullptr;
    const Tensor *t = nullptr;
    TraversalMatrix(const VersionedMatrix *v) : vm(v) {}
    TraversalMatrix(const Tensor *x) : t(x) {}
    uint64_t ncols() const { return vm ? vm->ncols() : t->fwd_m().ncols(); }
    void delta_lmxm_into(Matrix<bool> &f) const {                // cond_traverse.rs:77-85
        if (vm) f.delta_lmxm(vm->m(), vm->dp(), vm->dm());
        else f.delta_lmxm(t->fwd_m(), t->fwd_dp(), t->fwd_dm());
    }
};

// fuse_dst_labels: label matrices are diagonal (graph.rs:1191), so the destination filter F*A*L_dst
// (graph.rs:2600-2626) is one more delta_lmxm against the label's VersionedMatrix -- a column filter in frontier form on
// the device -- instead of a node_has_label probe (1-3 extractElement calls) per output pair on the host.
inline ExpandResult expand_batch(const std::vector<uint64_t> &src_ids, const std::vector<TraversalMatrix> &hops,
                                 const std::vector<const VersionedMatrix *> &src_labels,
                                 const std::vector<const VersionedMatrix *> &dst_labels, bool fuse_dst_labels = false) {
    ExpandResult out;
    if (hops.empty() || src_ids.empty()) return out;
    uint64_t ncols = hops[0].ncols();
    std::vector<uint64_t> row_idx_buf, col_idx_buf;
    row_idx_buf.reserve(src_ids.size());
    col_idx_buf.reserve(src_ids.size());
    for (size_t i = 0; i < src_ids.size(); i++) {
        bool ok = true;
        for (const VersionedMatrix *l : src_labels) if (!node_has_label(*l, src_ids[i])) { ok = false; break; }
        if (!ok) continue;                       // pre-filter src by label (= L_src * F)
        row_idx_buf.push_back((uint64_t)i);
        col_idx_buf.push_back(src_ids[i]);
    }
    if (row_idx_buf.empty()) return out;
    Matrix<bool> f(src_ids.size(), ncols);
    f.build(row_idx_buf, col_idx_buf);
    for (const TraversalMatrix &h : hops) h.delta_lmxm_into(f);
    if (fuse_dst_labels)
        for (const VersionedMatrix *l : dst_labels) TraversalMatrix(l).delta_lmxm_into(f);
    f.wait();                                    // flush pending mxm work before attaching the row iterator
    auto it = f.iter(0, UINT64_MAX);
    std::tuple<uint64_t, uint64_t> t;
    while (it.next(t)) {
        uint64_t dest = std::get<1>(t);
        bool ok = true;
        if (!fuse_dst_labels)
            for (const VersionedMatrix *l : dst_labels) if (!node_has_label(*l, dest)) { ok = false; break; }
        if (!ok) continue;                       // post-filter final-hop dst label (= F * A * R_dst)
        out.row_idx.push_back(std::get<0>(t));
        out.dest.push_back(dest);
    }
    return out;
}

// The batch boundary on the way out (graph/src/runtime/batch.rs:81, 274-287, 832): the (parent row, destination) pairs leave the
// operator re-packed into batches of <= BATCH_SIZE rows -- a `Column::NodeIds` with the destinations plus a selection vector
// (`Vec<u16>`: index of the parent row in the INPUT batch) that `Batch::gather` uses to copy the parent's columns next to it.
// Pairs arrive in ascending (row, dest) order (the iterator contract), so every output batch is a contiguous, ordered slice.
struct OutBatch {
    std::vector<uint64_t> node_ids;    // Column::NodeIds(Vec<NodeId>)  (graph.rs:139: NodeId(u64))
    std::vector<uint16_t> selection;   // parent row per output row; input batches hold <= BATCH_SIZE rows, so u16 suffices
};
inline std::vector<OutBatch> repack(const ExpandResult &r, size_t batch = BATCH_SIZE) {
    std::vector<OutBatch> out;
    for (size_t i = 0; i < r.dest.size(); i += batch) {
        const size_t e = std::min(r.dest.size(), i + batch);
        OutBatch b;
        b.node_ids.assign(r.dest.begin() + i, r.dest.begin() + e);
        b.selection.reserve(e - i);
        for (size_t k = i; k < e; k++) {
            if (r.row_idx[k] >= BATCH_SIZE) throw std::runtime_error("repack: parent row outside a BATCH_SIZE input batch");
            b.selection.push_back((uint16_t)r.row_idx[k]);
        }
        out.push_back(std::move(b));
    }
    return out;
}

} // namespace fdb
