// cond_var_len_traverse.hpp -- C++ mirror of the trail enumerator of the reference's CondVarLenTraverse operator
// (graph/src/runtime/ops/cond_var_len_traverse.rs:152-386): an explicit-stack DFS over relationship-unique paths (trails) of
// min_hops..max_hops edges from one start node, emitting (from, to) per qualifying trail in the reference's order.
//   begin_start_node :152-183   0-hop emission, initial frame
//   advance          :196-386   pop a frame; enumerate its adjacency (cached per node, :241-244) in storage order; skip used
//                               edges (:248-251); direction handling (:253-265); will_emit / will_continue (:319-326); emissions
//                               of one frame come out in adjacency order (:374-377), continuing frames are pushed in adjacency
//                               order and therefore popped in reverse
// What sits on the GraphBLAS boundary is the adjacency fetch: Graph::get_node_relationships_by_type walks the relationship
// tensors with the row iterator (forward rows for outgoing edges, the transpose mirror for incoming ones: tensor.rs:886,
// matrix.rs:1471-1605) -- here Tensor::fwd_iter / matrix_t over the same C calls.  Path materialisation, edge-attribute and
// WHERE filters are runtime code above this layer and are not mirrored.  The multiplicity-insensitive fast path (emit_path off,
// min_hops <= 1) is the reach loop of grb.py: multi_source_reach / the frontier kernels, not this DFS.
#pragma once
#include "tensor.hpp"
#include <unordered_map>

namespace fdb {

enum class EdgeDirection { Outgoing, Incoming, Both };
struct Edge { uint64_t src, dst, id; };

// Graph::get_node_relationships_by_type for one relationship tensor: outgoing edges of `node` from the forward tensor (ascending
// destination, then ascending edge id), incoming ones through the transpose mirror (ascending source), both for Both.
inline std::vector<Edge> node_relationships(const Tensor &t, uint64_t node, EdgeDirection dir) {
    std::vector<Edge> out;
    if (dir != EdgeDirection::Incoming) {
        auto it = t.fwd_iter(node, node);
        std::tuple<uint64_t, uint64_t, uint64_t> e;
        while (it.next(e))
            for (uint64_t id : t.get(std::get<0>(e), std::get<1>(e))) out.push_back({std::get<0>(e), std::get<1>(e), id});
    }
    if (dir != EdgeDirection::Outgoing) {
        auto it = t.matrix_t().iter(node, node);
        std::tuple<uint64_t, uint64_t> e;
        while (it.next(e)) {
            if (dir == EdgeDirection::Both && std::get<1>(e) == node) continue;      // a self-loop was listed with the outgoing half
            for (uint64_t id : t.get(std::get<1>(e), node)) out.push_back({std::get<1>(e), node, id});
        }
    }
    return out;
}

struct VarLenResult { uint64_t from, to; std::vector<uint64_t> edges; };   // edges: the trail, filled when emit_path is set

// Adj: callable (node, EdgeDirection) -> std::vector<Edge>, the adjacency fetch (the only part that touches GraphBLAS)
template <class Adj>
class VarLenIterT {
  public:
    VarLenIterT(Adj adj, uint64_t min_hops, uint64_t max_hops, bool reversed = false, bool bidirectional = false,
                int64_t dest_id = -1, bool emit_path = false)
        : fetch_(adj), min_(min_hops), max_(max_hops), reversed_(reversed), bidir_(bidirectional), dest_(dest_id), emit_path_(emit_path) {}

    void begin_start_node(uint64_t start) {                       // :152-183
        start_ = start;
        stack_.clear(); buf_.clear();
        if (min_ == 0 && (dest_ < 0 || (uint64_t)dest_ == start)) buf_.push_back({start, start, {}});
        stack_.push_back({start, {}, 0});
    }
    bool next(VarLenResult &out) {                                // Iterator::next :388-
        while (true) {
            if (!buf_.empty()) { out = std::move(buf_.back()); buf_.pop_back(); return true; }
            if (stack_.empty()) return false;
            advance();
        }
    }

  private:
    struct Frame { uint64_t node; std::vector<uint64_t> used; uint64_t depth; };
    void advance() {                                              // :196-386
        const EdgeDirection dir = bidir_ ? EdgeDirection::Both : reversed_ ? EdgeDirection::Incoming : EdgeDirection::Outgoing;
        while (!stack_.empty()) {
            Frame f = std::move(stack_.back());
            stack_.pop_back();
            const uint64_t hop = f.depth + 1;
            if (hop > max_) continue;
            auto found = adj_.find(f.node);                      // adjacency lists are cached per node (:241-244)
            if (found == adj_.end()) found = adj_.emplace(f.node, fetch_(f.node, dir)).first;
            std::vector<std::pair<uint64_t, uint64_t>> scratch;   // (edge id, neighbour)
            for (const Edge &e : found->second) {
                if (std::find(f.used.begin(), f.used.end(), e.id) != f.used.end()) continue;   // relationship uniqueness
                bool ok = false;
                uint64_t nb = 0;
                if (reversed_) { if (e.dst == f.node) { ok = true; nb = e.src; } }
                else if (e.src == f.node) { ok = true; nb = e.dst; }
                else if (bidir_ && e.dst == f.node) { ok = true; nb = e.src; }
                if (ok) scratch.push_back({e.id, nb});
            }
            for (auto &pr : scratch) {
                const uint64_t dest = pr.second;
                const bool will_emit = hop >= min_ && (dest_ < 0 || (uint64_t)dest_ == dest);
                const bool will_continue = hop < max_;
                if (!will_emit && !will_continue) continue;
                std::vector<uint64_t> used = f.used;
                used.push_back(pr.first);
                const uint64_t from = reversed_ ? dest : start_, to = reversed_ ? start_ : dest;
                if (will_emit) buf_.push_back({from, to, emit_path_ ? used : std::vector<uint64_t>()});
                if (will_continue) stack_.push_back({dest, std::move(used), hop});
            }
            if (!buf_.empty()) { std::reverse(buf_.begin(), buf_.end()); return; }   // pop() then yields adjacency order (:374-377)
        }
    }
    Adj fetch_;
    uint64_t min_, max_;
    bool reversed_, bidir_;
    int64_t dest_;
    bool emit_path_;
    uint64_t start_ = 0;
    std::vector<Frame> stack_;
    std::vector<VarLenResult> buf_;
    std::unordered_map<uint64_t, std::vector<Edge>> adj_;
};

struct TensorAdjacency {
    const Tensor *t;
    std::vector<Edge> operator()(uint64_t node, EdgeDirection dir) const { return node_relationships(*t, node, dir); }
};
struct VarLenIter : VarLenIterT<TensorAdjacency> {
    VarLenIter(const Tensor &t, uint64_t min_hops, uint64_t max_hops, bool reversed = false, bool bidirectional = false,
               int64_t dest_id = -1, bool emit_path = false)
        : VarLenIterT<TensorAdjacency>(TensorAdjacency{&t}, min_hops, max_hops, reversed, bidirectional, dest_id, emit_path) {}
};

} // namespace fdb
