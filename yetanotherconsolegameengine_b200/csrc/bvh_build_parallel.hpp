// bvh_build_parallel.hpp — the same trees as bvh_build.hpp, built on several host threads.
//
// The reference builds its trees on one thread (BVH.cs:258-459, MeshBVH.cs:371-576); a 280 k-triangle mesh costs ~0.4 s at
// every scene switch.  The recursion only couples a node to its children through (a) the in-place permutation of the item
// range, which the two children split disjointly, and (b) the numbering: nodes in pre-order (the parent's index is reserved
// before its children are built), leaf references in the order the leaves are created.  So the top of the tree is cut
// exactly as the serial builder cuts it (the two halves of a cut on two threads), the subtrees below are built independently
// into their own arrays, and the pieces are concatenated in pre-order with their indices shifted — node for node, leaf for leaf the tree
// of build_reference_tree (tests/test_bvh_builder_literal.py compares the two on whole meshes).
// Host mirror only; libycge.so keeps the serial builder.
#pragma once
#include "bvh_build.hpp"

#include <atomic>
#include <memory>
#include <mutex>
#include <thread>

namespace ycge {
namespace detail {

struct PlanNode { // the top of the tree: either cut further (left / right) or handed to a task
    int start = 0, count = 0, task = -1;
    std::unique_ptr<PlanNode> left, right;
};

class ParallelSah {
  public:
    ParallelSah(int leaf_size, bool mesh_variant, int task_items) : leaf_(leaf_size), mesh_(mesh_variant), task_items_(task_items) {}

    void run(BuildItem *arr, int n, unsigned threads, FlatTree &out) {
        for (par_depth_ = 0; (1u << par_depth_) < threads; par_depth_++) {} // the top levels are cut on 1, 2, 4, ... threads
        std::unique_ptr<PlanNode> root = plan(arr, 0, n, 0);
        std::atomic<size_t> next{0};
        auto work = [&] {
            for (size_t t; (t = next.fetch_add(1)) < tasks_.size();) {
                Task &k = tasks_[t];
                k.b.reset(new SahBuilder(leaf_, mesh_));
                k.b->build(arr, k.start, k.count);
            }
        };
        std::vector<std::thread> pool;
        for (unsigned i = 1; i < threads && i < tasks_.size(); i++) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
        SahBuilder all(leaf_, mesh_);
        all.nodes.reserve(2 * (size_t)n);
        all.leaves.reserve((size_t)n);
        all.fallbacks = top_fallbacks_.load();
        out.root = emit(*root, all);
        to_flat(all, out);
    }

  private:
    struct Task { int start, count; std::unique_ptr<SahBuilder> b; };
    int leaf_;
    bool mesh_;
    int task_items_;
    int par_depth_ = 0;
    std::atomic<uint64_t> top_fallbacks_{0}; // sort fallbacks taken while cutting the top ranges
    std::mutex tasks_mutex_;
    std::vector<Task> tasks_; // in no particular order: a plan node names its task by index

    std::unique_ptr<PlanNode> plan(BuildItem *arr, int start, int count, int depth) {
        std::unique_ptr<PlanNode> p(new PlanNode());
        p->start = start; p->count = count;
        if (count <= task_items_ || count <= leaf_) {
            std::lock_guard<std::mutex> lock(tasks_mutex_);
            p->task = (int)tasks_.size();
            tasks_.push_back(Task{start, count, nullptr});
            return p;
        }
        SahBuilder cut(leaf_, mesh_); // choose_mid reads and permutes only this range
        const int mid = cut.choose_mid(arr, start, count);
        top_fallbacks_ += cut.fallbacks;
        if (depth < par_depth_) { // the two halves are disjoint: cut the left one on another thread
            std::thread other([&] { p->left = plan(arr, start, mid - start, depth + 1); });
            struct Join { std::thread &t; ~Join() { if (t.joinable()) t.join(); } } guard{other}; // joined also when the sibling call throws
            p->right = plan(arr, mid, start + count - mid, depth + 1);
        } else {
            p->left = plan(arr, start, mid - start, depth + 1);
            p->right = plan(arr, mid, start + count - mid, depth + 1);
        }
        return p;
    }

    // pre-order concatenation: what SahBuilder::build would have pushed, in the order it would have pushed it
    int32_t emit(const PlanNode &p, SahBuilder &all) {
        if (p.task >= 0) {
            SahBuilder &b = *tasks_[(size_t)p.task].b;
            const int32_t node_base = (int32_t)all.nodes.size(), leaf_base = (int32_t)all.leaves.size();
            for (TmpNode t : b.nodes) {
                if (t.count > 0) t.start += leaf_base;
                else { if (t.left >= 0) t.left += node_base; if (t.right >= 0) t.right += node_base; }
                all.nodes.push_back(t);
            }
            all.leaves.insert(all.leaves.end(), b.leaves.begin(), b.leaves.end());
            all.fallbacks += b.fallbacks;
            return node_base; // a subtree's own root is its first node
        }
        const int32_t me = (int32_t)all.nodes.size();
        all.nodes.push_back(TmpNode{});
        const int32_t l = emit(*p.left, all);
        const int32_t r = emit(*p.right, all);
        all.nodes[(size_t)me] = all.join(l, r);
        return me;
    }
};

} // namespace detail

// build_reference_tree on `threads` host threads (0: one per core, at most 32); ranges of at most `task_items` items become tasks.
inline void build_reference_tree_parallel(std::vector<BuildItem> &items, int leaf_size, bool mesh_variant, FlatTree &out, unsigned threads = 0, int task_items = 0) {
    out = FlatTree();
    if (items.empty()) return;
    if (threads == 0) threads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (task_items <= 0) task_items = std::max(1024, (int)(items.size() / (8 * (size_t)threads)));
    if (threads == 1 || (int)items.size() <= task_items) { build_reference_tree(items, leaf_size, mesh_variant, out); return; }
    if (task_items < leaf_size) task_items = leaf_size;
    detail::ParallelSah(leaf_size, mesh_variant, task_items).run(items.data(), (int)items.size(), threads, out);
}

} // namespace ycge
