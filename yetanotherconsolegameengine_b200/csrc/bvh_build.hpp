// bvh_build.hpp — host-side binned-SAH tree construction that reproduces the reference's two builders
// bit for bit, so that a tree built here equals the tree the C# host would export:
//   top level : BVH.BuildRecursive      (ConsoleGame/RayTracing/Objects/BVH.cs:258-459), leaf <= 4
//   triangles : MeshBVH.BuildRecursive  (ConsoleGame/RayTracing/Objects/MeshBVH.cs:371-576), leaf <= 8
// The two differ only in how the partition pass re-derives the bin mapping (BVH.cs:394-396 vs
// MeshBVH.cs:511-513).  Tie-breaking of exact-t hits depends on node order and leaf order, hence the care.
// The fallback ordering is System.Array.Sort = dotnet/runtime's introsort, re-expressed below.
#pragma once
#include <cstdint>
#include <cmath>
#include <limits>
#include <utility>
#include <vector>

namespace ycge {

struct Aabb {
    float lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; k++) { lo[k] = std::numeric_limits<float>::infinity(); hi[k] = -std::numeric_limits<float>::infinity(); }
    }
    void grow(const Aabb &o) { // "Surround": strict comparisons, so NaN/equal never replace
        for (int k = 0; k < 3; k++) { if (o.lo[k] < lo[k]) lo[k] = o.lo[k]; }
        for (int k = 0; k < 3; k++) { if (o.hi[k] > hi[k]) hi[k] = o.hi[k]; }
    }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.0f * (dx * dy + dx * dz + dy * dz);
    }
};

struct BuildItem {
    int index;
    Aabb box;
    float c[3];
};

struct FlatTree { // the reference's SoA node arrays (BVH.cs:11-25)
    std::vector<float> min_x, min_y, min_z, max_x, max_y, max_z;
    std::vector<int32_t> left, right, start, count, leaf_index;
    int32_t root = -1;
    uint64_t sort_fallbacks = 0;
    int n_nodes() const { return (int)left.size(); }
};

namespace detail {

inline int single_compare(float a, float b) { // System.Single.CompareTo
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    if (a != a) return (b != b) ? 0 : -1;
    return 1;
}
// IEEE 754-2019 minimum/maximum as MathF.Min/Max implement them
inline float net_min(float a, float b) {
    if (a != b) return (a != a) ? a : (a < b ? a : b);
    return std::signbit(a) ? a : b;
}
inline float net_max(float a, float b) {
    if (a != b) return (a != a) ? a : (b < a ? a : b);
    return std::signbit(b) ? a : b;
}

// Array.Sort(keys, start, count, comparer) for BuildItem keyed on one centroid axis.
class NetIntroSort {
  public:
    NetIntroSort(BuildItem *base, int axis) : a_(base), ax_(axis) {}
    void run(int start, int count) {
        if (count < 2) return;
        int lg = 0;
        for (unsigned v = (unsigned)count; v > 1; v >>= 1) lg++;
        intro(start, count, 2 * (lg + 1));
    }

  private:
    BuildItem *a_;
    int ax_;
    int cmp(const BuildItem &p, const BuildItem &q) const { return single_compare(p.c[ax_], q.c[ax_]); }
    void order2(int i, int j) { if (cmp(a_[i], a_[j]) > 0) std::swap(a_[i], a_[j]); }
    void insertion(int lo, int n) {
        for (int i = 0; i + 1 < n; i++) {
            BuildItem t = a_[lo + i + 1];
            int j = i;
            for (; j >= 0 && cmp(t, a_[lo + j]) < 0; j--) a_[lo + j + 1] = a_[lo + j];
            a_[lo + j + 1] = t;
        }
    }
    void sift(int lo, int i, int n) {
        BuildItem d = a_[lo + i - 1];
        while (i <= n / 2) {
            int ch = 2 * i;
            if (ch < n && cmp(a_[lo + ch - 1], a_[lo + ch]) < 0) ch++;
            if (!(cmp(d, a_[lo + ch - 1]) < 0)) break;
            a_[lo + i - 1] = a_[lo + ch - 1];
            i = ch;
        }
        a_[lo + i - 1] = d;
    }
    void heap(int lo, int n) {
        for (int i = n / 2; i >= 1; i--) sift(lo, i, n);
        for (int i = n; i > 1; i--) { std::swap(a_[lo], a_[lo + i - 1]); sift(lo, 1, i - 1); }
    }
    int partition(int lo, int n) {
        int hi = n - 1, mid = hi >> 1;
        order2(lo, lo + mid);
        order2(lo, lo + hi);
        order2(lo + mid, lo + hi);
        BuildItem pivot = a_[lo + mid];
        std::swap(a_[lo + mid], a_[lo + hi - 1]);
        int l = 0, r = hi - 1;
        while (l < r) {
            while (cmp(a_[lo + (++l)], pivot) < 0) {}
            while (cmp(pivot, a_[lo + (--r)]) < 0) {}
            if (l >= r) break;
            std::swap(a_[lo + l], a_[lo + r]);
        }
        if (l != hi - 1) std::swap(a_[lo + l], a_[lo + hi - 1]);
        return l;
    }
    void intro(int lo, int n, int depth) {
        while (n > 1) {
            if (n <= 16) {
                if (n == 2) { order2(lo, lo + 1); return; }
                if (n == 3) { order2(lo, lo + 1); order2(lo, lo + 2); order2(lo + 1, lo + 2); return; }
                insertion(lo, n);
                return;
            }
            if (depth == 0) { heap(lo, n); return; }
            depth--;
            int p = partition(lo, n);
            intro(lo + p + 1, n - (p + 1), depth);
            n = p;
        }
    }
};

struct TmpNode { Aabb box; int32_t left, right, start, count; };

class SahBuilder {
  public:
    SahBuilder(int leaf_size, bool mesh_variant) : leaf_(leaf_size), mesh_(mesh_variant) {}
    std::vector<TmpNode> nodes;
    std::vector<int32_t> leaves;
    uint64_t fallbacks = 0;

    int32_t build(BuildItem *arr, int start, int count) {
        if (count <= 0) return -1;
        if (count <= leaf_) return make_leaf(arr, start, count);
        const int mid = choose_mid(arr, start, count);
        int32_t me = (int32_t)nodes.size();
        nodes.push_back(TmpNode{});
        int32_t l = build(arr, start, mid - start);
        int32_t r = build(arr, mid, start + count - mid);
        nodes[me] = join(l, r);
        return me;
    }
    bool is_leaf_range(int count) const { return count <= leaf_; }
    // an inner node from its finished children (BVH.cs:430-457)
    TmpNode join(int32_t l, int32_t r) const {
        TmpNode cur{};
        cur.left = l; cur.right = r; cur.start = 0; cur.count = 0;
        if (l >= 0 && r >= 0) {
            for (int k = 0; k < 3; k++) {
                cur.box.lo[k] = net_min(nodes[l].box.lo[k], nodes[r].box.lo[k]);
                cur.box.hi[k] = net_max(nodes[l].box.hi[k], nodes[r].box.hi[k]);
            }
        } else cur.box = nodes[l >= 0 ? l : r].box;
        return cur;
    }
    // where a range of more than leaf-size items is cut: items [start, mid) go left (the range is permuted in place)
    int choose_mid(BuildItem *arr, int start, int count) {
        float cmin[3] = {arr[start].c[0], arr[start].c[1], arr[start].c[2]};
        float cmax[3] = {cmin[0], cmin[1], cmin[2]};
        for (int i = start + 1; i < start + count; i++)
            for (int k = 0; k < 3; k++) { // per item: three mins, then three maxes — order irrelevant for independent scalars
                float v = arr[i].c[k];
                if (v < cmin[k]) cmin[k] = v;
                if (v > cmax[k]) cmax[k] = v;
            }
        float ext[3] = {cmax[0] - cmin[0], cmax[1] - cmin[1], cmax[2] - cmin[2]};
        int best_axis = 0;
        if (ext[1] > ext[0] && ext[1] >= ext[2]) best_axis = 1;
        else if (ext[2] > ext[0] && ext[2] >= ext[1]) best_axis = 2;

        int split = -1;
        float best_cost = std::numeric_limits<float>::infinity();
        for (int ax = 0; ax < 3; ax++) {
            if (!(ext[ax] > 0.0f)) continue;
            const float origin = cmin[ax], inv = 1.0f / ext[ax];
            int cnt[kBins];
            Aabb bin[kBins];
            for (int b = 0; b < kBins; b++) { cnt[b] = 0; bin[b].reset(); }
            for (int i = start; i < start + count; i++) {
                int b = (int)((arr[i].c[ax] - origin) * inv * (kBins - 1));
                if (b < 0) b = 0;
                if (b >= kBins) b = kBins - 1;
                cnt[b]++;
                bin[b].grow(arr[i].box);
            }
            int lcnt[kBins], rcnt[kBins];
            float larea[kBins], rarea[kBins];
            Aabb run;
            run.reset();
            int acc = 0;
            for (int b = 0; b < kBins; b++) {
                if (cnt[b] > 0) run.grow(bin[b]);
                acc += cnt[b];
                lcnt[b] = acc;
                larea[b] = run.area();
            }
            run.reset();
            acc = 0;
            for (int b = kBins - 1; b >= 0; b--) {
                if (cnt[b] > 0) run.grow(bin[b]);
                acc += cnt[b];
                rcnt[b] = acc;
                rarea[b] = run.area();
            }
            for (int b = 0; b + 1 < kBins; b++) {
                int lc = lcnt[b], rc = rcnt[b + 1];
                if (lc == 0 || rc == 0) continue;
                float cost = larea[b] * lc + rarea[b + 1] * rc;
                if (cost < best_cost) { best_cost = cost; best_axis = ax; split = b; }
            }
        }

        int mid;
        if (split < 0) {
            mid = median_split(arr, start, count, best_axis);
        } else {
            float origin, inv;
            bool degenerate = false;
            if (mesh_) {
                origin = cmin[best_axis];
                inv = 1.0f / ext[best_axis];
            } else { // BVH.cs:394-396: mapping re-derived from the first and last item of the (unsorted) range
                origin = arr[start].c[best_axis];
                float e = arr[start + count - 1].c[best_axis] - origin;
                inv = e != 0.0f ? 1.0f / e : 0.0f;
                degenerate = (inv == 0.0f);
            }
            int i0 = start, i1 = start + count - 1;
            while (i0 <= i1) {
                int b0 = degenerate ? 0 : (int)((arr[i0].c[best_axis] - origin) * inv * (kBins - 1));
                if (b0 <= split) i0++;
                else { std::swap(arr[i0], arr[i1]); i1--; }
            }
            mid = i0;
            if (mid == start || mid == start + count) mid = median_split(arr, start, count, best_axis);
        }
        return mid;
    }

  private:
    static constexpr int kBins = 16;
    int leaf_;
    bool mesh_;
    int32_t make_leaf(BuildItem *arr, int start, int count) {
        TmpNode leaf{};
        leaf.box = arr[start].box;
        for (int i = 1; i < count; i++) leaf.box.grow(arr[start + i].box);
        leaf.left = leaf.right = -1;
        leaf.start = (int32_t)leaves.size();
        leaf.count = count;
        for (int i = 0; i < count; i++) leaves.push_back(arr[start + i].index);
        nodes.push_back(leaf);
        return (int32_t)nodes.size() - 1;
    }
    int median_split(BuildItem *arr, int start, int count, int axis) {
        fallbacks++;
        NetIntroSort(arr, axis).run(start, count);
        return start + (count >> 1);
    }
};

} // namespace detail

namespace detail { inline void to_flat(SahBuilder &b, FlatTree &out); }
// Build the reference tree over `items` (which is permuted in place, as the reference permutes its Item[]).
inline void build_reference_tree(std::vector<BuildItem> &items, int leaf_size, bool mesh_variant, FlatTree &out) {
    out = FlatTree();
    if (items.empty()) return;
    detail::SahBuilder b(leaf_size, mesh_variant);
    b.nodes.reserve(2 * items.size());
    b.leaves.reserve(items.size());
    out.root = b.build(items.data(), 0, (int)items.size());
    detail::to_flat(b, out);
}
namespace detail {
inline void to_flat(SahBuilder &b, FlatTree &out) {
    size_t n = b.nodes.size();
    out.min_x.resize(n); out.min_y.resize(n); out.min_z.resize(n);
    out.max_x.resize(n); out.max_y.resize(n); out.max_z.resize(n);
    out.left.resize(n); out.right.resize(n); out.start.resize(n); out.count.resize(n);
    for (size_t i = 0; i < n; i++) {
        const detail::TmpNode &t = b.nodes[i];
        out.min_x[i] = t.box.lo[0]; out.min_y[i] = t.box.lo[1]; out.min_z[i] = t.box.lo[2];
        out.max_x[i] = t.box.hi[0]; out.max_y[i] = t.box.hi[1]; out.max_z[i] = t.box.hi[2];
        out.left[i] = t.left; out.right[i] = t.right; out.start[i] = t.start; out.count[i] = t.count;
    }
    out.leaf_index = std::move(b.leaves);
    out.sort_fallbacks = b.fallbacks;
}
} // namespace detail

// Triangle -> build item, MeshBVH.TryComputeBounds (MeshBVH.cs:351-361) and the centroid rule (MeshBVH.cs:55-57).
inline BuildItem triangle_item(int index, const float *abc) {
    BuildItem it;
    it.index = index;
    const float pad = 1e-4f;
    for (int k = 0; k < 3; k++) {
        float a = abc[k], b = abc[3 + k], c = abc[6 + k];
        it.box.lo[k] = detail::net_min(a, detail::net_min(b, c)) - pad;
        it.box.hi[k] = detail::net_max(a, detail::net_max(b, c)) + pad;
        it.c[k] = 0.5f * (it.box.lo[k] + it.box.hi[k]);
    }
    return it;
}

} // namespace ycge
