// bvh_device.cuh — SURVEY 8(f-2): MeshBVH.BuildRecursive (ConsoleGame/RayTracing/Objects/MeshBVH.cs:371-576) ON THE DEVICE,
// node for node the tree the reference (and bvh_build.hpp, its host restatement) builds: traversal order decides which of two
// hits at exactly the same t wins, so a "better" tree would not be bit-consistent.  A scene switch then costs a few
// milliseconds of GPU time instead of a ~100 ms host build (RaytraceEntity.cs:234-246 rebuilds on every switch).
//
// What makes the reference's builder parallel after all:
//   * every quantity a node needs from its items is a min / max / count (centroid bounds, bin boxes, bin counts, leaf boxes):
//     exact and order independent;
//   * the SAH sweep touches 3 x 16 bins: one thread, the reference's statements in the reference's order;
//   * the in-place two-pointer partition (MeshBVH.cs:511-530: scan from the left, swap a right-hand item with the current
//     last one, look at what came back) is a fixed permutation of the range that can be written down in closed form:
//     with r_0 < r_1 < .. the positions of the right-hand items and l_0 > l_1 > .. those of the left-hand items,
//     K = #{k : r_k < l_k} rounds complete, the scan from the left consumes the positions <= X, the swaps consume the rest from
//     the back, and
//         a left-hand item the scan meets stays where it is,            the k-th right-hand item the scan meets goes to l_{k-1} - 1,
//         a right-hand item pulled from the back moves down by one,     the k-th left-hand item pulled from the back goes to r_k
//     (l_{-1} = count; derivation and a brute-force check: tests/test_bvh_device_partition.py);
//   * the rare Array.Sort fallback (all centroids of a range equal) runs the same introsort as bvh_build.hpp on one thread;
//   * node numbers are the reference's (pre-order) and follow from subtree sizes once the topology exists.
// Nodes are processed from a device-wide work queue by persistent CTAs (a node is ready as soon as its parent has
// partitioned its range; no level-by-level launches, no host round trips).  The result is written straight in the device
// layout (64-byte pair nodes, leaf-ordered triangles, device_types.h).
// One documented difference: Surround's strict comparisons keep the FIRST of +0 / -0 it meets, an atomic min/max the
// smaller bit pattern; a box coordinate is min(vertices) - 1e-4 or max + 1e-4 and is never a zero in practice.
#pragma once
#include "device_types.h"

namespace ycge {

#define YCGE_DB_LEAF 8
#define YCGE_DB_BINS 16
#define YCGE_DB_THREADS 512

struct DbNode { int start, count, parent, left, right, pad0, pad1, pad2; };
struct DbCounters { unsigned int n_nodes, q_tail, q_head; int done_items; unsigned int fallbacks, pad0, pad1, pad2; };

struct DbArgs {
    int n;
    const float *abc;       // n x 9 vertex coordinates
    float *blo, *bhi, *cen; // [3][n]
    float *soa12;           // n x (A, e1, e2, normal)
    int *idx, *tmp, *pre, *rpos, *lpos;
    DbNode *nodes;          // capacity 2 n
    int *queue, *ready;     // capacity 2 n + grid
    DbCounters *cnt;
    // finalisation
    float *nbox;            // [2 n][6]
    int *nsize, *ninner, *nflag;
    PairNode *pairs;
    DevTri *tris;
    int *tri_id;
    TreeRoot *root;
};

__device__ __forceinline__ float db_net_min(float a, float b) { // MathF.Min (IEEE 754-2019 minimum), as bvh_build.hpp
    if (a != b) return (a != a) ? a : (a < b ? a : b);
    return (__float_as_int(a) < 0) ? a : b;
}
__device__ __forceinline__ float db_net_max(float a, float b) {
    if (a != b) return (a != a) ? a : (b < a ? a : b);
    return (__float_as_int(b) < 0) ? a : b;
}
__device__ __forceinline__ void db_atomic_min(float *addr, float v) {
    if (v >= 0.0f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void db_atomic_max(float *addr, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ int db_ld(const int *p) { return *reinterpret_cast<const volatile int *>(p); }

// MeshBVH ctor :83-97 + TryComputeBounds :351-361 + centroid :55-57
__global__ void db_items_kernel(DbArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const float *t = a.abc + 9 * (size_t)i;
    const float pad = 1e-4f;
    for (int k = 0; k < 3; k++) {
        const float va = t[k], vb = t[3 + k], vc = t[6 + k];
        const float lo = db_net_min(va, db_net_min(vb, vc)) - pad, hi = db_net_max(va, db_net_max(vb, vc)) + pad;
        a.blo[(size_t)k * a.n + i] = lo;
        a.bhi[(size_t)k * a.n + i] = hi;
        a.cen[(size_t)k * a.n + i] = 0.5f * (lo + hi);
    }
    float *d = a.soa12 + 12 * (size_t)i;
    d[0] = t[0]; d[1] = t[1]; d[2] = t[2];
    const float lx = t[3] - t[0], ly = t[4] - t[1], lz = t[5] - t[2];
    const float mx = t[6] - t[0], my = t[7] - t[1], mz = t[8] - t[2];
    d[3] = lx; d[4] = ly; d[5] = lz; d[6] = mx; d[7] = my; d[8] = mz;
    const float nnx = ly * mz - lz * my, nny = lz * mx - lx * mz, nnz = lx * my - ly * mx;
    const float invLen = 1.0f / db_net_max(1e-20f, __fsqrt_rn(nnx * nnx + nny * nny + nnz * nnz));
    d[9] = nnx * invLen; d[10] = nny * invLen; d[11] = nnz * invLen;
    a.idx[i] = i;
}

__global__ void db_init_kernel(DbArgs a) {
    DbNode r; r.start = 0; r.count = a.n; r.parent = -1; r.left = -1; r.right = -1; r.pad0 = r.pad1 = r.pad2 = 0;
    a.nodes[0] = r;
    a.cnt->n_nodes = 1; a.cnt->q_head = 0; a.cnt->fallbacks = 0;
    if (a.n > YCGE_DB_LEAF) { a.queue[0] = 0; a.ready[0] = 1; a.cnt->q_tail = 1; a.cnt->done_items = 0; }
    else { a.cnt->q_tail = 0; a.cnt->done_items = a.n; }
}

// System.Array.Sort on a range of the index array keyed by one centroid axis: dotnet/runtime's introsort as restated in
// bvh_build.hpp (NetIntroSort), item swaps become index swaps.  One thread.  (__host__ as well: tests/test_bvh_device_partition.py
// runs it on the CPU against NetIntroSort, permutation for permutation.)
struct DbIntroSort {
    int *a_; const float *key_;
    __host__ __device__ int cmp(int p, int q) const {
        const float x = key_[p], y = key_[q];
        if (x < y) return -1;
        if (x > y) return 1;
        if (x == y) return 0;
        if (x != x) return (y != y) ? 0 : -1;
        return 1;
    }
    __host__ __device__ void swp(int i, int j) { const int t = a_[i]; a_[i] = a_[j]; a_[j] = t; }
    __host__ __device__ void order2(int i, int j) { if (cmp(a_[i], a_[j]) > 0) swp(i, j); }
    __host__ __device__ void insertion(int lo, int n) {
        for (int i = 0; i + 1 < n; i++) {
            const int t = a_[lo + i + 1];
            int j = i;
            for (; j >= 0 && cmp(t, a_[lo + j]) < 0; j--) a_[lo + j + 1] = a_[lo + j];
            a_[lo + j + 1] = t;
        }
    }
    __host__ __device__ void sift(int lo, int i, int n) {
        const int d = a_[lo + i - 1];
        while (i <= n / 2) {
            int ch = 2 * i;
            if (ch < n && cmp(a_[lo + ch - 1], a_[lo + ch]) < 0) ch++;
            if (!(cmp(d, a_[lo + ch - 1]) < 0)) break;
            a_[lo + i - 1] = a_[lo + ch - 1];
            i = ch;
        }
        a_[lo + i - 1] = d;
    }
    __host__ __device__ void heap(int lo, int n) {
        for (int i = n / 2; i >= 1; i--) sift(lo, i, n);
        for (int i = n; i > 1; i--) { swp(lo, lo + i - 1); sift(lo, 1, i - 1); }
    }
    __host__ __device__ int partition(int lo, int n) {
        const int hi = n - 1, mid = hi >> 1;
        order2(lo, lo + mid);
        order2(lo, lo + hi);
        order2(lo + mid, lo + hi);
        const int pivot = a_[lo + mid];
        swp(lo + mid, lo + hi - 1);
        int l = 0, r = hi - 1;
        while (l < r) {
            while (cmp(a_[lo + (++l)], pivot) < 0) {}
            while (cmp(pivot, a_[lo + (--r)]) < 0) {}
            if (l >= r) break;
            swp(lo + l, lo + r);
        }
        if (l != hi - 1) swp(lo + l, lo + hi - 1);
        return l;
    }
    // the recursion of IntroSort on the right part, the loop on the left one; an explicit stack keeps it off the call stack
    __host__ __device__ void run(int start, int count) {
        if (count < 2) return;
        int lg = 0;
        for (unsigned v = (unsigned)count; v > 1; v >>= 1) lg++;
        int st_lo[72], st_n[72], st_d[72], sp = 0;
        st_lo[sp] = start; st_n[sp] = count; st_d[sp] = 2 * (lg + 1); sp++;
        while (sp > 0) {
            sp--;
            int lo = st_lo[sp], n = st_n[sp], depth = st_d[sp];
            while (n > 1) {
                if (n <= 16) {
                    if (n == 2) order2(lo, lo + 1);
                    else if (n == 3) { order2(lo, lo + 1); order2(lo, lo + 2); order2(lo + 1, lo + 2); }
                    else insertion(lo, n);
                    break;
                }
                if (depth == 0) { heap(lo, n); break; }
                depth--;
                const int p = partition(lo, n);
                // the reference sorts the RIGHT part first (recursive call), then continues with the left one; the two parts are
                // disjoint, so the order in which they are finished does not change the result: the right part is parked
                if (sp < 72) { st_lo[sp] = lo + p + 1; st_n[sp] = n - (p + 1); st_d[sp] = depth; sp++; }
                n = p;
            }
        }
    }
};

__global__ void __launch_bounds__(YCGE_DB_THREADS) db_build_kernel(DbArgs a) {
    const int T = YCGE_DB_THREADS, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ int s_node;
    __shared__ float s_red[2][3][YCGE_DB_THREADS / 32];
    __shared__ float s_cmin[3], s_cmax[3];
    __shared__ int s_bcnt[3][YCGE_DB_BINS];
    __shared__ float s_blo[3][YCGE_DB_BINS][3], s_bhi[3][YCGE_DB_BINS][3];
    __shared__ int s_split, s_axis, s_warp_sum[YCGE_DB_THREADS / 32], s_run, s_K;
    __shared__ float s_origin, s_inv;
    const float INF = __int_as_float(0x7F800000);
    for (;;) {
        if (tid == 0) {
            const unsigned int t = atomicAdd(&a.cnt->q_head, 1u);
            int node = -1;
            for (;;) {
                if (db_ld(&a.ready[t])) { node = db_ld(&a.queue[t]); break; }
                if (db_ld(&a.cnt->done_items) >= a.n) break;
                __nanosleep(200);
            }
            s_node = node;
        }
        __syncthreads();
        const int node = s_node;
        if (node < 0) return;
        __threadfence(); // the parent's writes to idx[] were fenced before `ready` was set
        const int start = a.nodes[node].start, count = a.nodes[node].count;
        int *idx = a.idx + start;
        // ---- centroid bounds (:398-405)
        float mn[3] = {INF, INF, INF}, mx[3] = {-INF, -INF, -INF};
        for (int i = tid; i < count; i += T) {
            const int id = idx[i];
            for (int k = 0; k < 3; k++) { const float v = a.cen[(size_t)k * a.n + id]; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
        }
        for (int k = 0; k < 3; k++) {
            for (int o = 16; o > 0; o >>= 1) { mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o)); mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o)); }
            if (lane == 0) { s_red[0][k][wid] = mn[k]; s_red[1][k][wid] = mx[k]; }
        }
        for (int k = tid; k < 3 * YCGE_DB_BINS; k += T) {
            const int ax = k / YCGE_DB_BINS, b = k % YCGE_DB_BINS;
            s_bcnt[ax][b] = 0;
            for (int j = 0; j < 3; j++) { s_blo[ax][b][j] = INF; s_bhi[ax][b][j] = -INF; }
        }
        __syncthreads();
        if (tid < 3) {
            float lo = INF, hi = -INF;
            for (int w = 0; w < T / 32; w++) { lo = fminf(lo, s_red[0][tid][w]); hi = fmaxf(hi, s_red[1][tid][w]); }
            s_cmin[tid] = lo; s_cmax[tid] = hi;
        }
        __syncthreads();
        float ext[3], org[3], inv[3];
        for (int k = 0; k < 3; k++) { org[k] = s_cmin[k]; ext[k] = s_cmax[k] - s_cmin[k]; inv[k] = 1.0f / ext[k]; }
        // ---- bins (:430-437): counts and boxes per axis
        for (int i = tid; i < count; i += T) {
            const int id = idx[i];
            float lo[3], hi[3];
            for (int j = 0; j < 3; j++) { lo[j] = a.blo[(size_t)j * a.n + id]; hi[j] = a.bhi[(size_t)j * a.n + id]; }
            for (int ax = 0; ax < 3; ax++) {
                if (!(ext[ax] > 0.0f)) continue;
                int b = (int)((a.cen[(size_t)ax * a.n + id] - org[ax]) * inv[ax] * (float)(YCGE_DB_BINS - 1));
                b = b < 0 ? 0 : (b >= YCGE_DB_BINS ? YCGE_DB_BINS - 1 : b);
                atomicAdd(&s_bcnt[ax][b], 1);
                for (int j = 0; j < 3; j++) { db_atomic_min(&s_blo[ax][b][j], lo[j]); db_atomic_max(&s_bhi[ax][b][j], hi[j]); }
            }
        }
        __syncthreads();
        // ---- SAH sweep (:407-493), one thread, the reference's statements
        if (tid == 0) {
            int best_axis = 0;
            if (ext[1] > ext[0] && ext[1] >= ext[2]) best_axis = 1;
            else if (ext[2] > ext[0] && ext[2] >= ext[1]) best_axis = 2;
            int split = -1;
            float best_cost = INF;
            for (int ax = 0; ax < 3; ax++) {
                if (!(ext[ax] > 0.0f)) continue;
                int lcnt[YCGE_DB_BINS], rcnt[YCGE_DB_BINS];
                float larea[YCGE_DB_BINS], rarea[YCGE_DB_BINS];
                float lo[3] = {INF, INF, INF}, hi[3] = {-INF, -INF, -INF};
                int acc = 0;
                for (int b = 0; b < YCGE_DB_BINS; b++) {
                    if (s_bcnt[ax][b] > 0) for (int j = 0; j < 3; j++) { if (s_blo[ax][b][j] < lo[j]) lo[j] = s_blo[ax][b][j]; if (s_bhi[ax][b][j] > hi[j]) hi[j] = s_bhi[ax][b][j]; }
                    acc += s_bcnt[ax][b];
                    lcnt[b] = acc;
                    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
                    larea[b] = 2.0f * (dx * dy + dx * dz + dy * dz);
                }
                for (int j = 0; j < 3; j++) { lo[j] = INF; hi[j] = -INF; }
                acc = 0;
                for (int b = YCGE_DB_BINS - 1; b >= 0; b--) {
                    if (s_bcnt[ax][b] > 0) for (int j = 0; j < 3; j++) { if (s_blo[ax][b][j] < lo[j]) lo[j] = s_blo[ax][b][j]; if (s_bhi[ax][b][j] > hi[j]) hi[j] = s_bhi[ax][b][j]; }
                    acc += s_bcnt[ax][b];
                    rcnt[b] = acc;
                    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
                    rarea[b] = 2.0f * (dx * dy + dx * dz + dy * dz);
                }
                for (int b = 0; b + 1 < YCGE_DB_BINS; b++) {
                    const int lc = lcnt[b], rc = rcnt[b + 1];
                    if (lc == 0 || rc == 0) continue;
                    const float cost = larea[b] * (float)lc + rarea[b + 1] * (float)rc;
                    if (cost < best_cost) { best_cost = cost; best_axis = ax; split = b; }
                }
            }
            s_split = split; s_axis = best_axis;
            s_origin = org[best_axis]; s_inv = 1.0f / ext[best_axis];
            s_run = 0;
        }
        __syncthreads();
        const int axis = s_axis, split = s_split;
        const float *key = a.cen + (size_t)axis * a.n;
        bool sorted_fallback = split < 0;
        int nL = 0;
        if (!sorted_fallback) {
            // ---- the two-pointer partition (:511-530) in closed form, see the header
            const float origin = s_origin, invE = s_inv;
            int *pre = a.pre + start, *rpos = a.rpos + start, *lpos = a.lpos + start, *tmp = a.tmp + start;
            for (int base = 0; base < count; base += T) {
                const int i = base + tid;
                int isR = 0;
                if (i < count) isR = ((int)((key[idx[i]] - origin) * invE * (float)(YCGE_DB_BINS - 1)) > split) ? 1 : 0;
                // block exclusive scan of isR, carried over tiles in s_run
                int v = isR;
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
                if (lane == 31) s_warp_sum[wid] = v;
                __syncthreads();
                int woff = 0;
                for (int w = 0; w < wid; w++) woff += s_warp_sum[w];
                const int run = s_run;
                const int kR = run + woff + v - isR; // right-hand items before i
                if (i < count) {
                    pre[i] = (kR << 1) | isR;
                    if (isR) rpos[kR] = i; else lpos[i - kR] = i;
                }
                __syncthreads();
                if (tid == T - 1) s_run = run + woff + v;
                __syncthreads();
            }
            const int nR = s_run;
            nL = count - nR;
            if (nL == 0 || nL == count) sorted_fallback = true; // (:531) cannot happen with the bin mapping of the sweep; kept for safety
            else {
                // K = #{k : r_k < l_k}, l_k = lpos[nL - 1 - k]
                const int m = nR < nL ? nR : nL;
                int c = 0;
                for (int k = tid; k < m; k += T) c += (rpos[k] < lpos[nL - 1 - k]) ? 1 : 0;
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                if (lane == 0) s_warp_sum[wid] = c;
                __syncthreads();
                if (tid == 0) { int K = 0; for (int w = 0; w < T / 32; w++) K += s_warp_sum[w]; s_K = K; }
                __syncthreads();
                const int K = s_K;
                const int lKm1 = K == 0 ? count : lpos[nL - K]; // l_{K-1}
                const int X = (K < nR && rpos[K] < lKm1) ? rpos[K] : lKm1 - 1;
                for (int i = tid; i < count; i += T) {
                    const int p = pre[i], isR = p & 1, kR = p >> 1;
                    int dest;
                    if (i <= X) dest = isR ? ((kR == 0 ? count : lpos[nL - kR]) - 1) : i;
                    else dest = isR ? i - 1 : rpos[nL - 1 - (i - kR)];
                    tmp[dest] = idx[i];
                }
                __syncthreads();
                for (int i = tid; i < count; i += T) idx[i] = tmp[i];
            }
        }
        if (sorted_fallback) { // Array.Sort + median (:495-503, :531-540)
            __syncthreads();
            if (tid == 0) {
                DbIntroSort s; s.a_ = idx; s.key_ = key;
                s.run(0, count);
                atomicAdd(&a.cnt->fallbacks, 1u);
            }
            nL = count >> 1;
        }
        __threadfence();
        __syncthreads();
        // ---- children
        if (tid == 0) {
            const unsigned int base = atomicAdd(&a.cnt->n_nodes, 2u);
            const int cs[2] = {start, start + nL}, cc[2] = {nL, count - nL};
            for (int k = 0; k < 2; k++) {
                DbNode ch; ch.start = cs[k]; ch.count = cc[k]; ch.parent = node; ch.left = -1; ch.right = -1; ch.pad0 = ch.pad1 = ch.pad2 = 0;
                a.nodes[base + k] = ch;
            }
            a.nodes[node].left = (int)base; a.nodes[node].right = (int)base + 1;
            __threadfence();
            for (int k = 0; k < 2; k++) {
                if (cc[k] <= YCGE_DB_LEAF) atomicAdd(&a.cnt->done_items, cc[k]);
                else {
                    const unsigned int slot = atomicAdd(&a.cnt->q_tail, 1u);
                    a.queue[slot] = (int)base + k;
                    __threadfence();
                    *reinterpret_cast<volatile int *>(&a.ready[slot]) = 1;
                }
            }
        }
        __syncthreads();
    }
}

// leaf boxes (:378-384), then inner boxes (:553-572), subtree sizes and inner-node counts bottom-up: the second child to
// arrive at a parent finishes it
__global__ void db_boxes_kernel(DbArgs a) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= (int)a.cnt->n_nodes) return;
    const DbNode nd = a.nodes[v];
    if (nd.left >= 0) return; // inner nodes are reached from below
    const float INF = __int_as_float(0x7F800000);
    float lo[3], hi[3];
    if (nd.count > 0) {
        for (int j = 0; j < 3; j++) { lo[j] = a.blo[(size_t)j * a.n + a.idx[nd.start]]; hi[j] = a.bhi[(size_t)j * a.n + a.idx[nd.start]]; }
        for (int i = 1; i < nd.count; i++) {
            const int id = a.idx[nd.start + i];
            for (int j = 0; j < 3; j++) { const float l = a.blo[(size_t)j * a.n + id], h = a.bhi[(size_t)j * a.n + id]; if (l < lo[j]) lo[j] = l; if (h > hi[j]) hi[j] = h; }
        }
    } else for (int j = 0; j < 3; j++) { lo[j] = INF; hi[j] = -INF; }
    float *b = a.nbox + 6 * (size_t)v;
    for (int j = 0; j < 3; j++) { b[j] = lo[j]; b[3 + j] = hi[j]; }
    a.nsize[v] = 1; a.ninner[v] = 0;
    __threadfence();
    int cur = nd.parent;
    while (cur >= 0) {
        if (atomicAdd(&a.nflag[cur], 1) == 0) return; // the sibling subtree is not finished yet
        __threadfence();
        const int l = a.nodes[cur].left, r = a.nodes[cur].right;
        const volatile float *bl = a.nbox + 6 * (size_t)l, *br = a.nbox + 6 * (size_t)r;
        float *bc = a.nbox + 6 * (size_t)cur;
        for (int j = 0; j < 3; j++) { bc[j] = db_net_min(bl[j], br[j]); bc[3 + j] = db_net_max(bl[3 + j], br[3 + j]); }
        a.nsize[cur] = 1 + db_ld(&a.nsize[l]) + db_ld(&a.nsize[r]);
        a.ninner[cur] = 1 + db_ld(&a.ninner[l]) + db_ld(&a.ninner[r]);
        __threadfence();
        cur = a.nodes[cur].parent;
    }
}

// pair nodes in the reference's node order: an inner node's pair index is its rank among the inner nodes in pre-order
// (ycge_lib.cu: flatten), i.e. the sum over the path from the root of 1 (the parent) + the inner nodes of a left sibling
__global__ void db_emit_kernel(DbArgs a) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= (int)a.cnt->n_nodes) return;
    const DbNode nd = a.nodes[v];
    auto leaf_ref = [&](const DbNode &x) { return ~(((x.count - 1) << 26) | x.start); };
    if (v == 0) {
        TreeRoot r;
        const float *b = a.nbox;
        for (int j = 0; j < 3; j++) { r.lo[j] = b[j]; r.hi[j] = b[3 + j]; }
        r.ref = nd.left >= 0 ? 0 : leaf_ref(nd);
        r.pad = 0;
        *a.root = r;
    }
    if (nd.left < 0) return;
    int pair = 0, cur = v;
    while (a.nodes[cur].parent >= 0) {
        const int p = a.nodes[cur].parent;
        pair += 1 + (a.nodes[p].right == cur ? a.ninner[a.nodes[p].left] : 0);
        cur = p;
    }
    const DbNode L = a.nodes[nd.left], R = a.nodes[nd.right];
    const int lref = L.left >= 0 ? pair + 1 : leaf_ref(L);
    const int rref = R.left >= 0 ? pair + 1 + a.ninner[nd.left] : leaf_ref(R);
    const float *lb = a.nbox + 6 * (size_t)nd.left, *rb = a.nbox + 6 * (size_t)nd.right;
    PairNode pn;
    pn.q0 = make_float4(lb[0], lb[1], lb[2], lb[3]);
    pn.q1 = make_float4(lb[4], lb[5], rb[0], rb[1]);
    pn.q2 = make_float4(rb[2], rb[3], rb[4], rb[5]);
    pn.q3 = make_float4(__int_as_float(lref), __int_as_float(rref), 0.0f, 0.0f);
    a.pairs[pair] = pn;
}

__global__ void db_tris_kernel(DbArgs a) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    const int t = a.idx[s];
    const float *q = a.soa12 + 12 * (size_t)t;
    DevTri d;
    d.t0 = make_float4(q[0], q[1], q[2], q[3]);
    d.t1 = make_float4(q[4], q[5], q[6], q[7]);
    d.t2 = make_float4(q[8], q[9], q[10], q[11]);
    a.tris[s] = d;
    a.tri_id[s] = t;
}

} // namespace ycge
