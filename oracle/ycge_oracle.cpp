/*
 * ycge_oracle.cpp — CPU restatement of YetAnotherConsoleGameEngine's per-frame ray tracing path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under yetanotherconsolegameengine_b200/ may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, as the checker and as the timed CPU baseline.
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCE TEXT (oracle/_ref).  The reference (C#/.NET 8) has no tests, golden
 * vectors or fixtures for this path and no .NET toolchain exists here, so it cannot be run as shipped.  Instead
 * oracle/ref_transpile.py rewrites its C# sources into C++ SYNTACTICALLY at build time (no arithmetic expression is touched;
 * nothing generated is committed) and oracle/Makefile compiles them into oracle/_ref/libycge_ref.so: Vec3, Ray, Material,
 * HitRecord, RaytraceSampler, every analytic primitive, Triangle (scalar path), MeshBVH and BVH (builders and traversal),
 * Mesh, VolumeGrid, Scene.Hit / Occluded, TraceFull, ComputeTransmittanceToLight, the BSDF helpers, MakeJitteredRay, the
 * verbatim head and tail of TryFlipAndBlit (ray generation, per-pixel trace loop; TAA, a-trous with the reference's own
 * buffer swap, exposure, cell loop), ToneMapper, Chexel, the ANSI-256 quantiser.  This file must equal that library bit for
 * bit -- whole frames, every plane (tests/test_reference_transpiled.py; the GPU is compared with it directly in
 * tests/test_gpu_parity.py::test_gpu_equals_the_transpiled_reference).  One substitution is shared by all sides: MathF.Exp /
 * Log / Pow / Sin / Cos / Tan = include/ycge_detmath.h (the platform libm behind them is not bit-reproducible).  Not covered by the
 * transpiled library: Texture.SampleBilinear (static textures), TemporalAA.ShouldResetHistory, the scene factories and MeshLoader
 * (host side) -- those keep the earlier pins: hand-derived known answers (tests/test_oracle_kat.py), a second
 * restatement in numpy binary32 (tests/test_oracle_trace_literal.py, test_oracle_render.py, test_bvh_builder_literal.py), scene
 * literals extracted by executing the C# factories' text (tools/extract_scene_literals.py).  This file follows the reference
 * source line by line (citations are relative to /root/reference/ConsoleGame/).
 *
 * Arithmetic rules: binary32 everywhere the reference uses float, evaluated in the reference's
 * order, no FMA contraction (build with -O2 -ffp-contract=off, no -ffast-math, x86-64 SSE2).
 * Transcendentals go through include/ycge_detmath.h (math_mode 0, bit-identical with the GPU) or glibc
 * libm (math_mode 1, to measure the "documented float tie" bucket).
 * Third-party algorithm on the path: System.Array.Sort (dotnet/runtime ArraySortHelper<T>.IntrospectiveSort,
 * .NET 8, restated from its published algorithm; version unpinned) in the BVH builders' fallback.
 */
#include "../include/ycge.h"
#include "../include/ycge_detmath.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace yo {

static int g_math_mode = 0; /* 0 = detmath, 1 = glibc */
static int g_sort_mode = 0; /* 0 = .NET introsort restatement, 1 = std::stable_sort */

static inline float m_sin(float x) { return g_math_mode ? sinf(x) : ycge_sinf(x); }
static inline float m_cos(float x) { return g_math_mode ? cosf(x) : ycge_cosf(x); }
static inline float m_tan(float x) { return g_math_mode ? tanf(x) : ycge_tanf(x); }
static inline float m_exp(float x) { return g_math_mode ? expf(x) : ycge_expf(x); }
static inline float m_log(float x) { return g_math_mode ? logf(x) : ycge_logf(x); }
static inline float m_pow(float x, float y) { return g_math_mode ? powf(x, y) : ycge_powf(x, y); }
static inline void m_sincos(float x, float *s, float *c) {
    if (g_math_mode) { *s = sinf(x); *c = cosf(x); } else ycge_sincosf(x, s, c);
}

/* MathF.Max / MathF.Min: IEEE 754-2019 maximum/minimum (NaN propagates, +0 > -0). */
static inline float MaxF(float a, float b) {
    if (a != b) { if (!(a != a)) return b < a ? a : b; return a; }
    return std::signbit(b) ? a : b;
}
static inline float MinF(float a, float b) {
    if (a != b) { if (!(a != a)) return a < b ? a : b; return a; }
    return std::signbit(a) ? a : b;
}
static inline float AbsF(float a) { return std::fabs(a); }
static inline float CopySignF(float a, float b) { return std::copysign(a, b); }
static inline float FloorF(float a) { return std::floor(a); }
static inline float SqrtF(float a) { return std::sqrt(a); }
static const float FloatMax = std::numeric_limits<float>::max();
static const float PosInf = std::numeric_limits<float>::infinity();
static const float NegInf = -std::numeric_limits<float>::infinity();

/* ---- Vec3 (RayTracing/Vec3.cs:6-128) ---- */
struct Vec3 {
    float X, Y, Z;
    Vec3() : X(0), Y(0), Z(0) {}
    Vec3(float x, float y, float z) : X(x), Y(y), Z(z) {}
    static Vec3 D(double x, double y, double z) { return Vec3((float)x, (float)y, (float)z); } /* Vec3.cs:21-26 */
    float Dot(const Vec3 &b) const { return X * b.X + Y * b.Y + Z * b.Z; }                        /* :74-77 */
    Vec3 Cross(const Vec3 &b) const { return Vec3(Y * b.Z - Z * b.Y, Z * b.X - X * b.Z, X * b.Y - Y * b.X); } /* :80-83 */
    Vec3 Normalized() const {                                                                     /* :98-107 */
        float lenSq = X * X + Y * Y + Z * Z;
        if (lenSq <= 0.0f) return *this;
        float invLen = 1.0f / SqrtF(lenSq);
        return Vec3(X * invLen, Y * invLen, Z * invLen);
    }
    static float Clamp01(float v) { if (v < 0.0f) return 0.0f; if (v > 1.0f) return 1.0f; return v; } /* :116-127 */
    Vec3 Saturate() const { return Vec3(Clamp01(X), Clamp01(Y), Clamp01(Z)); }
};
static inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.X + b.X, a.Y + b.Y, a.Z + b.Z); }
static inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.X - b.X, a.Y - b.Y, a.Z - b.Z); }
static inline Vec3 operator-(Vec3 a) { return Vec3(-a.X, -a.Y, -a.Z); }
static inline Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a.X * b.X, a.Y * b.Y, a.Z * b.Z); }
static inline Vec3 operator*(Vec3 a, float s) { return Vec3(a.X * s, a.Y * s, a.Z * s); }
static inline Vec3 operator/(Vec3 a, float s) { float inv = 1.0f / s; return Vec3(a.X * inv, a.Y * inv, a.Z * inv); } /* :68-71 */

struct Ray { /* Ray.cs:3-18: the constructor normalises */
    Vec3 Origin, Dir;
    Ray() {}
    Ray(Vec3 o, Vec3 d) : Origin(o), Dir(d.Normalized()) {}
    Vec3 At(float t) const { return Origin + Dir * t; }
};

struct Material { /* Material.cs:5-61, float-rounded scalars (exact for the path, see ycge.h) */
    Vec3 Albedo;
    float Specular = 0, Reflectivity = 0;
    Vec3 Emission;
    float Transparency = 0, IndexOfRefraction = 1.5f;
    Vec3 TransmissionColor = Vec3(1, 1, 1);
    int Tex = -1;
    float TextureWeight = 1, UVScale = 1;
};
static Material FromAbi(const ycge_material &m) {
    Material r;
    r.Albedo = Vec3(m.albedo[0], m.albedo[1], m.albedo[2]);
    r.Specular = m.specular; r.Reflectivity = m.reflectivity;
    r.Emission = Vec3(m.emission[0], m.emission[1], m.emission[2]);
    r.Transparency = m.transparency; r.IndexOfRefraction = m.ior;
    r.TransmissionColor = Vec3(m.transmission[0], m.transmission[1], m.transmission[2]);
    r.Tex = m.tex_id; r.TextureWeight = m.tex_weight; r.UVScale = m.uv_scale;
    return r;
}

struct HitRecord { /* HitRecord.cs:3-11 (+ ids, SURVEY 8c "Primitive-ID definition") */
    float T = 0;
    Vec3 P, N;
    Material Mat;
    float U = 0, V = 0;
    int ObjId = -1, SubId = -1;
};

struct Counters {
    uint64_t rays = 0, top_nodes = 0, mesh_nodes = 0, leaf_refs = 0, tris = 0, prims = 0, dda = 0;
    void add(const Counters &o) {
        rays += o.rays; top_nodes += o.top_nodes; mesh_nodes += o.mesh_nodes; leaf_refs += o.leaf_refs;
        tris += o.tris; prims += o.prims; dda += o.dda;
    }
};
static thread_local Counters *tl_cnt = nullptr;
#define CNT(field) do { if (tl_cnt) tl_cnt->field++; } while (0)

/* ======================================================================================
 * System.Array.Sort(arr, start, count, comparer) — dotnet/runtime ArraySortHelper<T>
 * (IntrospectiveSort: insertion sort <= 16, median-of-three quicksort, heapsort at depth limit).
 * ====================================================================================== */
template <class T, class Cmp> struct DotnetSort {
    T *k; Cmp cmp;
    void SwapIfGreater(int i, int j) { if (cmp(k[i], k[j]) > 0) std::swap(k[i], k[j]); }
    void InsertionSort(int lo, int n) {
        for (int i = 0; i < n - 1; i++) {
            T t = k[lo + i + 1];
            int j = i;
            while (j >= 0 && cmp(t, k[lo + j]) < 0) { k[lo + j + 1] = k[lo + j]; j--; }
            k[lo + j + 1] = t;
        }
    }
    void DownHeap(int lo, int i, int n) {
        T d = k[lo + i - 1];
        while (i <= (n >> 1)) {
            int child = 2 * i;
            if (child < n && cmp(k[lo + child - 1], k[lo + child]) < 0) child++;
            if (!(cmp(d, k[lo + child - 1]) < 0)) break;
            k[lo + i - 1] = k[lo + child - 1];
            i = child;
        }
        k[lo + i - 1] = d;
    }
    void HeapSort(int lo, int n) {
        for (int i = n >> 1; i >= 1; i--) DownHeap(lo, i, n);
        for (int i = n; i > 1; i--) { std::swap(k[lo], k[lo + i - 1]); DownHeap(lo, 1, i - 1); }
    }
    int PickPivotAndPartition(int lo, int n) {
        T *s = k + lo;
        int hi = n - 1, middle = hi >> 1;
        if (cmp(s[0], s[middle]) > 0) std::swap(s[0], s[middle]);
        if (cmp(s[0], s[hi]) > 0) std::swap(s[0], s[hi]);
        if (cmp(s[middle], s[hi]) > 0) std::swap(s[middle], s[hi]);
        T pivot = s[middle];
        std::swap(s[middle], s[hi - 1]);
        int left = 0, right = hi - 1;
        while (left < right) {
            while (cmp(s[++left], pivot) < 0) {}
            while (cmp(pivot, s[--right]) < 0) {}
            if (left >= right) break;
            std::swap(s[left], s[right]);
        }
        if (left != hi - 1) std::swap(s[left], s[hi - 1]);
        return left;
    }
    void IntroSort(int lo, int n, int depthLimit) {
        int partitionSize = n;
        while (partitionSize > 1) {
            if (partitionSize <= 16) {
                if (partitionSize == 2) { SwapIfGreater(lo, lo + 1); return; }
                if (partitionSize == 3) { SwapIfGreater(lo, lo + 1); SwapIfGreater(lo, lo + 2); SwapIfGreater(lo + 1, lo + 2); return; }
                InsertionSort(lo, partitionSize);
                return;
            }
            if (depthLimit == 0) { HeapSort(lo, partitionSize); return; }
            depthLimit--;
            int p = PickPivotAndPartition(lo, partitionSize);
            IntroSort(lo + p + 1, partitionSize - (p + 1), depthLimit);
            partitionSize = p;
        }
    }
    void Sort(int start, int count) {
        if (count > 1) {
            int log2 = 31 - __builtin_clz((unsigned)count);
            IntroSort(start, count, 2 * (log2 + 1));
        }
    }
};
static inline int FloatCompareTo(float a, float b) { /* System.Single.CompareTo */
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    if (a != a) return (b != b) ? 0 : -1;
    return 1;
}

/* ======================================================================================
 * Flat BVH storage + the two binned-SAH builders (Objects/BVH.cs:258-459, MeshBVH.cs:371-576)
 * ====================================================================================== */
struct FlatBVH {
    std::vector<float> minX, minY, minZ, maxX, maxY, maxZ;
    std::vector<int> left, right, start, count, leafIndex;
    int root = -1;
    int nodes() const { return (int)minX.size(); }
};
struct Item { int Index; float MinX, MinY, MinZ, MaxX, MaxY, MaxZ, Cx, Cy, Cz; };
struct NodeTmp { float MinX, MinY, MinZ, MaxX, MaxY, MaxZ; int Left, Right, Start, Count; };

static inline void Surround(float &minX, float &minY, float &minZ, float &maxX, float &maxY, float &maxZ,
                            float oMinX, float oMinY, float oMinZ, float oMaxX, float oMaxY, float oMaxZ) {
    if (oMinX < minX) minX = oMinX; if (oMinY < minY) minY = oMinY; if (oMinZ < minZ) minZ = oMinZ;
    if (oMaxX > maxX) maxX = oMaxX; if (oMaxY > maxY) maxY = oMaxY; if (oMaxZ > maxZ) maxZ = oMaxZ;
}
static inline float SurfaceArea(float minX, float minY, float minZ, float maxX, float maxY, float maxZ) {
    float dx = maxX - minX, dy = maxY - minY, dz = maxZ - minZ;
    return 2.0f * (dx * dy + dx * dz + dy * dz);
}
static inline float AxisC(const Item &it, int ax) { return ax == 0 ? it.Cx : ax == 1 ? it.Cy : it.Cz; }

struct Builder {
    int TargetLeafSize; /* 4 (BVH.cs:7) or 8 (MeshBVH.cs:14) */
    bool meshVariant;   /* partition pass: BVH.cs:394-396 re-derives origin/extent from arr[start]/arr[last]; MeshBVH.cs:511-513 reuses cmin/ext */
    std::vector<NodeTmp> nodes;
    std::vector<int> leafIndices;
    uint64_t sortFallbacks = 0;
    static const int SAH_Bins = 16;

    void SortRange(Item *arr, int start, int count, int axis) {
        sortFallbacks++;
        auto cmp = [axis](const Item &a, const Item &b) { return FloatCompareTo(AxisC(a, axis), AxisC(b, axis)); };
        if (g_sort_mode == 0) {
            DotnetSort<Item, decltype(cmp)> s{arr, cmp};
            s.Sort(start, count);
        } else {
            std::stable_sort(arr + start, arr + start + count, [&](const Item &a, const Item &b) { return cmp(a, b) < 0; });
        }
    }

    int BuildRecursive(Item *arr, int start, int count) {
        if (count <= 0) return -1;
        if (count <= TargetLeafSize) {
            NodeTmp leaf{};
            float mnx = arr[start].MinX, mny = arr[start].MinY, mnz = arr[start].MinZ;
            float mxx = arr[start].MaxX, mxy = arr[start].MaxY, mxz = arr[start].MaxZ;
            for (int i = 1; i < count; i++)
                Surround(mnx, mny, mnz, mxx, mxy, mxz, arr[start + i].MinX, arr[start + i].MinY, arr[start + i].MinZ,
                         arr[start + i].MaxX, arr[start + i].MaxY, arr[start + i].MaxZ);
            int baseIndex = (int)leafIndices.size();
            for (int i = 0; i < count; i++) leafIndices.push_back(arr[start + i].Index);
            leaf.MinX = mnx; leaf.MinY = mny; leaf.MinZ = mnz; leaf.MaxX = mxx; leaf.MaxY = mxy; leaf.MaxZ = mxz;
            leaf.Left = -1; leaf.Right = -1; leaf.Start = baseIndex; leaf.Count = count;
            int idx = (int)nodes.size();
            nodes.push_back(leaf);
            return idx;
        }

        float cminx = arr[start].Cx, cminy = arr[start].Cy, cminz = arr[start].Cz;
        float cmaxx = cminx, cmaxy = cminy, cmaxz = cminz;
        for (int i = start + 1; i < start + count; i++) {
            float cx = arr[i].Cx, cy = arr[i].Cy, cz = arr[i].Cz;
            if (cx < cminx) cminx = cx; if (cy < cminy) cminy = cy; if (cz < cminz) cminz = cz;
            if (cx > cmaxx) cmaxx = cx; if (cy > cmaxy) cmaxy = cy; if (cz > cmaxz) cmaxz = cz;
        }
        float extX = cmaxx - cminx, extY = cmaxy - cminy, extZ = cmaxz - cminz;
        int axis = 0;
        if (extY > extX && extY >= extZ) axis = 1; else if (extZ > extX && extZ >= extY) axis = 2;

        int splitBin = -1;
        int bestAxis = axis;
        float bestCost = PosInf;

        for (int ax = 0; ax < 3; ax++) {
            float extent = ax == 0 ? extX : ax == 1 ? extY : extZ;
            if (!(extent > 0.0f)) continue;
            float origin = ax == 0 ? cminx : ax == 1 ? cminy : cminz;
            float invExtent = 1.0f / extent;

            int counts[SAH_Bins];
            float lminx[SAH_Bins], lminy[SAH_Bins], lminz[SAH_Bins], lmaxx[SAH_Bins], lmaxy[SAH_Bins], lmaxz[SAH_Bins];
            for (int b = 0; b < SAH_Bins; b++) {
                lminx[b] = lminy[b] = lminz[b] = PosInf;
                lmaxx[b] = lmaxy[b] = lmaxz[b] = NegInf;
                counts[b] = 0;
            }
            for (int i = start; i < start + count; i++) {
                float c = AxisC(arr[i], ax);
                int b = (int)((c - origin) * invExtent * (SAH_Bins - 1));
                if (b < 0) b = 0; if (b >= SAH_Bins) b = SAH_Bins - 1;
                counts[b]++;
                Surround(lminx[b], lminy[b], lminz[b], lmaxx[b], lmaxy[b], lmaxz[b], arr[i].MinX, arr[i].MinY, arr[i].MinZ, arr[i].MaxX, arr[i].MaxY, arr[i].MaxZ);
            }
            int leftCount[SAH_Bins], rightCount[SAH_Bins];
            float leftArea[SAH_Bins], rightArea[SAH_Bins];
            float cLminx = PosInf, cLminy = PosInf, cLminz = PosInf, cLmaxx = NegInf, cLmaxy = NegInf, cLmaxz = NegInf;
            int acc = 0;
            for (int b = 0; b < SAH_Bins; b++) {
                if (counts[b] > 0) Surround(cLminx, cLminy, cLminz, cLmaxx, cLmaxy, cLmaxz, lminx[b], lminy[b], lminz[b], lmaxx[b], lmaxy[b], lmaxz[b]);
                acc += counts[b];
                leftCount[b] = acc;
                leftArea[b] = SurfaceArea(cLminx, cLminy, cLminz, cLmaxx, cLmaxy, cLmaxz);
            }
            float cRminx = PosInf, cRminy = PosInf, cRminz = PosInf, cRmaxx = NegInf, cRmaxy = NegInf, cRmaxz = NegInf;
            acc = 0;
            for (int b = SAH_Bins - 1; b >= 0; b--) {
                if (counts[b] > 0) Surround(cRminx, cRminy, cRminz, cRmaxx, cRmaxy, cRmaxz, lminx[b], lminy[b], lminz[b], lmaxx[b], lmaxy[b], lmaxz[b]);
                acc += counts[b];
                rightCount[b] = acc;
                rightArea[b] = SurfaceArea(cRminx, cRminy, cRminz, cRmaxx, cRmaxy, cRmaxz);
            }
            for (int b = 0; b < SAH_Bins - 1; b++) {
                int lc = leftCount[b];
                int rc = rightCount[b + 1];
                if (lc == 0 || rc == 0) continue;
                float cost = leftArea[b] * lc + rightArea[b + 1] * rc;
                if (cost < bestCost) { bestCost = cost; bestAxis = ax; splitBin = b; }
            }
        }

        int mid;
        if (splitBin < 0) {
            SortRange(arr, start, count, bestAxis);
            mid = start + (count >> 1);
        } else {
            float origin, extent, invExtent;
            if (meshVariant) { /* MeshBVH.cs:511-513 */
                origin = bestAxis == 0 ? cminx : bestAxis == 1 ? cminy : cminz;
                extent = bestAxis == 0 ? extX : bestAxis == 1 ? extY : extZ;
                invExtent = 1.0f / extent;
            } else { /* BVH.cs:394-396 — origin/extent from the first/last (unsorted) items */
                origin = AxisC(arr[start], bestAxis);
                extent = AxisC(arr[start + count - 1], bestAxis) - origin;
                invExtent = extent != 0.0f ? 1.0f / extent : 0.0f;
            }
            int i0 = start, i1 = start + count - 1;
            while (i0 <= i1) {
                float c0 = AxisC(arr[i0], bestAxis);
                int b0;
                if (meshVariant) b0 = (int)((c0 - origin) * invExtent * (SAH_Bins - 1));
                else b0 = invExtent != 0.0f ? (int)((c0 - origin) * invExtent * (SAH_Bins - 1)) : 0;
                if (b0 <= splitBin) i0++;
                else { Item tmp = arr[i0]; arr[i0] = arr[i1]; arr[i1] = tmp; i1--; }
            }
            mid = i0;
            if (mid == start || mid == start + count) {
                SortRange(arr, start, count, bestAxis);
                mid = start + (count >> 1);
            }
        }

        int myIndex = (int)nodes.size();
        nodes.push_back(NodeTmp{});
        int leftIndex = BuildRecursive(arr, start, mid - start);
        int rightIndex = BuildRecursive(arr, mid, start + count - mid);
        NodeTmp cur{};
        cur.Left = leftIndex; cur.Right = rightIndex;
        if (leftIndex >= 0 && rightIndex >= 0) {
            const NodeTmp &L = nodes[leftIndex], &R = nodes[rightIndex];
            cur.MinX = MinF(L.MinX, R.MinX); cur.MinY = MinF(L.MinY, R.MinY); cur.MinZ = MinF(L.MinZ, R.MinZ);
            cur.MaxX = MaxF(L.MaxX, R.MaxX); cur.MaxY = MaxF(L.MaxY, R.MaxY); cur.MaxZ = MaxF(L.MaxZ, R.MaxZ);
        } else if (leftIndex >= 0) {
            const NodeTmp &L = nodes[leftIndex];
            cur.MinX = L.MinX; cur.MinY = L.MinY; cur.MinZ = L.MinZ; cur.MaxX = L.MaxX; cur.MaxY = L.MaxY; cur.MaxZ = L.MaxZ;
        } else {
            const NodeTmp &R = nodes[rightIndex];
            cur.MinX = R.MinX; cur.MinY = R.MinY; cur.MinZ = R.MinZ; cur.MaxX = R.MaxX; cur.MaxY = R.MaxY; cur.MaxZ = R.MaxZ;
        }
        cur.Start = 0; cur.Count = 0;
        nodes[myIndex] = cur;
        return myIndex;
    }

    void Build(std::vector<Item> &items, FlatBVH &out) {
        nodes.clear(); leafIndices.clear();
        nodes.reserve(2 * items.size());
        out = FlatBVH();
        if (items.empty()) { out.root = -1; return; }
        out.root = BuildRecursive(items.data(), 0, (int)items.size());
        int n = (int)nodes.size();
        out.minX.resize(n); out.minY.resize(n); out.minZ.resize(n); out.maxX.resize(n); out.maxY.resize(n); out.maxZ.resize(n);
        out.left.resize(n); out.right.resize(n); out.start.resize(n); out.count.resize(n);
        for (int i = 0; i < n; i++) {
            const NodeTmp &nd = nodes[i];
            out.minX[i] = nd.MinX; out.minY[i] = nd.MinY; out.minZ[i] = nd.MinZ;
            out.maxX[i] = nd.MaxX; out.maxY[i] = nd.MaxY; out.maxZ[i] = nd.MaxZ;
            out.left[i] = nd.Left; out.right[i] = nd.Right; out.start[i] = nd.Start; out.count[i] = nd.Count;
        }
        out.leafIndex = leafIndices;
    }
};

static void BvhFromAbi(const ycge_bvh *b, FlatBVH &out) {
    int n = b->n_nodes;
    out.minX.assign(b->min_x, b->min_x + n); out.minY.assign(b->min_y, b->min_y + n); out.minZ.assign(b->min_z, b->min_z + n);
    out.maxX.assign(b->max_x, b->max_x + n); out.maxY.assign(b->max_y, b->max_y + n); out.maxZ.assign(b->max_z, b->max_z + n);
    out.left.assign(b->left, b->left + n); out.right.assign(b->right, b->right + n);
    out.start.assign(b->start, b->start + n); out.count.assign(b->count, b->count + n);
    out.leafIndex.assign(b->leaf_index, b->leaf_index + b->n_leaf_refs);
    out.root = b->root;
}

/* ======================================================================================
 * MeshBVH (Objects/MeshBVH.cs)
 * ====================================================================================== */
struct MeshBVH {
    FlatBVH bvh;
    std::vector<float> ax, ay, az, e1x, e1y, e1z, e2x, e2y, e2z, nx, ny, nz;
    Material mat;
    uint64_t sortFallbacks = 0;

    /* new MeshBVH(tris)  MeshBVH.cs:41-130; abc = A,B,C per triangle */
    void FromTriangles(int n, const float *abc, const Material &m) {
        mat = m;
        std::vector<Item> items(n);
        ax.resize(n); ay.resize(n); az.resize(n); e1x.resize(n); e1y.resize(n); e1z.resize(n);
        e2x.resize(n); e2y.resize(n); e2z.resize(n); nx.resize(n); ny.resize(n); nz.resize(n);
        for (int i = 0; i < n; i++) {
            const float *t = abc + 9 * (size_t)i;
            float Ax = t[0], Ay = t[1], Az = t[2], Bx = t[3], By = t[4], Bz = t[5], Cx = t[6], Cy = t[7], Cz = t[8];
            const float Eps = 1e-4f; /* TryComputeBounds MeshBVH.cs:351-361 */
            Item it;
            it.MinX = MinF(Ax, MinF(Bx, Cx)) - Eps; it.MinY = MinF(Ay, MinF(By, Cy)) - Eps; it.MinZ = MinF(Az, MinF(Bz, Cz)) - Eps;
            it.MaxX = MaxF(Ax, MaxF(Bx, Cx)) + Eps; it.MaxY = MaxF(Ay, MaxF(By, Cy)) + Eps; it.MaxZ = MaxF(Az, MaxF(Bz, Cz)) + Eps;
            it.Index = i;
            it.Cx = 0.5f * (it.MinX + it.MaxX); it.Cy = 0.5f * (it.MinY + it.MaxY); it.Cz = 0.5f * (it.MinZ + it.MaxZ);
            items[i] = it;
            ax[i] = Ax; ay[i] = Ay; az[i] = Az;
            float lx = Bx - Ax, ly = By - Ay, lz = Bz - Az;
            float mx = Cx - Ax, my = Cy - Ay, mz = Cz - Az;
            e1x[i] = lx; e1y[i] = ly; e1z[i] = lz; e2x[i] = mx; e2y[i] = my; e2z[i] = mz;
            float nnx = ly * mz - lz * my, nny = lz * mx - lx * mz, nnz = lx * my - ly * mx;
            float invLen = 1.0f / MaxF(1e-20f, SqrtF(nnx * nnx + nny * nny + nnz * nnz));
            nx[i] = nnx * invLen; ny[i] = nny * invLen; nz[i] = nnz * invLen;
        }
        Builder b; b.TargetLeafSize = 8; b.meshVariant = true;
        b.Build(items, bvh);
        sortFallbacks = b.sortFallbacks;
    }
    void FromSoa(const ycge_mesh_soa *m) {
        int n = m->n_tris;
        mat = FromAbi(m->material);
        ax.assign(m->ax, m->ax + n); ay.assign(m->ay, m->ay + n); az.assign(m->az, m->az + n);
        e1x.assign(m->e1x, m->e1x + n); e1y.assign(m->e1y, m->e1y + n); e1z.assign(m->e1z, m->e1z + n);
        e2x.assign(m->e2x, m->e2x + n); e2y.assign(m->e2y, m->e2y + n); e2z.assign(m->e2z, m->e2z + n);
        nx.assign(m->nx, m->nx + n); ny.assign(m->ny, m->ny + n); nz.assign(m->nz, m->nz + n);
        BvhFromAbi(m->bvh, bvh);
    }

    /* MeshBVH.BoxHitFast  MeshBVH.cs:308-332 */
    static bool BoxHitFast(float minX, float minY, float minZ, float maxX, float maxY, float maxZ, const Ray &r, float tMin, float tMax,
                           float invDx, float invDy, float invDz, int signX, int signY, int signZ, float &tNear, float &tFar) {
        float ox = r.Origin.X, oy = r.Origin.Y, oz = r.Origin.Z;
        float txEnter = ((signX == 0 ? minX : maxX) - ox) * invDx;
        float txExit = ((signX == 0 ? maxX : minX) - ox) * invDx;
        if (txEnter > tMin) tMin = txEnter;
        if (txExit < tMax) tMax = txExit;
        if (tMax < tMin) { tNear = tMin; tFar = tMax; return false; }
        float tyEnter = ((signY == 0 ? minY : maxY) - oy) * invDy;
        float tyExit = ((signY == 0 ? maxY : minY) - oy) * invDy;
        if (tyEnter > tMin) tMin = tyEnter;
        if (tyExit < tMax) tMax = tyExit;
        if (tMax < tMin) { tNear = tMin; tFar = tMax; return false; }
        float tzEnter = ((signZ == 0 ? minZ : maxZ) - oz) * invDz;
        float tzExit = ((signZ == 0 ? maxZ : minZ) - oz) * invDz;
        if (tzEnter > tMin) tMin = tzEnter;
        if (tzExit < tMax) tMax = tzExit;
        tNear = tMin; tFar = tMax;
        return tMax >= tMin;
    }

    /* MeshBVH.TriHit  MeshBVH.cs:239-304 */
    bool TriHit(int i, const Ray &r, float tMin, float tMax, float &t, float &u, float &v) const {
        CNT(tris);
        float dirx = r.Dir.X, diry = r.Dir.Y, dirz = r.Dir.Z;
        float e1x_i = e1x[i], e1y_i = e1y[i], e1z_i = e1z[i];
        float e2x_i = e2x[i], e2y_i = e2y[i], e2z_i = e2z[i];
        float ax_i = ax[i], ay_i = ay[i], az_i = az[i];
        float px = diry * e2z_i - dirz * e2y_i;
        float py = dirz * e2x_i - dirx * e2z_i;
        float pz = dirx * e2y_i - diry * e2x_i;
        float det = e1x_i * px + e1y_i * py + e1z_i * pz;
        const float Eps = 1e-8f;
        if (det > -Eps && det < Eps) { t = 0; u = 0; v = 0; return false; }
        float sx = r.Origin.X - ax_i, sy = r.Origin.Y - ay_i, sz = r.Origin.Z - az_i;
        float uNum = sx * px + sy * py + sz * pz;
        float sgn = det > 0.0f ? 1.0f : -1.0f;
        float detAbs = det * sgn;
        float uNumS = uNum * sgn;
        if (uNumS < 0.0f || uNumS > detAbs) { t = 0; u = 0; v = 0; return false; }
        float qx = sy * e1z_i - sz * e1y_i;
        float qy = sz * e1x_i - sx * e1z_i;
        float qz = sx * e1y_i - sy * e1x_i;
        float vNum = dirx * qx + diry * qy + dirz * qz;
        float vNumS = vNum * sgn;
        float uvSumS = uNumS + vNumS;
        if (vNumS < 0.0f || uvSumS > detAbs) { t = 0; u = 0; v = 0; return false; }
        float tNum = e2x_i * qx + e2y_i * qy + e2z_i * qz;
        float tNumS = tNum * sgn;
        float tMinScaled = tMin * detAbs;
        float tMaxScaled = tMax * detAbs;
        if (tNumS < tMinScaled || tNumS > tMaxScaled) { t = 0; u = 0; v = 0; return false; }
        float invDet = 1.0f / det;
        t = tNum * invDet; u = uNum * invDet; v = vNum * invDet;
        return true;
    }

    /* MeshBVH.Hit  MeshBVH.cs:132-236 */
    bool Hit(const Ray &r, float tMin, float tMax, HitRecord &rec) const {
        if (bvh.root < 0) return false;
        float invDx = 1.0f / r.Dir.X, invDy = 1.0f / r.Dir.Y, invDz = 1.0f / r.Dir.Z;
        int signX = invDx < 0.0f ? 1 : 0, signY = invDy < 0.0f ? 1 : 0, signZ = invDz < 0.0f ? 1 : 0;
        bool hitAnything = false;
        float closest = tMax;
        HitRecord best;
        int stack[64];
        int sp = 0;
        stack[sp++] = bvh.root;
        while (sp > 0) {
            int ni = stack[--sp];
            CNT(mesh_nodes);
            float tNear, tFar;
            if (!BoxHitFast(bvh.minX[ni], bvh.minY[ni], bvh.minZ[ni], bvh.maxX[ni], bvh.maxY[ni], bvh.maxZ[ni], r, tMin, closest,
                            invDx, invDy, invDz, signX, signY, signZ, tNear, tFar))
                continue;
            int cnt = bvh.count[ni];
            if (cnt > 0) {
                int start = bvh.start[ni];
                for (int i = 0; i < cnt; i++) {
                    int tri = bvh.leafIndex[start + i];
                    CNT(leaf_refs);
                    float tHit, u, v;
                    if (TriHit(tri, r, tMin, closest, tHit, u, v)) {
                        closest = tHit;
                        hitAnything = true;
                        best.T = tHit;
                        best.P = Vec3(r.Origin.X + tHit * r.Dir.X, r.Origin.Y + tHit * r.Dir.Y, r.Origin.Z + tHit * r.Dir.Z);
                        float ndotd = nx[tri] * r.Dir.X + ny[tri] * r.Dir.Y + nz[tri] * r.Dir.Z;
                        best.N = ndotd < 0.0f ? Vec3(nx[tri], ny[tri], nz[tri]) : Vec3(-nx[tri], -ny[tri], -nz[tri]);
                        best.Mat = mat;
                        best.U = u; best.V = v;
                        best.SubId = tri;
                    }
                }
            } else {
                int l = bvh.left[ni], rr = bvh.right[ni];
                float lNear = 0, lFar = 0, rNear = 0, rFar = 0;
                bool hitL = false, hitR = false;
                if (l >= 0) hitL = BoxHitFast(bvh.minX[l], bvh.minY[l], bvh.minZ[l], bvh.maxX[l], bvh.maxY[l], bvh.maxZ[l], r, tMin, closest, invDx, invDy, invDz, signX, signY, signZ, lNear, lFar);
                if (rr >= 0) hitR = BoxHitFast(bvh.minX[rr], bvh.minY[rr], bvh.minZ[rr], bvh.maxX[rr], bvh.maxY[rr], bvh.maxZ[rr], r, tMin, closest, invDx, invDy, invDz, signX, signY, signZ, rNear, rFar);
                if (hitL & hitR) {
                    if (lNear < rNear) { stack[sp++] = rr; stack[sp++] = l; }
                    else { stack[sp++] = l; stack[sp++] = rr; }
                } else if (hitL) stack[sp++] = l;
                else if (hitR) stack[sp++] = rr;
            }
        }
        if (hitAnything) rec = best;
        return hitAnything;
    }
    bool TryGetBounds(float b[9]) const { /* MeshBVH.cs:585-603 */
        if (bvh.root < 0) return false;
        int r = bvh.root;
        b[0] = bvh.minX[r]; b[1] = bvh.minY[r]; b[2] = bvh.minZ[r]; b[3] = bvh.maxX[r]; b[4] = bvh.maxY[r]; b[5] = bvh.maxZ[r];
        b[6] = 0.5f * (b[0] + b[3]); b[7] = 0.5f * (b[1] + b[4]); b[8] = 0.5f * (b[2] + b[5]);
        return true;
    }
};

/* ======================================================================================
 * VolumeGrid (Objects/VolumeGrid.cs)
 * ====================================================================================== */
struct VolumeGrid {
    int nx = 0, ny = 0, nz = 0, nbx = 0, nby = 0, nbz = 0;
    std::vector<int> mat, meta;
    Vec3 minCorner, voxelSize;
    bool wireframe = true;
    float wireWidthFrac = 0.06f, wireMaxDistance = 16.0f;
    int palN = 0, palLevels = 1, palDefault = 0;
    std::vector<int> palette;
    /* cached "center block" (VolumeGrid.cs:47-50) — racy in the reference; serial-order semantics here */
    mutable int centerIx = INT32_MIN, centerIy = INT32_MIN, centerIz = INT32_MIN;
    mutable bool centerValid = false;

    static int Morton3_3bits(int x, int y, int z) { /* VolumeGrid.cs:246-252 */
        return ((x & 1) << 0) | ((y & 1) << 1) | ((z & 1) << 2) | ((x & 2) << 2) | ((y & 2) << 3) | ((z & 2) << 4) |
               ((x & 4) << 4) | ((y & 4) << 5) | ((z & 4) << 6);
    }
    int IndexOf(int ix, int iy, int iz) const { /* VolumeGrid.cs:235-242 */
        int bx = ix >> 3, by = iy >> 3, bz = iz >> 3;
        int lx = ix & 7, ly = iy & 7, lz = iz & 7;
        int brickLinear = ((bz * nby) + by) * nbx + bx;
        return brickLinear * 512 + Morton3_3bits(lx, ly, lz);
    }
    void FromAbi(const ycge_volume *v) {
        nx = v->nx; ny = v->ny; nz = v->nz;
        nbx = (nx + 7) >> 3; nby = (ny + 7) >> 3; nbz = (nz + 7) >> 3;
        size_t cap = (size_t)nbx * nby * nbz * 512;
        mat.assign(v->mat, v->mat + cap); meta.assign(v->meta, v->meta + cap);
        minCorner = Vec3(v->min_corner[0], v->min_corner[1], v->min_corner[2]);
        voxelSize = Vec3(v->voxel_size[0], v->voxel_size[1], v->voxel_size[2]);
        wireframe = v->wireframe != 0; wireWidthFrac = v->wire_width_frac; wireMaxDistance = v->wire_max_distance;
        palN = v->palette_n_ids; palLevels = v->palette_meta_levels < 1 ? 1 : v->palette_meta_levels; palDefault = v->palette_default;
        palette.assign(v->palette, v->palette + (size_t)palN * palLevels);
    }
    int LookupMaterialIndex(int id, int metaId) const {
        if (id < 0 || id >= palN) return palDefault;
        int m = metaId < 0 ? 0 : (metaId >= palLevels ? palLevels - 1 : metaId);
        return palette[(size_t)id * palLevels + m];
    }
    static bool Slab(float ro, float rd, float mn, float mx, float &tEnter, float &tExit, int axis, int &enterAxis) { /* :334-355 */
        const float eps = 1e-12f;
        if (AbsF(rd) < eps) { if (ro < mn || ro > mx) return false; return true; }
        float inv = 1.0f / rd;
        float t0 = (mn - ro) * inv, t1 = (mx - ro) * inv;
        if (t0 > t1) { float tmp = t0; t0 = t1; t1 = tmp; }
        if (t0 > tEnter) { tEnter = t0; enterAxis = axis; }
        if (t1 < tExit) tExit = t1;
        return tExit >= tEnter;
    }
    static bool RayAabb(const Ray &r, Vec3 bmin, Vec3 bmax, float &tEnter, float &tExit, int &enterAxis) { /* :319-331 */
        tEnter = NegInf; tExit = PosInf; enterAxis = -1;
        if (!Slab(r.Origin.X, r.Dir.X, bmin.X, bmax.X, tEnter, tExit, 0, enterAxis)) return false;
        if (!Slab(r.Origin.Y, r.Dir.Y, bmin.Y, bmax.Y, tEnter, tExit, 1, enterAxis)) return false;
        if (!Slab(r.Origin.Z, r.Dir.Z, bmin.Z, bmax.Z, tEnter, tExit, 2, enterAxis)) return false;
        return tExit >= MaxF(0.0f, tEnter);
    }
    static double EdgeDistance(double v, double v0, double v1) { /* :291-297 */
        double a = v - v0, b = v1 - v;
        if (a < 0.0) a = 0.0; if (b < 0.0) b = 0.0;
        return std::min(a, b); /* Math.Min(double): no NaN/-0 can reach here */
    }
    bool IsWireOnFace(Vec3 p, int ix, int iy, int iz, int axis) const { /* :256-283; Vec3 fields are float: float math, then widened */
        double x0 = minCorner.X + ix * voxelSize.X; double x1 = x0 + voxelSize.X;
        double y0 = minCorner.Y + iy * voxelSize.Y; double y1 = y0 + voxelSize.Y;
        double z0 = minCorner.Z + iz * voxelSize.Z; double z1 = z0 + voxelSize.Z;
        if (axis == 0) {
            double dy = EdgeDistance((double)p.Y, y0, y1), dz = EdgeDistance((double)p.Z, z0, z1);
            double w = wireWidthFrac * MinF(voxelSize.Y, voxelSize.Z);
            return dy <= w || dz <= w;
        } else if (axis == 1) {
            double dx = EdgeDistance((double)p.X, x0, x1), dz = EdgeDistance((double)p.Z, z0, z1);
            double w = wireWidthFrac * MinF(voxelSize.X, voxelSize.Z);
            return dx <= w || dz <= w;
        } else {
            double dx = EdgeDistance((double)p.X, x0, x1), dy = EdgeDistance((double)p.Y, y0, y1);
            double w = wireWidthFrac * MinF(voxelSize.X, voxelSize.Y);
            return dx <= w || dy <= w;
        }
    }
    static Vec3 FaceNormalFromAxis(int axis, int stepX, int stepY, int stepZ) { /* :302-308 */
        if (axis == 0) return Vec3(stepX > 0 ? -1.0f : 1.0f, 0, 0);
        if (axis == 1) return Vec3(0, stepY > 0 ? -1.0f : 1.0f, 0);
        if (axis == 2) return Vec3(0, 0, stepZ > 0 ? -1.0f : 1.0f);
        return Vec3(0, 0, 0);
    }

    /* VolumeGrid.Hit  VolumeGrid.cs:99-231.  Returns the material *index*; albedoOverride: 0 none, 1 black wire, 2 white wire */
    bool Hit(const Ray &r, float tMin, float tMax, HitRecord &rec, float screenU, float screenV, int &matIndex, int &albedoOverride) const {
        float minX = minCorner.X, minY = minCorner.Y, minZ = minCorner.Z;
        float sizeX = voxelSize.X, sizeY = voxelSize.Y, sizeZ = voxelSize.Z;
        float maxX = minX + nx * sizeX, maxY = minY + ny * sizeY, maxZ = minZ + nz * sizeZ;
        int enterAxis;
        float tEnter, tExit;
        if (!RayAabb(r, Vec3(minX, minY, minZ), Vec3(maxX, maxY, maxZ), tEnter, tExit, enterAxis)) return false;
        float t = tEnter; if (t < tMin) t = tMin; if (t > tMax || t > tExit) return false;
        const float eps = 1e-6f;
        t += eps;
        float ox = r.Origin.X, oy = r.Origin.Y, oz = r.Origin.Z;
        float dx = r.Dir.X, dy = r.Dir.Y, dz = r.Dir.Z;
        float px = ox + dx * t, py = oy + dy * t, pz = oz + dz * t;
        int ix = (int)FloorF((px - minX) / sizeX); if (ix < 0) ix = 0; else if (ix >= nx) ix = nx - 1;
        int iy = (int)FloorF((py - minY) / sizeY); if (iy < 0) iy = 0; else if (iy >= ny) iy = ny - 1;
        int iz = (int)FloorF((pz - minZ) / sizeZ); if (iz < 0) iz = 0; else if (iz >= nz) iz = nz - 1;
        int stepX = dx > 0.0f ? 1 : dx < 0.0f ? -1 : 0;
        int stepY = dy > 0.0f ? 1 : dy < 0.0f ? -1 : 0;
        int stepZ = dz > 0.0f ? 1 : dz < 0.0f ? -1 : 0;
        float invDx = stepX == 0 ? 0.0f : 1.0f / dx;
        float invDy = stepY == 0 ? 0.0f : 1.0f / dy;
        float invDz = stepZ == 0 ? 0.0f : 1.0f / dz;
        float nextVx = minX + (stepX > 0 ? (ix + 1) * sizeX : ix * sizeX);
        float nextVy = minY + (stepY > 0 ? (iy + 1) * sizeY : iy * sizeY);
        float nextVz = minZ + (stepZ > 0 ? (iz + 1) * sizeZ : iz * sizeZ);
        float tMaxX = stepX == 0 ? PosInf : (nextVx - ox) * invDx;
        float tMaxY = stepY == 0 ? PosInf : (nextVy - oy) * invDy;
        float tMaxZ = stepZ == 0 ? PosInf : (nextVz - oz) * invDz;
        float tDeltaX = stepX == 0 ? PosInf : AbsF(sizeX * invDx);
        float tDeltaY = stepY == 0 ? PosInf : AbsF(sizeY * invDy);
        float tDeltaZ = stepZ == 0 ? PosInf : AbsF(sizeZ * invDz);
        int lastAxis = enterAxis < 0 ? (tMaxX <= tMaxY && tMaxX <= tMaxZ ? 0 : tMaxY <= tMaxZ ? 1 : 2) : enterAxis;
        bool wf = wireframe;
        float wireMax2 = wireMaxDistance <= 0.0f ? -1.0f : wireMaxDistance * wireMaxDistance;
        float dirLen2 = dx * dx + dy * dy + dz * dz;

        while (t <= tExit && t <= tMax) {
            if ((unsigned)ix < (unsigned)nx && (unsigned)iy < (unsigned)ny && (unsigned)iz < (unsigned)nz) {
                CNT(dda);
                int idx = IndexOf(ix, iy, iz);
                int matId = mat[idx];
                if (matId > 0) {
                    int metaId = meta[idx];
                    int normalAxis = lastAxis;
                    float hitT = MaxF(t, tMin);
                    if (normalAxis < 0) {
                        if (tMaxX <= tMaxY && tMaxX <= tMaxZ) { normalAxis = 0; hitT = MaxF(tMaxX, tMin); }
                        else if (tMaxY <= tMaxZ) { normalAxis = 1; hitT = MaxF(tMaxY, tMin); }
                        else { normalAxis = 2; hitT = MaxF(tMaxZ, tMin); }
                    }
                    Vec3 n = FaceNormalFromAxis(normalAxis, stepX, stepY, stepZ);
                    Vec3 hitPoint = r.At(hitT);
                    bool withinWireRange = false;
                    if (wf && wireMax2 >= 0.0f) {
                        float dist2 = hitT * hitT * dirLen2;
                        withinWireRange = dist2 <= wireMax2;
                    }
                    bool isCenterBlock = false;
                    if (wf) {
                        bool isCenterRay = AbsF(screenU - 0.5f) <= 0.000001f && AbsF(screenV - 0.5f) <= 0.000001f;
                        if (isCenterRay) { centerIx = ix; centerIy = iy; centerIz = iz; centerValid = true; }
                        isCenterBlock = centerValid && ix == centerIx && iy == centerIy && iz == centerIz;
                    }
                    matIndex = LookupMaterialIndex(matId, metaId);
                    albedoOverride = 0;
                    if (wf && withinWireRange && IsWireOnFace(hitPoint, ix, iy, iz, normalAxis)) albedoOverride = isCenterBlock ? 2 : 1;
                    rec.T = hitT; rec.P = hitPoint; rec.N = n; rec.U = 0; rec.V = 0;
                    rec.SubId = ix + nx * (iy + ny * iz);
                    return true;
                }
            }
            if (tMaxX <= tMaxY && tMaxX <= tMaxZ) { ix += stepX; t = tMaxX; tMaxX += tDeltaX; lastAxis = 0; }
            else if (tMaxY <= tMaxZ) { iy += stepY; t = tMaxY; tMaxY += tDeltaY; lastAxis = 1; }
            else { iz += stepZ; t = tMaxZ; tMaxZ += tDeltaZ; lastAxis = 2; }
            if ((unsigned)ix >= (unsigned)nx || (unsigned)iy >= (unsigned)ny || (unsigned)iz >= (unsigned)nz) break;
        }
        return false;
    }
    bool TryGetBounds(float b[9]) const { /* VolumeGrid.cs:386-403 */
        if (nx <= 0 || ny <= 0 || nz <= 0) return false;
        b[0] = minCorner.X; b[1] = minCorner.Y; b[2] = minCorner.Z;
        b[3] = minCorner.X + nx * voxelSize.X; b[4] = minCorner.Y + ny * voxelSize.Y; b[5] = minCorner.Z + nz * voxelSize.Z;
        b[6] = 0.5f * (b[0] + b[3]); b[7] = 0.5f * (b[1] + b[4]); b[8] = 0.5f * (b[2] + b[5]);
        return true;
    }
};

/* ======================================================================================
 * Scene objects (Objects/BoundedObjects.cs, Surfaces.cs, Triangle.cs) and the top-level BVH
 * ====================================================================================== */
/* Renderer/Texture.cs (static images): int[] pixels = RGBA bytes (byte 0 = R, InitializeFromRgbaMat :81-90), row-major,
   row 0 first; SampleBilinear :143-162 */
struct Texture {
    int width = 0, height = 0;
    std::vector<uint32_t> pixels;
    static Vec3 Lerp(Vec3 a, Vec3 b, float t) { return a * (1.0f - t) + b * t; } /* :164-167 */
    Vec3 Texel(int idx) const { /* new RGBA32(int).toVec3()  RGBA32.cs:14-21,:82-85 */
        uint32_t v = pixels[(size_t)idx];
        return Vec3((float)(v & 255u) / 255.0f, (float)((v >> 8) & 255u) / 255.0f, (float)((v >> 16) & 255u) / 255.0f);
    }
    Vec3 SampleBilinear(float u, float v) const {
        if (width <= 0 || height <= 0 || pixels.empty()) return Vec3(1.0f, 1.0f, 1.0f);
        u = u - FloorF(u);
        v = v - FloorF(v);
        float fx = u * (float)(width - 1);
        float fy = v * (float)(height - 1);
        int x0 = (int)FloorF(fx);
        int y0 = (int)FloorF(fy);
        int x1 = (x0 + 1) % width;
        int y1 = (y0 + 1) % height;
        float tx = fx - (float)x0;
        float ty = fy - (float)y0;
        Vec3 c00 = Texel(y0 * width + x0), c10 = Texel(y0 * width + x1), c01 = Texel(y1 * width + x0), c11 = Texel(y1 * width + x1);
        Vec3 a = Lerp(c00, c10, tx);
        Vec3 b = Lerp(c01, c11, tx);
        Vec3 c = Lerp(a, b, ty);
        return c.Saturate();
    }
};

struct Scene;
struct Object {
    int kind = 0, matA = 0, matB = 0, overrideSR = 0, refId = -1;
    float checkerScale = 0, specular = 0, reflectivity = 0;
    float p[12] = {0};
    /* derived in the reference's constructors */
    float ndot = 0, radius2 = 0, invSpanA = 0, invSpanB = 0;
    float e1x = 0, e1y = 0, e1z = 0, e2x = 0, e2y = 0, e2z = 0, tnx = 0, tny = 0, tnz = 0;
};

struct Scene {
    Vec3 bgTop, bgBottom, ambientColor;
    float ambientIntensity = 0;
    bool isVolumeScene = false;
    std::vector<ycge_light> lights;
    std::vector<Material> materials;
    std::vector<Object> objects;
    FlatBVH bvh;
    uint64_t sortFallbacks = 0;
    std::map<int, std::shared_ptr<MeshBVH>> *meshes = nullptr;
    std::map<int, std::shared_ptr<VolumeGrid>> *volumes = nullptr;
    std::map<int, std::shared_ptr<Texture>> *textures = nullptr;
    std::vector<const MeshBVH *> objMesh;
    std::vector<const VolumeGrid *> objVol;

    Material MatFunc(const Object &o, Vec3 pos) const { /* Scenes/Scenes.cs:408-428 + override Surfaces.cs:64-66 */
        Material m;
        if (o.checkerScale != 0.0f) {
            int cx = (int)FloorF(pos.X / o.checkerScale);
            int cz = (int)FloorF(pos.Z / o.checkerScale);
            bool check = ((cx + cz) & 1) == 0;
            m = materials[check ? o.matA : o.matB];
        } else m = materials[o.matA];
        if (o.overrideSR) { m.Specular = o.specular; m.Reflectivity = o.reflectivity; }
        return m;
    }

    bool TryGetBounds(int i, float b[9]) const {
        const Object &o = objects[i];
        const float *p = o.p;
        switch (o.kind) {
            case YCGE_SPHERE: case YCGE_DISK: { /* BoundedObjects.cs:20-29, Surfaces.cs:96-105 */
                float R = o.kind == YCGE_SPHERE ? p[3] : p[6];
                b[0] = p[0] - R; b[1] = p[1] - R; b[2] = p[2] - R; b[3] = p[0] + R; b[4] = p[1] + R; b[5] = p[2] + R;
                break; }
            case YCGE_PLANE: { float B = 1e6f; b[0] = b[1] = b[2] = -B; b[3] = b[4] = b[5] = B; b[6] = b[7] = b[8] = 0.0f; return true; } /* Surfaces.cs:30-36 */
            case YCGE_XYRECT: { const float E = 1e-4f; b[0] = p[0]; b[1] = p[2]; b[2] = p[4] - E; b[3] = p[1]; b[4] = p[3]; b[5] = p[4] + E; break; }
            case YCGE_XZRECT: { const float E = 1e-4f; b[0] = p[0]; b[1] = p[4] - E; b[2] = p[2]; b[3] = p[1]; b[4] = p[4] + E; b[5] = p[3]; break; }
            case YCGE_YZRECT: { const float E = 1e-4f; b[0] = p[4] - E; b[1] = p[0]; b[2] = p[2]; b[3] = p[4] + E; b[4] = p[1]; b[5] = p[3]; break; }
            case YCGE_BOX: b[0] = p[0]; b[1] = p[1]; b[2] = p[2]; b[3] = p[3]; b[4] = p[4]; b[5] = p[5]; break;
            case YCGE_CYLINDER_Y: b[0] = p[0] - p[3]; b[1] = p[4]; b[2] = p[2] - p[3]; b[3] = p[0] + p[3]; b[4] = p[5]; b[5] = p[2] + p[3]; break;
            case YCGE_TRIANGLE: { /* Triangle.cs:54-66 */
                const float E = 1e-4f;
                b[0] = MinF(p[0], MinF(p[3], p[6])) - E; b[1] = MinF(p[1], MinF(p[4], p[7])) - E; b[2] = MinF(p[2], MinF(p[5], p[8])) - E;
                b[3] = MaxF(p[0], MaxF(p[3], p[6])) + E; b[4] = MaxF(p[1], MaxF(p[4], p[7])) + E; b[5] = MaxF(p[2], MaxF(p[5], p[8])) + E;
                break; }
            case YCGE_MESH: return objMesh[i] && objMesh[i]->TryGetBounds(b);
            case YCGE_VOLUME: return objVol[i] && objVol[i]->TryGetBounds(b);
            default: return false;
        }
        b[6] = 0.5f * (b[0] + b[3]); b[7] = 0.5f * (b[1] + b[4]); b[8] = 0.5f * (b[2] + b[5]);
        return true;
    }

    /* XYRect/XZRect/YZRect.Hit  Surfaces.cs:184-214, 256-286, 328-358. axis: 2 = XY (normal Z), 1 = XZ, 0 = YZ */
    bool RectHit(const Object &o, int axis, const float *q, const Ray &r, float tMin, float tMax, HitRecord &rec) const {
        float a0 = q[0], a1 = q[1], b0 = q[2], b1 = q[3], k = q[4];
        float invA = 1.0f / (a1 - a0), invB = 1.0f / (b1 - b0);
        float dirK = axis == 2 ? r.Dir.Z : axis == 1 ? r.Dir.Y : r.Dir.X;
        float oK = axis == 2 ? r.Origin.Z : axis == 1 ? r.Origin.Y : r.Origin.X;
        float adir = AbsF(dirK);
        float safeDir = CopySignF(MaxF(adir, 1e-8f), dirK);
        float t = (k - oK) / safeDir;
        float pa, pb;
        if (axis == 2) { pa = r.Origin.X + t * r.Dir.X; pb = r.Origin.Y + t * r.Dir.Y; }
        else if (axis == 1) { pa = r.Origin.X + t * r.Dir.X; pb = r.Origin.Z + t * r.Dir.Z; }
        else { pa = r.Origin.Y + t * r.Dir.Y; pb = r.Origin.Z + t * r.Dir.Z; }
        bool ok = adir >= 1e-8f;
        ok &= (t >= tMin) & (t <= tMax);
        ok &= (pa >= a0) & (pa <= a1) & (pb >= b0) & (pb <= b1);
        if (!ok) return false;
        rec.T = t;
        float nk = CopySignF(1.0f, -dirK);
        if (axis == 2) { rec.P = Vec3(pa, pb, k); rec.N = Vec3(0, 0, nk); }
        else if (axis == 1) { rec.P = Vec3(pa, k, pb); rec.N = Vec3(0, nk, 0); }
        else { rec.P = Vec3(k, pa, pb); rec.N = Vec3(nk, 0, 0); }
        rec.Mat = MatFunc(o, rec.P);
        rec.U = (pa - a0) * invA;
        rec.V = (pb - b0) * invB;
        return true;
    }

    bool ObjectHit(int objId, const Ray &r, float tMin, float tMax, HitRecord &rec, float screenU, float screenV) const {
        const Object &o = objects[objId];
        const float *p = o.p;
        CNT(prims);
        switch (o.kind) {
            case YCGE_SPHERE: { /* BoundedObjects.cs:31-69 */
                float Cx = p[0], Cy = p[1], Cz = p[2], Radius = p[3];
                float ox = r.Origin.X - Cx, oy = r.Origin.Y - Cy, oz = r.Origin.Z - Cz;
                float dx = r.Dir.X, dy = r.Dir.Y, dz = r.Dir.Z;
                float a = dx * dx + dy * dy + dz * dz;
                float halfB = ox * dx + oy * dy + oz * dz;
                float c = ox * ox + oy * oy + oz * oz - Radius * Radius;
                float disc = halfB * halfB - a * c;
                if (disc < 0.0f) return false;
                float s = SqrtF(disc);
                float invA = 1.0f / a;
                float t = (-halfB - s) * invA;
                if (t < tMin || t > tMax) {
                    t = (-halfB + s) * invA;
                    if (t < tMin || t > tMax) return false;
                }
                float px = r.Origin.X + t * dx, py = r.Origin.Y + t * dy, pz = r.Origin.Z + t * dz;
                float invR = 1.0f / Radius;
                rec.T = t; rec.P = Vec3(px, py, pz);
                rec.N = Vec3((px - Cx) * invR, (py - Cy) * invR, (pz - Cz) * invR);
                rec.Mat = materials[o.matA]; rec.U = 0; rec.V = 0; rec.SubId = 0;
                return true; }
            case YCGE_PLANE: { /* Surfaces.cs:39-71 */
                float nx = p[3], ny = p[4], nz = p[5];
                float dx = r.Dir.X, dy = r.Dir.Y, dz = r.Dir.Z, ox = r.Origin.X, oy = r.Origin.Y, oz = r.Origin.Z;
                float denom = nx * dx + ny * dy + nz * dz;
                const float Eps = 1e-6f;
                if (denom > -Eps && denom < Eps) return false;
                float t = (o.ndot - (nx * ox + ny * oy + nz * oz)) / denom;
                if (t < tMin || t > tMax) return false;
                float px = ox + t * dx, py = oy + t * dy, pz = oz + t * dz;
                rec.T = t; rec.P = Vec3(px, py, pz);
                rec.N = denom < 0.0f ? Vec3(nx, ny, nz) : Vec3(-nx, -ny, -nz);
                rec.Mat = MatFunc(o, rec.P); rec.U = 0; rec.V = 0; rec.SubId = 0;
                return true; }
            case YCGE_DISK: { /* Surfaces.cs:108-142 */
                Vec3 Center(p[0], p[1], p[2]), Normal(p[3], p[4], p[5]);
                float denom = Normal.Dot(r.Dir);
                float adenom = AbsF(denom);
                float safeDenom = CopySignF(MaxF(adenom, 1e-8f), denom);
                float t = (o.ndot - Normal.Dot(r.Origin)) / safeDenom;
                float px = r.Origin.X + t * r.Dir.X, py = r.Origin.Y + t * r.Dir.Y, pz = r.Origin.Z + t * r.Dir.Z;
                float dx = px - Center.X, dz = pz - Center.Z;
                float rr = dx * dx + dz * dz;
                bool ok = adenom >= 1e-6f;
                ok &= (t >= tMin) & (t <= tMax);
                ok &= rr <= o.radius2;
                if (!ok) return false;
                rec.T = t; rec.P = Vec3(px, py, pz);
                rec.N = denom < 0.0f ? Normal : -Normal;
                rec.Mat = MatFunc(o, rec.P); rec.U = 0; rec.V = 0; rec.SubId = 0;
                return true; }
            case YCGE_XYRECT: rec.SubId = 0; return RectHit(o, 2, p, r, tMin, tMax, rec);
            case YCGE_XZRECT: rec.SubId = 0; return RectHit(o, 1, p, r, tMin, tMax, rec);
            case YCGE_YZRECT: rec.SubId = 0; return RectHit(o, 0, p, r, tMin, tMax, rec);
            case YCGE_BOX: { /* BoundedObjects.cs:78-115: six rects, fixed order, shrinking closest */
                float mnx = p[0], mny = p[1], mnz = p[2], mxx = p[3], mxy = p[4], mxz = p[5];
                float f[6][5] = {{mnx, mxx, mny, mxy, mxz}, {mnx, mxx, mny, mxy, mnz}, {mnx, mxx, mnz, mxz, mxy},
                                 {mnx, mxx, mnz, mxz, mny}, {mny, mxy, mnz, mxz, mxx}, {mny, mxy, mnz, mxz, mnx}};
                static const int axes[6] = {2, 2, 1, 1, 0, 0};
                bool hitAnything = false;
                float closest = tMax;
                HitRecord temp;
                for (int i = 0; i < 6; i++) {
                    if (RectHit(o, axes[i], f[i], r, tMin, closest, temp)) {
                        hitAnything = true; closest = temp.T; rec = temp; rec.SubId = i;
                    }
                }
                return hitAnything; }
            case YCGE_CYLINDER_Y: { /* BoundedObjects.cs:148-247 */
                float Cx = p[0], Cz = p[2], Radius = p[3], YMin = p[4], YMax = p[5];
                bool Capped = p[6] != 0.0f;
                float ox = r.Origin.X - Cx, oy = r.Origin.Y, oz = r.Origin.Z - Cz;
                float dx = r.Dir.X, dy = r.Dir.Y, dz = r.Dir.Z;
                float a = dx * dx + dz * dz;
                float hitT = FloatMax;
                Vec3 hitN;
                bool hit = false;
                if (a > 1e-12f) {
                    float halfB = ox * dx + oz * dz;
                    float c = ox * ox + oz * oz - o.radius2;
                    float disc = halfB * halfB - a * c;
                    if (disc >= 0.0f) {
                        float s = SqrtF(disc);
                        float invA = 1.0f / a;
                        float t1 = (-halfB - s) * invA;
                        if (t1 > tMin && t1 < tMax) {
                            float y1 = oy + t1 * dy;
                            if (y1 >= YMin && y1 <= YMax) {
                                hitT = t1;
                                float nx = (ox + t1 * dx) / Radius, nz = (oz + t1 * dz) / Radius;
                                hitN = Vec3(nx, 0.0f, nz); hit = true;
                            }
                        }
                        if (!hit) {
                            float t2 = (-halfB + s) * invA;
                            if (t2 > tMin && t2 < tMax) {
                                float y2 = oy + t2 * dy;
                                if (y2 >= YMin && y2 <= YMax) {
                                    hitT = t2;
                                    float nx = (ox + t2 * dx) / Radius, nz = (oz + t2 * dz) / Radius;
                                    hitN = Vec3(nx, 0.0f, nz); hit = true;
                                }
                            }
                        }
                    }
                }
                if (Capped && AbsF(dy) > 1e-8f) {
                    float tTop = (YMax - oy) / dy;
                    if (tTop > tMin && tTop < tMax) {
                        float rx = ox + tTop * dx, rz = oz + tTop * dz;
                        if (rx * rx + rz * rz <= o.radius2) if (tTop < hitT) { hitT = tTop; hitN = Vec3(0, 1, 0); hit = true; }
                    }
                    float tBot = (YMin - oy) / dy;
                    if (tBot > tMin && tBot < tMax) {
                        float rx = ox + tBot * dx, rz = oz + tBot * dz;
                        if (rx * rx + rz * rz <= o.radius2) if (tBot < hitT) { hitT = tBot; hitN = Vec3(0, -1, 0); hit = true; }
                    }
                }
                if (!hit) return false;
                float px = r.Origin.X + hitT * dx, py = r.Origin.Y + hitT * dy, pz = r.Origin.Z + hitT * dz;
                rec.T = hitT; rec.P = Vec3(px, py, pz);
                rec.N = hitN.Dot(r.Dir) < 0.0f ? hitN : -hitN;
                rec.Mat = materials[o.matA]; rec.U = 0; rec.V = 0; rec.SubId = 0;
                return true; }
            case YCGE_TRIANGLE: { /* Triangle.cs:69-128, the SSE4.1 path (x86-64): DPPS 0x71 = (x+y)+(z+0) */
                float Dx = r.Dir.X, Dy = r.Dir.Y, Dz = r.Dir.Z;
                float Sx = r.Origin.X - p[0], Sy = r.Origin.Y - p[1], Sz = r.Origin.Z - p[2];
                /* h = D x E2: lanes computed as D*E2yzx - E2*Dyzx, then rotated */
                float hx = Dy * o.e2z - o.e2y * Dz;
                float hy = Dz * o.e2x - o.e2z * Dx;
                float hz = Dx * o.e2y - o.e2x * Dy;
                auto dpps = [](float a0, float a1, float a2, float b0, float b1, float b2) { return (a0 * b0 + a1 * b1) + (a2 * b2 + 0.0f); };
                float det = dpps(o.e1x, o.e1y, o.e1z, hx, hy, hz);
                if (AbsF(det) < 1e-8f) return false;
                float invDet = 1.0f / det;
                float u = dpps(Sx, Sy, Sz, hx, hy, hz) * invDet;
                if (u < 0.0f || u > 1.0f) return false;
                float qx = Sy * o.e1z - o.e1y * Sz;
                float qy = Sz * o.e1x - o.e1z * Sx;
                float qz = Sx * o.e1y - o.e1x * Sy;
                float v = dpps(Dx, Dy, Dz, qx, qy, qz) * invDet;
                if (v < 0.0f || (u + v) > 1.0f) return false;
                float t = dpps(o.e2x, o.e2y, o.e2z, qx, qy, qz) * invDet;
                if (t < tMin || t > tMax) return false;
                rec.T = t;
                rec.P = Vec3(r.Origin.X + t * r.Dir.X, r.Origin.Y + t * r.Dir.Y, r.Origin.Z + t * r.Dir.Z);
                float ndotd = o.tnx * r.Dir.X + o.tny * r.Dir.Y + o.tnz * r.Dir.Z;
                rec.N = ndotd < 0.0f ? Vec3(o.tnx, o.tny, o.tnz) : Vec3(-o.tnx, -o.tny, -o.tnz);
                rec.Mat = materials[o.matA]; rec.U = u; rec.V = v; rec.SubId = 0;
                return true; }
            case YCGE_MESH: return objMesh[objId]->Hit(r, tMin, tMax, rec); /* Mesh.cs:26-29 */
            case YCGE_VOLUME: {
                int mi = 0, ov = 0;
                if (!objVol[objId]->Hit(r, tMin, tMax, rec, screenU, screenV, mi, ov)) return false;
                rec.Mat = materials[mi];
                if (ov == 1) rec.Mat.Albedo = Vec3(0, 0, 0); else if (ov == 2) rec.Mat.Albedo = Vec3(1, 1, 1);
                return true; }
        }
        return false;
    }

    /* BVH.BoxHitFast  Objects/BVH.cs:201-236 */
    static bool BoxHitFast(float minX, float minY, float minZ, float maxX, float maxY, float maxZ, const Ray &r, float tMin, float tMax,
                           float invDx, float invDy, float invDz, float &tNear, float &tFar) {
        float ox = r.Origin.X, oy = r.Origin.Y, oz = r.Origin.Z;
        float tEnterX = (minX - ox) * invDx, tExitX = (maxX - ox) * invDx;
        if (tEnterX > tExitX) { float tmp = tEnterX; tEnterX = tExitX; tExitX = tmp; }
        float tEnterY = (minY - oy) * invDy, tExitY = (maxY - oy) * invDy;
        if (tEnterY > tExitY) { float tmp = tEnterY; tEnterY = tExitY; tExitY = tmp; }
        float tEnterZ = (minZ - oz) * invDz, tExitZ = (maxZ - oz) * invDz;
        if (tEnterZ > tExitZ) { float tmp = tEnterZ; tEnterZ = tExitZ; tExitZ = tmp; }
        float tEnter = MaxF(tEnterX, MaxF(tEnterY, tEnterZ));
        float tExit = MinF(tExitX, MinF(tExitY, tExitZ));
        if (tEnter < tMin) tEnter = tMin;
        if (tExit > tMax) tExit = tMax;
        tNear = tEnter; tFar = tExit;
        return tExit >= tEnter;
    }

    /* Scene.Hit -> BVH.Hit  Scenes/Scene.cs:71-75, Objects/BVH.cs:99-198 */
    bool Hit(const Ray &r, float tMin, float tMax, HitRecord &rec, float screenU, float screenV) const {
        CNT(rays);
        if (bvh.root < 0) return false;
        float invDx = 1.0f / r.Dir.X, invDy = 1.0f / r.Dir.Y, invDz = 1.0f / r.Dir.Z;
        bool hitAnything = false;
        float closest = tMax;
        HitRecord best;
        int stack[128];
        int sp = 0;
        stack[sp++] = bvh.root;
        while (sp > 0) {
            int ni = stack[--sp];
            CNT(top_nodes);
            float tNear, tFar;
            if (!BoxHitFast(bvh.minX[ni], bvh.minY[ni], bvh.minZ[ni], bvh.maxX[ni], bvh.maxY[ni], bvh.maxZ[ni], r, tMin, closest, invDx, invDy, invDz, tNear, tFar))
                continue;
            int cnt = bvh.count[ni];
            if (cnt > 0) {
                int start = bvh.start[ni];
                for (int i = 0; i < cnt; i++) {
                    int objId = bvh.leafIndex[start + i];
                    CNT(leaf_refs);
                    HitRecord tmp;
                    if (ObjectHit(objId, r, tMin, closest, tmp, screenU, screenV)) {
                        hitAnything = true; closest = tmp.T; best = tmp; best.ObjId = objId;
                    }
                }
            } else {
                int l = bvh.left[ni], rr = bvh.right[ni];
                float lNear = 0, lFar = 0, rNear = 0, rFar = 0;
                bool hitL = false, hitR = false;
                if (l >= 0) hitL = BoxHitFast(bvh.minX[l], bvh.minY[l], bvh.minZ[l], bvh.maxX[l], bvh.maxY[l], bvh.maxZ[l], r, tMin, closest, invDx, invDy, invDz, lNear, lFar);
                if (rr >= 0) hitR = BoxHitFast(bvh.minX[rr], bvh.minY[rr], bvh.minZ[rr], bvh.maxX[rr], bvh.maxY[rr], bvh.maxZ[rr], r, tMin, closest, invDx, invDy, invDz, rNear, rFar);
                if (hitL & hitR) {
                    if (lNear < rNear) { stack[sp++] = rr; stack[sp++] = l; }
                    else { stack[sp++] = l; stack[sp++] = rr; }
                } else if (hitL) stack[sp++] = l;
                else if (hitR) stack[sp++] = rr;
            }
        }
        if (hitAnything) rec = best;
        return hitAnything;
    }
    bool Occluded(const Ray &r, float maxDist, float screenU, float screenV) const { /* Scenes/Scene.cs:77-82 */
        HitRecord rec;
        return Hit(r, 0.001f, maxDist, rec, screenU, screenV);
    }
    /* brute force over Scene.Objects in order, for BVH-vs-linear agreement tests */
    bool HitLinear(const Ray &r, float tMin, float tMax, HitRecord &rec) const {
        bool any = false; float closest = tMax;
        for (int i = 0; i < (int)objects.size(); i++) {
            HitRecord tmp;
            if (ObjectHit(i, r, tMin, closest, tmp, 0.25f, 0.25f)) { any = true; closest = tmp.T; rec = tmp; rec.ObjId = i; }
        }
        return any;
    }

    int Upload(const ycge_scene *s) {
        bgTop = Vec3(s->bg_top[0], s->bg_top[1], s->bg_top[2]);
        bgBottom = Vec3(s->bg_bottom[0], s->bg_bottom[1], s->bg_bottom[2]);
        ambientColor = Vec3(s->ambient_color[0], s->ambient_color[1], s->ambient_color[2]);
        ambientIntensity = s->ambient_intensity;
        isVolumeScene = s->is_volume_scene != 0;
        lights.assign(s->lights, s->lights + s->n_lights);
        materials.clear();
        for (int i = 0; i < s->n_materials; i++) materials.push_back(FromAbi(s->materials[i]));
        objects.clear(); objMesh.clear(); objVol.clear();
        for (int i = 0; i < s->n_objects; i++) {
            const ycge_object &a = s->objects[i];
            Object o;
            o.kind = a.kind; o.matA = a.mat_a; o.matB = a.mat_b; o.checkerScale = a.checker_scale; o.overrideSR = a.override_sr;
            o.specular = a.specular; o.reflectivity = a.reflectivity; o.refId = a.ref_id;
            memcpy(o.p, a.p, sizeof o.p);
            const float *p = o.p;
            const MeshBVH *mp = nullptr; const VolumeGrid *vp = nullptr;
            switch (o.kind) {
                case YCGE_PLANE: o.ndot = p[3] * p[0] + p[4] * p[1] + p[5] * p[2]; break;               /* Surfaces.cs:26 */
                case YCGE_DISK: o.ndot = Vec3(p[3], p[4], p[5]).Dot(Vec3(p[0], p[1], p[2])); o.radius2 = p[6] * p[6]; break; /* :92-93 */
                case YCGE_CYLINDER_Y: o.radius2 = p[3] * p[3]; break;                                 /* BoundedObjects.cs:136 */
                case YCGE_TRIANGLE: { /* Triangle.cs:36-44 */
                    o.e1x = p[3] - p[0]; o.e1y = p[4] - p[1]; o.e1z = p[5] - p[2];
                    o.e2x = p[6] - p[0]; o.e2y = p[7] - p[1]; o.e2z = p[8] - p[2];
                    float nnx = o.e1y * o.e2z - o.e1z * o.e2y, nny = o.e1z * o.e2x - o.e1x * o.e2z, nnz = o.e1x * o.e2y - o.e1y * o.e2x;
                    float invLen = 1.0f / MaxF(1e-20f, SqrtF(nnx * nnx + nny * nny + nnz * nnz));
                    o.tnx = nnx * invLen; o.tny = nny * invLen; o.tnz = nnz * invLen;
                    break; }
                case YCGE_MESH: { auto it = meshes->find(o.refId); if (it == meshes->end()) return YCGE_ERR_INVALID; mp = it->second.get(); break; }
                case YCGE_VOLUME: { auto it = volumes->find(o.refId); if (it == volumes->end()) return YCGE_ERR_INVALID; vp = it->second.get(); break; }
                default: break;
            }
            objects.push_back(o); objMesh.push_back(mp); objVol.push_back(vp);
        }
        if (s->bvh) { BvhFromAbi(s->bvh, bvh); sortFallbacks = 0; }
        else { /* new BVH(Objects)  BVH.cs:29-97 */
            std::vector<Item> items;
            for (int i = 0; i < (int)objects.size(); i++) {
                float b[9];
                if (!TryGetBounds(i, b)) return YCGE_ERR_UNBOUNDED;
                Item it; it.Index = i;
                it.MinX = b[0]; it.MinY = b[1]; it.MinZ = b[2]; it.MaxX = b[3]; it.MaxY = b[4]; it.MaxZ = b[5]; it.Cx = b[6]; it.Cy = b[7]; it.Cz = b[8];
                items.push_back(it);
            }
            Builder bd; bd.TargetLeafSize = 4; bd.meshVariant = false;
            bd.Build(items, bvh);
            sortFallbacks = bd.sortFallbacks;
        }
        return 0;
    }
};

/* ======================================================================================
 * RaytraceSampler (RayTracing/RaytraceSampler.cs) and Rng.cs
 * ====================================================================================== */
static const uint8_t BlueNoise8x8[8][8] = { /* RaytraceSampler.cs:9-19 */
    {0, 32, 8, 40, 2, 34, 10, 42},   {48, 16, 56, 24, 50, 18, 58, 26}, {12, 44, 4, 36, 14, 46, 6, 38},
    {60, 28, 52, 20, 62, 30, 54, 22}, {3, 35, 11, 43, 1, 33, 9, 41},    {51, 19, 59, 27, 49, 17, 57, 25},
    {15, 47, 7, 39, 13, 45, 5, 37},  {63, 31, 55, 23, 61, 29, 53, 21}};
static inline float Frac(float v) { return v - FloorF(v); } /* :22-25 */
static float BlueNoiseSample(int x, int y, int frameIdx, int channel) { /* :27-34 */
    int ix = x & 7, iy = y & 7;
    float baseVal = (BlueNoise8x8[iy][ix] + 0.5f) * (1.0f / (8 * 8));
    float rot = Frac((frameIdx + 1) * (channel == 0 ? 0.7548776662466927f : 0.5698402909980532f));
    return Frac(baseVal + rot);
}
static inline uint64_t SplitMix64(uint64_t z) { /* :71-80 */
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static uint64_t PerFrameSeed(int x, int y, int64_t frame, int jx, int jy, uint64_t salt) { /* :56-68 */
    uint64_t h = 1469598103934665603ULL;
    h ^= (uint64_t)(int64_t)x * 0x9E3779B97F4A7C15ULL; h = SplitMix64(h);
    h ^= (uint64_t)(int64_t)y * 0xC2B2AE3D27D4EB4FULL; h = SplitMix64(h);
    h ^= (uint64_t)frame * 0x165667B19E3779F9ULL; h = SplitMix64(h);
    h ^= ((uint64_t)(uint8_t)jx << 8) ^ (uint64_t)(uint8_t)jy; h = SplitMix64(h);
    h ^= salt; h = SplitMix64(h);
    return h;
}
struct Rng { /* :36-53 */
    uint64_t state;
    explicit Rng(uint64_t seed) : state(seed != 0 ? seed : 0x9E3779B97F4A7C15ULL) {}
    float NextUnit() {
        state = SplitMix64(state);
        uint32_t m24 = (uint32_t)(state >> 40);
        return (m24 + 0.5f) * (1.0f / 16777216.0f);
    }
};
struct RngCs { /* ConsoleRayTracing.Rng  Rng.cs:3-29 (unused by the renderer; restated for completeness) */
    uint64_t state;
    static uint64_t Scramble(uint64_t x) {
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 27; x *= 0x94D049BB133111EBULL; x ^= x >> 31;
        return x;
    }
    explicit RngCs(uint64_t seed) { state = seed + 0x9E3779B97F4A7C15ULL; state = Scramble(state); }
    float NextUnit() {
        state += 0x9E3779B97F4A7C15ULL;
        uint64_t z = Scramble(state);
        return (float)((double)(z >> 11) * (1.0 / 9007199254740992.0));
    }
};
static Vec3 CosineSampleHemisphere(Vec3 n, Rng &rng) { /* :83-111 */
    float u1 = rng.NextUnit();
    float u2 = rng.NextUnit();
    float r = SqrtF(u1);
    float phi = 6.2831853071795864769f * u2;
    float sn, cs;
    m_sincos(phi, &sn, &cs);
    float x = r * cs, y = r * sn;
    float z = SqrtF(1.0f - u1);
    Vec3 w = n;
    float wz = w.Z;
    if (wz < -0.999999f) {
        Vec3 u(0.0f, -1.0f, 0.0f), v(-1.0f, 0.0f, 0.0f);
        return u * x + v * y + w * z;
    }
    float a = 1.0f / (1.0f + wz);
    float b = (-w.X * w.Y) * a;
    Vec3 uAxis = Vec3::D(1.0 - (double)((w.X * w.X) * a), (double)b, (double)(-w.X));
    Vec3 vAxis = Vec3::D((double)b, 1.0 - (double)((w.Y * w.Y) * a), (double)(-w.Y));
    return uAxis * x + vAxis * y + w * z;
}

/* ======================================================================================
 * Cell quantisation (Renderer/Chexel.cs, ANSITerminalRenderer.cs, Win32TerminalRenderer.cs)
 * ====================================================================================== */
static const float Palette16[16][3] = { /* Chexel.cs:11-29 */
    {0.00f, 0.00f, 0.00f}, {0.00f, 0.00f, 0.50f}, {0.00f, 0.50f, 0.00f}, {0.00f, 0.50f, 0.50f}, {0.50f, 0.00f, 0.00f}, {0.50f, 0.00f, 0.50f},
    {0.50f, 0.50f, 0.00f}, {0.75f, 0.75f, 0.75f}, {0.50f, 0.50f, 0.50f}, {0.00f, 0.00f, 1.00f}, {0.00f, 1.00f, 0.00f}, {0.00f, 1.00f, 1.00f},
    {1.00f, 0.00f, 0.00f}, {1.00f, 0.00f, 1.00f}, {1.00f, 1.00f, 0.00f}, {1.00f, 1.00f, 1.00f}};
static Vec3 ChexelClamp01(Vec3 c) { /* Chexel.cs:90-96 (double compare, value unchanged) */
    double rx = c.X < 0.0 ? 0.0 : (c.X > 1.0 ? 1.0 : c.X);
    double ry = c.Y < 0.0 ? 0.0 : (c.Y > 1.0 ? 1.0 : c.Y);
    double rz = c.Z < 0.0 ? 0.0 : (c.Z > 1.0 ? 1.0 : c.Z);
    return Vec3::D(rx, ry, rz);
}
static int NearestConsoleColorFrom(Vec3 v) { /* Chexel.cs:70-88 */
    int best = 0;
    float bestD = FloatMax;
    for (int i = 0; i < 16; i++) {
        float dr = v.X - Palette16[i][0], dg = v.Y - Palette16[i][1], db = v.Z - Palette16[i][2];
        float d = dr * dr + dg * dg + db * db;
        if (d < bestD) { bestD = d; best = i; }
    }
    return best;
}
static uint8_t LinearToSrgb8(double c) { /* ANSITerminalRenderer.cs:298-307; Math.Pow = libm pow (binary64) */
    if (c < 0.0) c = 0.0;
    if (c > 1.0) c = 1.0;
    double s = c <= 0.0031308 ? 12.92 * c : 1.055 * std::pow(c, 1.0 / 2.4) - 0.055;
    int v = (int)std::nearbyint(s * 255.0); /* Math.Round: to nearest, ties to even */
    if (v < 0) v = 0;
    if (v > 255) v = 255;
    return (uint8_t)v;
}
static int ToCubeLevelSrgb(uint8_t v) { /* :288-296 */
    if (v < 48) return 0; if (v < 114) return 1; if (v < 154) return 2; if (v < 194) return 3; if (v < 234) return 4; return 5;
}
static int ChexelToAnsi256(Vec3 color_f32) { /* :246-286, including the never-filled s_graySrgb table (:26) */
    static const uint8_t s_cubeSrgb[6] = {0, 95, 135, 175, 215, 255};
    static const uint8_t s_graySrgb[24] = {0};
    double rLin = color_f32.X, gLin = color_f32.Y, bLin = color_f32.Z;
    if (rLin < 0.0) rLin = 0.0; if (rLin > 1.0) rLin = 1.0;
    if (gLin < 0.0) gLin = 0.0; if (gLin > 1.0) gLin = 1.0;
    if (bLin < 0.0) bLin = 0.0; if (bLin > 1.0) bLin = 1.0;
    uint8_t rS = LinearToSrgb8(rLin), gS = LinearToSrgb8(gLin), bS = LinearToSrgb8(bLin);
    int ir = ToCubeLevelSrgb(rS), ig = ToCubeLevelSrgb(gS), ib = ToCubeLevelSrgb(bS);
    int idxCube = 16 + 36 * ir + 6 * ig + ib;
    int cubeR = s_cubeSrgb[ir], cubeG = s_cubeSrgb[ig], cubeB = s_cubeSrgb[ib];
    uint8_t ySrgb = LinearToSrgb8(0.2126 * rLin + 0.7152 * gLin + 0.0722 * bLin);
    int grayIdx = (int)std::nearbyint((ySrgb - 8.0) / 10.0);
    if (grayIdx < 0) grayIdx = 0;
    if (grayIdx > 23) grayIdx = 23;
    int grayV = s_graySrgb[grayIdx];
    int idxGray = 232 + grayIdx;
    int drg = std::abs(rS - gS), drb = std::abs(rS - bS), dgb = std::abs(gS - bS);
    int chroma = std::max(drg, std::max(drb, dgb));
    bool allowGray = chroma <= 18;
    auto Dist2 = [](int r1, int g1, int b1, int r2, int g2, int b2) { int dr = r1 - r2, dg = g1 - g2, db = b1 - b2; return dr * dr + dg * dg + db * db; };
    int dCube = Dist2(rS, gS, bS, cubeR, cubeG, cubeB);
    int dGray = allowGray ? Dist2(rS, gS, bS, grayV, grayV, grayV) + 64 : INT32_MAX;
    return dGray < dCube ? idxGray : idxCube;
}

/* ======================================================================================
 * ToneMapper (RayTracing/ToneMapper.cs)
 * ====================================================================================== */
struct ToneMapper {
    float toneExposure = 1.0f, toneGamma = 2.2f;
    bool autoExposure = true;
    float aeKey = 0.18f, aeSpeed = 0.2f, aeExposure = 1.0f, aeMin = 0.10f, aeMax = 1.50f;
    float effectiveExposure = 1.0f;
    float toneSaturation = 2.0f, toneVibrance = 0.0f;
    float lastLogSum = 0; int lastCnt = 0;

    static float Saturate01(float v) { if (v < 0.0f) return 0.0f; if (v > 1.0f) return 1.0f; return v; }
    static float ACESFilm(float x) { /* :247-260 */
        float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
        float num = x * (a * x + b);
        float den = x * (c * x + d) + e;
        float y = den > 0.0f ? num / den : 0.0f;
        if (y < 0.0f) y = 0.0f;
        if (y > 1.0f) y = 1.0f;
        return y;
    }
    Vec3 ApplySaturation(Vec3 srgb01) const { /* :223-238 */
        float r = Saturate01(srgb01.X), g = Saturate01(srgb01.Y), b = Saturate01(srgb01.Z);
        float y = 0.2126f * r + 0.7152f * g + 0.0722f * b;
        float maxc = MaxF(r, MaxF(g, b)), minc = MinF(r, MinF(g, b));
        float chroma = maxc - minc;
        float vibFactor = 1.0f + toneVibrance * (1.0f - chroma);
        float f = toneSaturation * vibFactor;
        float rr = y + (r - y) * f, gg = y + (g - y) * f, bb = y + (b - y) * f;
        return Vec3(Saturate01(rr), Saturate01(gg), Saturate01(bb));
    }
    Vec3 ToneMapAndEncode(Vec3 hdr, float exposure, float gamma) const { /* :204-221 */
        float r = MaxF(0.0f, hdr.X) * exposure, g = MaxF(0.0f, hdr.Y) * exposure, b = MaxF(0.0f, hdr.Z) * exposure;
        r = ACESFilm(r); g = ACESFilm(g); b = ACESFilm(b);
        float invGamma = 1.0f / MaxF(0.1f, gamma);
        float sr = m_pow(Saturate01(r), invGamma), sg = m_pow(Saturate01(g), invGamma), sb = m_pow(Saturate01(b), invGamma);
        return ApplySaturation(Vec3(sr, sg, sb));
    }
    Vec3 MapPixel(Vec3 hdr) const { return ToneMapAndEncode(hdr, effectiveExposure, toneGamma); } /* :155-159 */
};

/* ======================================================================================
 * RaytraceRenderer (RayTracing/RaytraceRenderer.cs) + TemporalAA camera test (TemporalAA.cs:58-76)
 * ====================================================================================== */
struct Renderer {
    ycge_params P;
    int fbW = 0, fbH = 0, ss = 1, hiW = 0, hiH = 0;
    float fovDeg = 45.0f;
    int64_t frameCounter = 0;
    Vec3 camPos = Vec3(0.0f, 1.0f, 0.0f);
    float yaw = 0, pitch = 0;
    /* TemporalAA camera state */
    float lastCamX = NAN, lastCamY = NAN, lastCamZ = NAN, lastYaw = NAN, lastPitch = NAN;
    bool forceResetOnce = false, alwaysReset = false;
    ToneMapper tone;
    Scene scene;
    bool haveScene = false;
    std::map<int, std::shared_ptr<MeshBVH>> meshes;
    std::map<int, std::shared_ptr<VolumeGrid>> volumes;
    std::map<int, std::shared_ptr<Texture>> textures;

    std::vector<Ray> rays;
    std::vector<Vec3> currentHdr, gAlbedo, gNormal, spatialA, spatialB, taaHistory, prevNormal;
    std::vector<float> gDepth, prevDepth;
    std::vector<uint8_t> skyMask, prevSky;
    std::vector<int> primObj, primSub;
    std::vector<float> logSamples;
    const std::vector<Vec3> *denoised = nullptr;
    bool taaHistoryValid = false;
    std::vector<ycge_cell> cells;
    Counters lastCounters;
    double msTrace = 0, msTaa = 0, msAtrous = 0, msExposure = 0, msCells = 0, msTotal = 0, msRaygen = 0;
    std::string err;

    void Alloc() { /* ctor :86-105 / Resize :118-136 */
        hiW = fbW * ss; hiH = fbH * 2 * ss;
        size_t n = (size_t)hiW * hiH;
        rays.assign(n, Ray()); currentHdr.assign(n, Vec3()); gAlbedo.assign(n, Vec3()); gNormal.assign(n, Vec3());
        spatialA.assign(n, Vec3()); spatialB.assign(n, Vec3()); taaHistory.assign(n, Vec3()); prevNormal.assign(n, Vec3());
        gDepth.assign(n, 0.0f); prevDepth.assign(n, 0.0f); skyMask.assign(n, 0); prevSky.assign(n, 0);
        primObj.assign(n, -1); primSub.assign(n, -1);
        cells.assign((size_t)fbW * fbH, ycge_cell());
        taaHistoryValid = false;
    }
    void Resize(int w, int h, int s) { /* :110-138 + TemporalAA.Resize TemporalAA.cs:33-45 */
        ss = s < 1 ? 1 : s; fbW = w; fbH = h;
        Alloc();
        lastCamX = lastCamY = lastCamZ = lastYaw = lastPitch = NAN;
    }
    bool ShouldResetHistory(Vec3 cam, float yw, float pt) const { /* TemporalAA.cs:58-67 */
        float dx = cam.X - lastCamX, dy = cam.Y - lastCamY, dz = cam.Z - lastCamZ;
        float trans = (dx != dx) ? 0.0f : SqrtF(dx * dx + dy * dy + dz * dz);
        float dyaw = (lastYaw != lastYaw) ? 0.0f : AbsF(yw - lastYaw);
        float dpitch = (lastPitch != lastPitch) ? 0.0f : AbsF(pt - lastPitch);
        return trans > MaxF(0.0f, P.motion_trans_reset) || dyaw > MaxF(0.0f, P.motion_rot_reset) || dpitch > MaxF(0.0f, P.motion_rot_reset);
    }

    static Vec3 ForwardFromYawPitch(float yw, float pt) { /* :413-417 */
        float cp = m_cos(pt);
        return Vec3(m_sin(yw) * cp, m_sin(pt), -m_cos(yw) * cp);
    }
    static Ray MakeJitteredRay(Vec3 cam, float yw, float pt, float fov, float aspect, int px, int py, int W, int H, float jitterRotX, float jitterRotY, int frameIdx) { /* :419-437 */
        float jxBase = BlueNoiseSample(px, py, frameIdx, 0);
        float jyBase = BlueNoiseSample(px, py, frameIdx, 1);
        float jx = Frac(jxBase + jitterRotX) - 0.5f;
        float jy = Frac(jyBase + jitterRotY) - 0.5f;
        float u = ((px + 0.5f + jx) / W) * 2.0f - 1.0f;
        float v = 1.0f - ((py + 0.5f + jy) / H) * 2.0f;
        float fovRad = fov * (3.14159274f / 180.0f); /* MathF.PI */
        float halfH = m_tan(0.5f * fovRad);
        float halfW = halfH * aspect;
        Vec3 fwd = ForwardFromYawPitch(yw, pt).Normalized();
        Vec3 worldUp(0.0f, 1.0f, 0.0f);
        Vec3 right = fwd.Cross(worldUp).Normalized();
        Vec3 up = right.Cross(fwd).Normalized();
        Vec3 dir = (fwd + right * (u * halfW) + up * (v * halfH)).Normalized();
        return Ray(cam, dir);
    }

    static Vec3 Reflect(Vec3 v, Vec3 n) { return v - n * (2.0f * v.Dot(n)); }                 /* :800-803 */
    static Vec3 Lerp(Vec3 a, Vec3 b, float t) { return a * (1.0f - t) + b * t; }               /* :805-808 */
    static bool Refract(Vec3 v, Vec3 n, float eta, Vec3 &refrDir) { /* :737-748 */
        float cosi = -MaxF(-1.0f, MinF(1.0f, v.Dot(n)));
        float k = 1.0f - eta * eta * (1.0f - cosi * cosi);
        if (k < 0.0f) { refrDir = Vec3(); return false; }
        refrDir = (v * eta) + (n * (eta * cosi - SqrtF(k)));
        return true;
    }
    static float FresnelSchlick(float cosTheta, float etaI, float etaT) { /* :750-755 */
        float r0 = (etaI - etaT) / (etaI + etaT);
        r0 = r0 * r0;
        return r0 + (1.0f - r0) * m_pow(1.0f - cosTheta, 5.0f);
    }
    static Vec3 OrenNayarBRDF(Vec3 albedo, Vec3 n, Vec3 wo, Vec3 wi, float sigmaRad) { /* :810-831 */
        const float Pi = 3.14159265358979323846f, InvPi = 1.0f / Pi;
        float cosThetaI = MaxF(0.0f, n.Dot(wi));
        float cosThetaO = MaxF(0.0f, n.Dot(wo));
        if (cosThetaI <= 0.0f || cosThetaO <= 0.0f) return Vec3();
        float sinThetaI = SqrtF(MaxF(0.0f, 1.0f - cosThetaI * cosThetaI));
        float sinThetaO = SqrtF(MaxF(0.0f, 1.0f - cosThetaO * cosThetaO));
        Vec3 projI = (wi - n * cosThetaI).Normalized();
        Vec3 projO = (wo - n * cosThetaO).Normalized();
        float cosPhiDiff = MaxF(0.0f, projI.Dot(projO));
        float sigma2 = sigmaRad * sigmaRad;
        float A = 1.0f - (sigma2 / (2.0f * (sigma2 + 0.33f)));
        float B = 0.45f * sigma2 / (sigma2 + 0.09f);
        float sinAlpha = MaxF(sinThetaI, sinThetaO);
        float tanBeta = MinF(sinThetaI / MaxF(1e-6f, cosThetaI), sinThetaO / MaxF(1e-6f, cosThetaO));
        float on = (A + B * cosPhiDiff * sinAlpha * tanBeta);
        Vec3 f = albedo * (on * InvPi);
        return f.Saturate();
    }
    Vec3 SampleAlbedo(const Material &mat, float u, float v) const { /* :724-735 */
        const Texture *tex = nullptr;
        if (mat.Tex >= 0 && scene.textures) { auto it = scene.textures->find(mat.Tex); if (it != scene.textures->end()) tex = it->second.get(); }
        if (tex == nullptr || mat.TextureWeight <= 0.0f) return mat.Albedo;
        float tiles = (float)std::max(1e-6, (double)mat.UVScale);
        Vec3 texel = tex->SampleBilinear(u * tiles, v * tiles);
        float t = mat.TextureWeight < 0.0f ? 0.0f : (mat.TextureWeight > 1.0f ? 1.0f : mat.TextureWeight);
        Vec3 outAlbedo = mat.Albedo * (1.0f - t) + texel * t;
        return outAlbedo.Saturate();
    }
    Vec3 ComputeTransmittanceToLight(const Ray &shadow, float maxDist, float screenU, float screenV) const { /* :757-798 */
        if (scene.isVolumeScene) {
            bool blocked = scene.Occluded(shadow, maxDist, screenU, screenV);
            return blocked ? Vec3() : Vec3(1.0f, 1.0f, 1.0f);
        }
        float transR = 1.0f, transG = 1.0f, transB = 1.0f;
        HitRecord block;
        float tmin = 0.0f + P.eps;
        int counter = 0;
        const float cutoff = 1e-6f;
        while (counter < P.max_refractions && scene.Hit(shadow, tmin, maxDist, block, screenU, screenV)) {
            counter++;
            float tr = block.Mat.Transparency;
            if (tr <= 0.0f) return Vec3();
            Vec3 tint = block.Mat.TransmissionColor;
            float trf = tr;
            transR *= tint.X * trf; transG *= tint.Y * trf; transB *= tint.Z * trf;
            if (transR <= cutoff && transG <= cutoff && transB <= cutoff) return Vec3();
            float tHit = block.T;
            if (tHit > maxDist) break;
            tmin = tHit + P.eps;
        }
        return Vec3(transR, transG, transB);
    }

    struct PathWorkItem { Ray ray; Vec3 Throughput; int MirrorDepth, DiffuseDepth; bool IsPrimary; };
    struct PrimaryGBuffer { Vec3 Albedo, Normal; float Depth; int ObjId, SubId; };

    Vec3 TraceFull(const Ray &r, Rng &rng, float screenU, float screenV, bool &isSky, PrimaryGBuffer &primary) const { /* :448-620 */
        const int MaxStack = 16;
        const float Pi = 3.14159265358979323846f;
        PathWorkItem stack[MaxStack];
        int sp = 0;
        stack[sp++] = PathWorkItem{r, Vec3(1, 1, 1), 0, 0, true};
        Vec3 radiance;
        bool primaryHitSomething = false;
        isSky = false;
        bool gbufValid = false;
        primary = PrimaryGBuffer{Vec3(), Vec3(), FloatMax, -1, -1};
        float sigmaRad = P.diffuse_sigma_deg * (3.14159274f / 180.0f);
        const float Eps = P.eps;
        while (sp > 0) {
            sp--;
            PathWorkItem item = stack[sp];
            Ray currentRay = item.ray;
            Vec3 beta = item.Throughput;
            int mirrorDepth = item.MirrorDepth;
            int diffuseDepth = item.DiffuseDepth;
            for (;;) {
                HitRecord rec;
                if (!scene.Hit(currentRay, 0.001f, FloatMax, rec, screenU, screenV)) {
                    float tbg = 0.5f * (currentRay.Dir.Y + 1.0f);
                    Vec3 sky = Lerp(scene.bgBottom, scene.bgTop, tbg);
                    if (item.IsPrimary && !primaryHitSomething) {
                        isSky = true;
                        if (!gbufValid) { primary = PrimaryGBuffer{Vec3(), Vec3(), FloatMax, -1, -1}; gbufValid = true; }
                    }
                    radiance = radiance + Vec3(beta.X * sky.X, beta.Y * sky.Y, beta.Z * sky.Z);
                    break;
                }
                if (item.IsPrimary) {
                    primaryHitSomething = true;
                    isSky = false;
                    if (!gbufValid) {
                        Vec3 baseAlb = SampleAlbedo(rec.Mat, rec.U, rec.V);
                        primary = PrimaryGBuffer{baseAlb, rec.N, rec.T, rec.ObjId, rec.SubId};
                        gbufValid = true;
                    }
                    item.IsPrimary = false;
                }
                if (rec.Mat.Emission.X != 0.0f || rec.Mat.Emission.Y != 0.0f || rec.Mat.Emission.Z != 0.0f) {
                    Vec3 e = rec.Mat.Emission;
                    radiance = radiance + Vec3(beta.X * e.X, beta.Y * e.Y, beta.Z * e.Z);
                }
                Vec3 baseAlbedo = SampleAlbedo(rec.Mat, rec.U, rec.V);
                if (rec.Mat.Transparency > 0.0f) {
                    if (mirrorDepth >= P.max_mirror_bounces) break;
                    Vec3 n = rec.N;
                    Vec3 wo = currentRay.Dir;
                    bool frontFace = n.Dot(wo) < 0.0f;
                    Vec3 nl = frontFace ? n : n * -1.0f;
                    float etaI = frontFace ? 1.0f : rec.Mat.IndexOfRefraction;
                    float etaT = frontFace ? rec.Mat.IndexOfRefraction : 1.0f;
                    float eta = etaI / etaT;
                    Vec3 reflDir = Reflect(wo, nl).Normalized();
                    Vec3 refrDir;
                    bool hasRefract = Refract(wo, nl, eta, refrDir);
                    float cosTheta = AbsF(nl.Dot(wo * -1.0f));
                    float fresnel = FresnelSchlick(cosTheta, etaI, etaT);
                    float R = fresnel;
                    float Tr = rec.Mat.Transparency < 0.0f ? 0.0f : (rec.Mat.Transparency > 1.0f ? 1.0f : rec.Mat.Transparency);
                    float T = hasRefract ? (1.0f - R) * Tr : 0.0f;
                    { float v = R + rec.Mat.Reflectivity * (1.0f - R); R = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
                    if (R > 0.0f) {
                        if (sp < MaxStack) {
                            PathWorkItem refl;
                            refl.ray = Ray(rec.P + nl * Eps, reflDir);
                            refl.Throughput = Vec3(beta.X * baseAlbedo.X * R, beta.Y * baseAlbedo.Y * R, beta.Z * baseAlbedo.Z * R);
                            refl.MirrorDepth = mirrorDepth + 1; refl.DiffuseDepth = diffuseDepth; refl.IsPrimary = false;
                            stack[sp++] = refl;
                        }
                    }
                    if (T > 0.0f) {
                        if (sp < MaxStack) {
                            PathWorkItem refr;
                            refr.ray = Ray(rec.P - nl * Eps, refrDir.Normalized());
                            Vec3 transTint = rec.Mat.TransmissionColor;
                            refr.Throughput = Vec3(beta.X * transTint.X * T, beta.Y * transTint.Y * T, beta.Z * transTint.Z * T);
                            refr.MirrorDepth = mirrorDepth + 1; refr.DiffuseDepth = diffuseDepth; refr.IsPrimary = false;
                            stack[sp++] = refr;
                        }
                    }
                    break;
                }
                if (rec.Mat.Reflectivity >= P.mirror_threshold) {
                    if (mirrorDepth >= P.max_mirror_bounces) break;
                    Vec3 reflDir = Reflect(currentRay.Dir, rec.N).Normalized();
                    currentRay = Ray(rec.P + rec.N * Eps, reflDir);
                    beta = Vec3(beta.X * baseAlbedo.X, beta.Y * baseAlbedo.Y, beta.Z * baseAlbedo.Z);
                    mirrorDepth++;
                    continue;
                }
                if (scene.ambientIntensity > 0.0f) {
                    Vec3 a(scene.ambientColor.X * scene.ambientIntensity, scene.ambientColor.Y * scene.ambientIntensity, scene.ambientColor.Z * scene.ambientIntensity);
                    Vec3 amb(a.X * baseAlbedo.X, a.Y * baseAlbedo.Y, a.Z * baseAlbedo.Z);
                    radiance = radiance + Vec3(beta.X * amb.X, beta.Y * amb.Y, beta.Z * amb.Z);
                }
                Vec3 woView = (currentRay.Dir * -1.0f).Normalized();
                for (size_t i = 0; i < scene.lights.size(); i++) {
                    const ycge_light &light = scene.lights[i];
                    Vec3 toL = Vec3(light.pos[0], light.pos[1], light.pos[2]) - rec.P;
                    float dist2 = toL.Dot(toL);
                    float dist = SqrtF(dist2);
                    Vec3 ldir = toL / dist;
                    float nDotL = MaxF(0.0f, rec.N.Dot(ldir));
                    if (nDotL <= 0.0f) continue;
                    Ray shadow(rec.P + rec.N * Eps, ldir);
                    Vec3 transToLight = ComputeTransmittanceToLight(shadow, dist - Eps, screenU, screenV);
                    if (transToLight.X <= 1e-6f && transToLight.Y <= 1e-6f && transToLight.Z <= 1e-6f) continue;
                    float atten = light.intensity / dist2;
                    Vec3 fDiffuse = OrenNayarBRDF(baseAlbedo, rec.N, woView, ldir, sigmaRad);
                    Vec3 Li = Vec3(light.color[0], light.color[1], light.color[2]) * atten;
                    Vec3 contrib = (fDiffuse * nDotL) * Li;
                    contrib = Vec3(contrib.X * transToLight.X, contrib.Y * transToLight.Y, contrib.Z * transToLight.Z);
                    radiance = radiance + Vec3(beta.X * contrib.X, beta.Y * contrib.Y, beta.Z * contrib.Z);
                }
                if (diffuseDepth < P.diffuse_bounces) {
                    Vec3 bounceDir = CosineSampleHemisphere(rec.N, rng);
                    Vec3 fON = OrenNayarBRDF(baseAlbedo, rec.N, woView, bounceDir, sigmaRad);
                    float factor = Pi;
                    Vec3 mult(fON.X * factor, fON.Y * factor, fON.Z * factor);
                    currentRay = Ray(rec.P + rec.N * Eps, bounceDir);
                    beta = Vec3(beta.X * mult.X, beta.Y * mult.Y, beta.Z * mult.Z);
                    diffuseDepth++;
                    continue;
                }
                break;
            }
        }
        return radiance;
    }

    static float Luma(Vec3 c) { return 0.2126f * c.X + 0.7152f * c.Y + 0.0722f * c.Z; } /* :269-272 */

    const std::vector<Vec3> &TemporalBlendWithClamp(bool forceReset) { /* :274-398 */
        int w = hiW, h = hiH;
        if (!taaHistoryValid || forceReset) {
            for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
                size_t i = (size_t)x + (size_t)y * w;
                taaHistory[i] = currentHdr[i]; prevNormal[i] = gNormal[i]; prevDepth[i] = gDepth[i]; prevSky[i] = skyMask[i];
            }
            taaHistoryValid = true;
            return taaHistory;
        }
        float alpha = MaxF(0.0f, MinF(1.0f, P.taa_alpha));
        int r = 1;
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) {
                size_t i = (size_t)x + (size_t)y * w;
                Vec3 cur = currentHdr[i];
                Vec3 prev = taaHistory[i];
                bool skyNow = skyMask[i], skyPrev = prevSky[i];
                float localAlpha = alpha;
                if (skyNow != skyPrev) localAlpha = 1.0f;
                else {
                    float zNow = gDepth[i], zPrev = prevDepth[i];
                    Vec3 nNow = gNormal[i].Normalized(), nPrev = prevNormal[i].Normalized();
                    if (!std::isfinite(zNow) || !std::isfinite(zPrev)) localAlpha = 1.0f;
                    else {
                        float dz = AbsF(zNow - zPrev);
                        float rel = dz / MaxF(1e-4f, MinF(zNow, zPrev));
                        float ndot = nNow.Dot(nPrev);
                        if (rel > 0.05f || ndot < 0.8f) localAlpha = 1.0f;
                    }
                }
                float minL = PosInf, maxL = NegInf;
                for (int oy = -r; oy <= r; oy++) {
                    int sy = y + oy; if (sy < 0) sy = 0; else if (sy >= h) sy = h - 1;
                    for (int ox = -r; ox <= r; ox++) {
                        int sx = x + ox; if (sx < 0) sx = 0; else if (sx >= w) sx = w - 1;
                        size_t si = (size_t)sx + (size_t)sy * w;
                        if (skyMask[si] != skyMask[i]) continue;
                        float l = Luma(currentHdr[si]);
                        if (l < minL) minL = l;
                        if (l > maxL) maxL = l;
                    }
                }
                float pad = P.luminance_pad;
                float range = maxL - minL;
                float lMin = minL - range * pad, lMax = maxL + range * pad;
                float prevL = Luma(prev);
                if (prevL > lMax) { float s = lMax / MaxF(1e-6f, prevL); prev = Vec3(prev.X * s, prev.Y * s, prev.Z * s); }
                else if (prevL < lMin) { float s = lMin / MaxF(1e-6f, prevL); prev = Vec3(prev.X * s, prev.Y * s, prev.Z * s); }
                Vec3 outC(prev.X * (1.0f - localAlpha) + cur.X * localAlpha, prev.Y * (1.0f - localAlpha) + cur.Y * localAlpha,
                          prev.Z * (1.0f - localAlpha) + cur.Z * localAlpha);
                taaHistory[i] = outC;
            }
        }
        for (size_t i = 0; i < (size_t)w * h; i++) { prevNormal[i] = gNormal[i]; prevDepth[i] = gDepth[i]; prevSky[i] = skyMask[i]; }
        return taaHistory;
    }

    /* One à-trous pass over rows [y0,y1)  :655-715.  (Row-parallel only in the non-reference "fast CPU" mode.) */
    void AtrousRows(const std::vector<Vec3> &cur, std::vector<Vec3> &dst, int step, int y0, int y1) const {
        int w = hiW, h = hiH;
        const float k[5] = {1.f / 16.f, 1.f / 4.f, 3.f / 8.f, 1.f / 4.f, 1.f / 16.f};
        float cPhi = P.c_phi, nPhi = P.n_phi, zPhi = P.z_phi, aPhi = P.a_phi;
        for (int y = y0; y < y1; y++) {
            for (int x = 0; x < w; x++) {
                size_t i0 = (size_t)x + (size_t)y * w;
                if (skyMask[i0]) { dst[i0] = cur[i0]; continue; }
                Vec3 c0 = cur[i0], a0 = gAlbedo[i0], n0 = gNormal[i0].Normalized();
                float z0 = gDepth[i0];
                float wsum = 0.0f;
                Vec3 accum;
                for (int ky = -2; ky <= 2; ky++) {
                    int sy = y + ky * step; if (sy < 0) sy = 0; else if (sy >= h) sy = h - 1;
                    float wy = k[ky + 2];
                    for (int kx = -2; kx <= 2; kx++) {
                        int sx = x + kx * step; if (sx < 0) sx = 0; else if (sx >= w) sx = w - 1;
                        size_t si = (size_t)sx + (size_t)sy * w;
                        if (skyMask[si] != skyMask[i0]) continue;
                        float wx = k[kx + 2];
                        float wBase = wx * wy;
                        Vec3 c = cur[si], a = gAlbedo[si], n = gNormal[si].Normalized();
                        float z = gDepth[si];
                        float lum0 = 0.2126f * c0.X + 0.7152f * c0.Y + 0.0722f * c0.Z;
                        float lum = 0.2126f * c.X + 0.7152f * c.Y + 0.0722f * c.Z;
                        float dl = AbsF(lum - lum0);
                        float dn = MaxF(0.0f, 1.0f - n0.Dot(n));
                        float dz = AbsF(z - z0);
                        float da = AbsF(a.X - a0.X) + AbsF(a.Y - a0.Y) + AbsF(a.Z - a0.Z);
                        float wc = m_exp(-dl / MaxF(1e-6f, cPhi));
                        float wn = m_exp(-dn / MaxF(1e-6f, nPhi));
                        float wz = m_exp(-dz / MaxF(1e-6f, zPhi));
                        float wa = m_exp(-(da) / MaxF(1e-6f, aPhi));
                        float wght = wBase * wc * wn * wz * wa;
                        accum = Vec3(accum.X + c.X * wght, accum.Y + c.Y * wght, accum.Z + c.Z * wght);
                        wsum += wght;
                    }
                }
                if (wsum > 1e-8f) { float inv = 1.0f / wsum; dst[i0] = Vec3(accum.X * inv, accum.Y * inv, accum.Z * inv); }
                else dst[i0] = c0;
            }
        }
    }

    void UpdateExposure(const std::vector<Vec3> &hdr, int sampleStep) { /* ToneMapper.cs:49-91 */
        int w = hiW, h = hiH;
        if (!tone.autoExposure) { tone.effectiveExposure = tone.toneExposure * tone.aeExposure; return; }
        int step = std::max(2, sampleStep);
        float logSum = 0.0f;
        int cnt = 0;
        int sw = (w + step - 1) / step, sh = (h + step - 1) / step;
        logSamples.assign((size_t)sw * sh, NAN);
        for (int py = 0; py < h; py += step) {
            for (int px = 0; px < w; px += step) {
                size_t i = (size_t)px + (size_t)py * w;
                if (skyMask[i]) continue;
                Vec3 c = hdr[i];
                float lum = 0.2126f * c.X + 0.7152f * c.Y + 0.0722f * c.Z;
                if (lum > 0.0f) {
                    float lv = m_log(1e-6f + lum);
                    logSamples[(size_t)(px / step) + (size_t)(py / step) * sw] = lv;
                    logSum += lv;
                    cnt++;
                }
            }
        }
        float avgLog = cnt > 0 ? logSum / std::max(1, cnt) : 0.0f;
        float avgLum = m_exp(avgLog);
        float targetExp = cnt > 0 ? tone.aeKey / MaxF(1e-6f, avgLum) : tone.aeExposure;
        if (targetExp < tone.aeMin) targetExp = tone.aeMin;
        if (targetExp > tone.aeMax) targetExp = tone.aeMax;
        float s = 1.0f - m_exp(-tone.aeSpeed);
        tone.aeExposure = tone.aeExposure + (targetExp - tone.aeExposure) * s;
        tone.effectiveExposure = tone.toneExposure * tone.aeExposure;
        tone.lastLogSum = logSum; tone.lastCnt = cnt;
    }

    template <class F> static void ParallelFor(int n, int threads, F f) {
        if (threads <= 1) { for (int i = 0; i < n; i++) f(i); return; }
        std::vector<std::thread> th;
        for (int i = 0; i < n; i++) th.emplace_back([=]() { f(i); });
        for (auto &t : th) t.join();
    }

    /* TryFlipAndBlit  :157-267.  threads = procCount of the reference's partitioning (raygen + cells in row
     * bands, trace over all threads, TAA / à-trous / exposure single-threaded). fastPost: non-reference option
     * that row-parallelises à-trous (identical results) — used only to shorten test time. */
    int RenderFrame(int threads, bool fastPost, bool countEvents) {
        if (!haveScene) { err = "Scene BVH not built; upload a scene first"; return YCGE_ERR_NO_SCENE; }
        using clk = std::chrono::steady_clock;
        auto t0 = clk::now();
        int procCount = threads < 1 ? 1 : threads;
        float aspect = hiW / (float)hiH;
        Vec3 camPosSnapshot = camPos; float yawSnapshot = yaw, pitchSnapshot = pitch;
        bool resetHistory = ShouldResetHistory(camPosSnapshot, yawSnapshot, pitchSnapshot) || alwaysReset || forceResetOnce;
        forceResetOnce = false;
        int64_t frame = ++frameCounter;
        int frameIdx = (int)(frame & 0x7fffffff);
        float jitterRotX = Frac((frameIdx + 1) * 0.61803398875f);
        float jitterRotY = Frac((frameIdx + 1) * 0.38196601125f);

        ParallelFor(procCount, procCount, [&](int worker) {
            int yStart = (int)((int64_t)worker * hiH / procCount), yEnd = (int)((int64_t)(worker + 1) * hiH / procCount);
            for (int py = yStart; py < yEnd; py++)
                for (int px = 0; px < hiW; px++)
                    rays[(size_t)px + (size_t)py * hiW] = MakeJitteredRay(camPosSnapshot, yawSnapshot, pitchSnapshot, fovDeg, aspect, px, py, hiW, hiH, jitterRotX, jitterRotY, frameIdx);
        });
        auto t1 = clk::now();

        std::vector<Counters> cnts(procCount);
        std::atomic<int> nextRow{0};
        ParallelFor(procCount, procCount, [&](int worker) {
            tl_cnt = &cnts[worker];
            if (!countEvents) { /* rays are always counted (cheap); the detailed events only on request */ }
            for (;;) {
                int py = nextRow.fetch_add(1);
                if (py >= hiH) break;
                for (int px = 0; px < hiW; px++) {
                    size_t i = (size_t)px + (size_t)py * hiW;
                    Rng rng(PerFrameSeed(px, py, frame, 0, 0, P.seed_salt));
                    float uCenter = (px + 0.5f) / hiW, vCenter = (py + 0.5f) / hiH;
                    bool isSky; PrimaryGBuffer gbuf;
                    Vec3 cur = TraceFull(rays[i], rng, uCenter, vCenter, isSky, gbuf);
                    skyMask[i] = isSky; currentHdr[i] = cur; gAlbedo[i] = gbuf.Albedo; gNormal[i] = gbuf.Normal; gDepth[i] = gbuf.Depth;
                    primObj[i] = gbuf.ObjId; primSub[i] = gbuf.SubId;
                }
            }
            tl_cnt = nullptr;
        });
        lastCounters = Counters();
        for (auto &c : cnts) lastCounters.add(c);
        auto t2 = clk::now();

        const std::vector<Vec3> &blended = TemporalBlendWithClamp(resetHistory);
        auto t3 = clk::now();

        const std::vector<Vec3> *cur = &blended;
        std::vector<Vec3> *dst = &spatialA;
        int iters = std::max(1, P.atrous_iterations);
        for (int it = 0; it < iters; it++) {
            int step = 1 << it;
            /* NOTE the reference's swap (:718) makes pass 1 run IN PLACE: after pass 0, cur == dst == scratchA, so taps
             * that precede the pixel in row-major order read already-filtered values. Kept literally (serial). */
            if (fastPost && procCount > 1 && cur != dst) {
                const std::vector<Vec3> *c = cur; std::vector<Vec3> *d = dst;
                ParallelFor(procCount, procCount, [&, c, d](int worker) {
                    int y0 = (int)((int64_t)worker * hiH / procCount), y1 = (int)((int64_t)(worker + 1) * hiH / procCount);
                    AtrousRows(*c, *d, step, y0, y1);
                });
            } else AtrousRows(*cur, *dst, step, 0, hiH);
            const std::vector<Vec3> *tmp = cur; cur = dst; dst = (tmp == &spatialA) ? &spatialB : &spatialA;
        }
        denoised = cur;
        auto t4 = clk::now();

        int step = std::max(2, ss * 2);
        UpdateExposure(*denoised, step);
        auto t5 = clk::now();

        const std::vector<Vec3> &den = *denoised;
        ParallelFor(procCount, procCount, [&](int worker) {
            int yStart = (int)((int64_t)worker * fbH / procCount), yEnd = (int)((int64_t)(worker + 1) * fbH / procCount);
            for (int cy = yStart; cy < yEnd; cy++) {
                int yTopPx0 = cy * 2 * ss, yBotPx0 = (cy * 2 + 1) * ss;
                for (int cx = 0; cx < fbW; cx++) {
                    int xPx0 = cx * ss;
                    Vec3 topSum, botSum;
                    for (int sy = 0; sy < ss; sy++) {
                        int yTop = yTopPx0 + sy, yBot = yBotPx0 + sy;
                        for (int sx = 0; sx < ss; sx++) {
                            int x = xPx0 + sx;
                            topSum = topSum + den[(size_t)x + (size_t)yTop * hiW];
                            botSum = botSum + den[(size_t)x + (size_t)yBot * hiW];
                        }
                    }
                    float inv = 1.0f / (ss * ss);
                    Vec3 topAvg(topSum.X * inv, topSum.Y * inv, topSum.Z * inv), botAvg(botSum.X * inv, botSum.Y * inv, botSum.Z * inv);
                    Vec3 topSDR = tone.MapPixel(topAvg), botSDR = tone.MapPixel(botAvg);
                    /* new Chexel('▀', topSDR, botSDR)  Chexel.cs:112-117 -> ChexelColor(Vec3) :37-41 */
                    Vec3 fg = ChexelClamp01(topSDR), bg = ChexelClamp01(botSDR);
                    ycge_cell &c = cells[(size_t)cx + (size_t)cy * fbW];
                    c.glyph = 0x2580;
                    c.fg16 = (uint8_t)NearestConsoleColorFrom(fg); c.bg16 = (uint8_t)NearestConsoleColorFrom(bg);
                    c.fg_ansi = (uint8_t)ChexelToAnsi256(fg); c.bg_ansi = (uint8_t)ChexelToAnsi256(bg);
                    c.attr = (uint16_t)((c.fg16 & 0x0F) | ((c.bg16 & 0x0F) << 4)); /* Win32TerminalRenderer.cs:109-112 */
                    c.fg[0] = fg.X; c.fg[1] = fg.Y; c.fg[2] = fg.Z; c.bg[0] = bg.X; c.bg[1] = bg.Y; c.bg[2] = bg.Z;
                }
            }
        });
        /* taa.CommitCamera  :266 / TemporalAA.cs:69-76 */
        lastCamX = camPosSnapshot.X; lastCamY = camPosSnapshot.Y; lastCamZ = camPosSnapshot.Z; lastYaw = yawSnapshot; lastPitch = pitchSnapshot;
        auto t6 = clk::now();
        auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        msRaygen = ms(t0, t1); msTrace = ms(t1, t2); msTaa = ms(t2, t3); msAtrous = ms(t3, t4); msExposure = ms(t4, t5); msCells = ms(t5, t6); msTotal = ms(t0, t6);
        return 0;
    }
};

} // namespace yo

/* ========================================================================================== C API (ctypes) */
using namespace yo;
#define YO_API extern "C" __attribute__((visibility("default")))

YO_API void yo_set_math_mode(int mode) { g_math_mode = mode; }
YO_API void yo_set_sort_mode(int mode) { g_sort_mode = mode; }

YO_API void yo_default_params(ycge_params *p) {
    memset(p, 0, sizeof *p);
    p->diffuse_bounces = 1; p->max_mirror_bounces = 2; p->max_refractions = 2; p->atrous_iterations = 3;
    p->mirror_threshold = 0.9f; p->eps = 1e-4f; p->taa_alpha = 0.01f; p->motion_trans_reset = 0.0025f; p->motion_rot_reset = 0.0025f;
    p->diffuse_sigma_deg = 25.0f; p->luminance_pad = 0.10f; p->c_phi = 3.0f; p->n_phi = 0.35f; p->z_phi = 2.0f; p->a_phi = 0.20f;
    p->tone_exposure = 1.0f; p->tone_gamma = 2.2f; p->ae_key = 0.18f; p->ae_speed = 0.2f; p->ae_min = 0.10f; p->ae_max = 1.50f;
    p->saturation = 2.0f; p->vibrance = 0.0f; p->auto_exposure = 1; p->seed_salt = 0x9E3779B97F4A7C15ULL;
}
YO_API void *yo_create(const ycge_config *cfg) {
    Renderer *r = new Renderer();
    r->P = cfg->params;
    r->tone.toneExposure = cfg->params.tone_exposure; r->tone.toneGamma = cfg->params.tone_gamma;
    r->tone.autoExposure = cfg->params.auto_exposure != 0; r->tone.aeKey = cfg->params.ae_key; r->tone.aeSpeed = cfg->params.ae_speed;
    r->tone.aeMin = cfg->params.ae_min; r->tone.aeMax = cfg->params.ae_max; r->tone.toneSaturation = cfg->params.saturation; r->tone.toneVibrance = cfg->params.vibrance;
    r->fbW = cfg->fb_w; r->fbH = cfg->fb_h; r->ss = cfg->ss < 1 ? 1 : cfg->ss;
    r->Alloc();
    r->scene.meshes = &r->meshes; r->scene.volumes = &r->volumes; r->scene.textures = &r->textures;
    return r;
}
YO_API void yo_destroy(void *h) { delete (Renderer *)h; }
YO_API int yo_resize(void *h, int w, int hh, int ss) { ((Renderer *)h)->Resize(w, hh, ss); return 0; }
YO_API int yo_mesh_upload_triangles(void *h, int id, int n, const float *abc, const ycge_material *m) {
    auto mb = std::make_shared<MeshBVH>();
    mb->FromTriangles(n, abc, FromAbi(*m));
    ((Renderer *)h)->meshes[id] = mb;
    return 0;
}
YO_API int yo_mesh_upload_soa(void *h, int id, const ycge_mesh_soa *m) {
    auto mb = std::make_shared<MeshBVH>();
    mb->FromSoa(m);
    ((Renderer *)h)->meshes[id] = mb;
    return 0;
}
YO_API int yo_volume_upload(void *h, int id, const ycge_volume *v) {
    auto vg = std::make_shared<VolumeGrid>();
    vg->FromAbi(v);
    ((Renderer *)h)->volumes[id] = vg;
    return 0;
}
YO_API int yo_texture_upload(void *h, int id, int w, int hh, const uint32_t *rgba) {
    auto t = std::make_shared<Texture>();
    t->width = w; t->height = hh;
    if (rgba && w > 0 && hh > 0) t->pixels.assign(rgba, rgba + (size_t)w * hh);
    ((Renderer *)h)->textures[id] = t;
    return 0;
}
YO_API void yo_texture_sample(int w, int hh, const uint32_t *rgba, float u, float v, float *out3) {
    Texture t;
    t.width = w; t.height = hh; t.pixels.assign(rgba, rgba + (size_t)w * hh);
    Vec3 c = t.SampleBilinear(u, v);
    out3[0] = c.X; out3[1] = c.Y; out3[2] = c.Z;
}
YO_API int yo_scene_upload(void *h, const ycge_scene *s) {
    Renderer *r = (Renderer *)h;
    int rc = r->scene.Upload(s);
    r->haveScene = rc == 0;
    return rc;
}
YO_API int yo_lights_update(void *h, int n, const ycge_light *l) { ((Renderer *)h)->scene.lights.assign(l, l + n); return 0; }
YO_API int yo_globals_update(void *h, const float *top, const float *bot, const float *amb, float ai) {
    Scene &s = ((Renderer *)h)->scene;
    s.bgTop = Vec3(top[0], top[1], top[2]); s.bgBottom = Vec3(bot[0], bot[1], bot[2]); s.ambientColor = Vec3(amb[0], amb[1], amb[2]); s.ambientIntensity = ai;
    return 0;
}
YO_API int yo_set_camera(void *h, const float *pos, float yaw, float pitch) {
    Renderer *r = (Renderer *)h; r->camPos = Vec3(pos[0], pos[1], pos[2]); r->yaw = yaw; r->pitch = pitch; return 0;
}
YO_API int yo_set_fov(void *h, float f) { ((Renderer *)h)->fovDeg = f; return 0; }
YO_API int yo_reset_history(void *h) { ((Renderer *)h)->forceResetOnce = true; return 0; }
YO_API int yo_render_frame(void *h, ycge_cell *out, int threads, int fast_post) {
    Renderer *r = (Renderer *)h;
    int rc = r->RenderFrame(threads, fast_post != 0, true);
    if (rc == 0 && out) memcpy(out, r->cells.data(), r->cells.size() * sizeof(ycge_cell));
    return rc;
}
YO_API int yo_get_stats(void *h, ycge_stats *s) {
    Renderer *r = (Renderer *)h;
    memset(s, 0, sizeof *s);
    s->frames = (uint64_t)r->frameCounter; s->rays = r->lastCounters.rays;
    s->top_nodes_popped = r->lastCounters.top_nodes; s->mesh_nodes_popped = r->lastCounters.mesh_nodes; s->leaf_refs = r->lastCounters.leaf_refs;
    s->tris_tested = r->lastCounters.tris; s->prims_tested = r->lastCounters.prims; s->dda_cells = r->lastCounters.dda;
    s->ms_trace = (float)(r->msRaygen + r->msTrace); s->ms_taa = (float)r->msTaa; s->ms_atrous = (float)r->msAtrous; s->ms_exposure = (float)r->msExposure;
    s->ms_cells = (float)r->msCells; s->ms_total = (float)r->msTotal;
    s->ae_exposure = r->tone.aeExposure; s->log_sum = r->tone.lastLogSum; s->log_cnt = r->tone.lastCnt;
    return 0;
}
YO_API int yo_debug_read(void *h, int kind, void *dst, size_t bytes) {
    Renderer *r = (Renderer *)h;
    size_t n = (size_t)r->hiW * r->hiH;
    float *f = (float *)dst;
    auto need = [&](size_t b) { return bytes >= b; };
    switch (kind) {
        case YCGE_DBG_RAYS: if (!need(n * 24)) return YCGE_ERR_INVALID;
            for (size_t i = 0; i < n; i++) { f[6 * i] = r->rays[i].Origin.X; f[6 * i + 1] = r->rays[i].Origin.Y; f[6 * i + 2] = r->rays[i].Origin.Z; f[6 * i + 3] = r->rays[i].Dir.X; f[6 * i + 4] = r->rays[i].Dir.Y; f[6 * i + 5] = r->rays[i].Dir.Z; }
            return 0;
        case YCGE_DBG_HDR: case YCGE_DBG_TAA: case YCGE_DBG_DENOISED: {
            if (!need(n * 16)) return YCGE_ERR_INVALID;
            const std::vector<Vec3> &v = kind == YCGE_DBG_HDR ? r->currentHdr : kind == YCGE_DBG_TAA ? r->taaHistory : *r->denoised;
            for (size_t i = 0; i < n; i++) { f[4 * i] = v[i].X; f[4 * i + 1] = v[i].Y; f[4 * i + 2] = v[i].Z; f[4 * i + 3] = Renderer::Luma(v[i]); }
            return 0; }
        case YCGE_DBG_ALBEDO_SKY: if (!need(n * 16)) return YCGE_ERR_INVALID;
            for (size_t i = 0; i < n; i++) { f[4 * i] = r->gAlbedo[i].X; f[4 * i + 1] = r->gAlbedo[i].Y; f[4 * i + 2] = r->gAlbedo[i].Z; f[4 * i + 3] = r->skyMask[i] ? 1.0f : 0.0f; }
            return 0;
        case YCGE_DBG_NORMAL_DEPTH: if (!need(n * 16)) return YCGE_ERR_INVALID;
            for (size_t i = 0; i < n; i++) { Vec3 nn = r->gNormal[i].Normalized(); f[4 * i] = nn.X; f[4 * i + 1] = nn.Y; f[4 * i + 2] = nn.Z; f[4 * i + 3] = r->gDepth[i]; }
            return 0;
        case YCGE_DBG_PRIM_ID: { if (!need(n * 8)) return YCGE_ERR_INVALID; int *d = (int *)dst;
            for (size_t i = 0; i < n; i++) { d[2 * i] = r->primObj[i]; d[2 * i + 1] = r->primSub[i]; }
            return 0; }
        case YCGE_DBG_LOG_SAMPLES: if (!need(r->logSamples.size() * 4)) return YCGE_ERR_INVALID;
            memcpy(dst, r->logSamples.data(), r->logSamples.size() * 4); return 0;
    }
    return YCGE_ERR_INVALID;
}
/* gNormal as TraceFull wrote it (not normalised), hiW*hiH*3 floats: input of the literal TAA transcription in the tests */
YO_API int yo_debug_raw_normal(void *h, float *dst) {
    Renderer *r = (Renderer *)h;
    for (size_t i = 0; i < r->gNormal.size(); i++) { dst[3 * i] = r->gNormal[i].X; dst[3 * i + 1] = r->gNormal[i].Y; dst[3 * i + 2] = r->gNormal[i].Z; }
    return 0;
}
/* ---- tree export for builder-parity tests: which = -1 top-level, else mesh id ---- */
YO_API int yo_bvh_info(void *h, int which, int *n_nodes, int *root, int *n_leaf, uint64_t *sort_fallbacks) {
    Renderer *r = (Renderer *)h;
    const FlatBVH *b; uint64_t sf;
    if (which < 0) { b = &r->scene.bvh; sf = r->scene.sortFallbacks; }
    else { auto it = r->meshes.find(which); if (it == r->meshes.end()) return YCGE_ERR_INVALID; b = &it->second->bvh; sf = it->second->sortFallbacks; }
    *n_nodes = b->nodes(); *root = b->root; *n_leaf = (int)b->leafIndex.size(); *sort_fallbacks = sf;
    return 0;
}
YO_API int yo_bvh_read(void *h, int which, float *boxes6, int *lrsc4, int *leaf) {
    Renderer *r = (Renderer *)h;
    const FlatBVH *b;
    if (which < 0) b = &r->scene.bvh; else { auto it = r->meshes.find(which); if (it == r->meshes.end()) return YCGE_ERR_INVALID; b = &it->second->bvh; }
    for (int i = 0; i < b->nodes(); i++) {
        boxes6[6 * i] = b->minX[i]; boxes6[6 * i + 1] = b->minY[i]; boxes6[6 * i + 2] = b->minZ[i];
        boxes6[6 * i + 3] = b->maxX[i]; boxes6[6 * i + 4] = b->maxY[i]; boxes6[6 * i + 5] = b->maxZ[i];
        lrsc4[4 * i] = b->left[i]; lrsc4[4 * i + 1] = b->right[i]; lrsc4[4 * i + 2] = b->start[i]; lrsc4[4 * i + 3] = b->count[i];
    }
    memcpy(leaf, b->leafIndex.data(), b->leafIndex.size() * sizeof(int));
    return 0;
}
YO_API int yo_mesh_soa_read(void *h, int id, float *soa12) { /* n*12 floats: ax..az,e1,e2,n per triangle */
    Renderer *r = (Renderer *)h;
    auto it = r->meshes.find(id); if (it == r->meshes.end()) return YCGE_ERR_INVALID;
    const MeshBVH &m = *it->second;
    for (size_t i = 0; i < m.ax.size(); i++) {
        float *d = soa12 + 12 * i;
        d[0] = m.ax[i]; d[1] = m.ay[i]; d[2] = m.az[i]; d[3] = m.e1x[i]; d[4] = m.e1y[i]; d[5] = m.e1z[i];
        d[6] = m.e2x[i]; d[7] = m.e2y[i]; d[8] = m.e2z[i]; d[9] = m.nx[i]; d[10] = m.ny[i]; d[11] = m.nz[i];
    }
    return 0;
}
/* ---- single-ray queries: use_bvh 1 = Scene.Hit, 0 = linear scan over Scene.Objects ---- */
YO_API int yo_scene_hit(void *h, int n, const float *org_dir6, float tmin, float tmax, int use_bvh, float *t_out, int *ids2_out, float *n_out3) {
    Renderer *r = (Renderer *)h;
    for (int i = 0; i < n; i++) {
        const float *q = org_dir6 + 6 * i;
        Ray ray(Vec3(q[0], q[1], q[2]), Vec3(q[3], q[4], q[5]));
        HitRecord rec;
        bool hit = use_bvh ? r->scene.Hit(ray, tmin, tmax, rec, 0.25f, 0.25f) : r->scene.HitLinear(ray, tmin, tmax, rec);
        t_out[i] = hit ? rec.T : -1.0f;
        ids2_out[2 * i] = hit ? rec.ObjId : -1; ids2_out[2 * i + 1] = hit ? rec.SubId : -1;
        if (n_out3) { n_out3[3 * i] = rec.N.X; n_out3[3 * i + 1] = rec.N.Y; n_out3[3 * i + 2] = rec.N.Z; }
    }
    return 0;
}
/* ---- unit-level entry points for known-answer tests ---- */
YO_API uint64_t yo_splitmix64(uint64_t z) { return SplitMix64(z); }
YO_API uint64_t yo_per_frame_seed(int x, int y, int64_t frame, int jx, int jy, uint64_t salt) { return PerFrameSeed(x, y, frame, jx, jy, salt); }
YO_API void yo_rng_draws(uint64_t seed, int n, uint32_t *bits_out, uint32_t *m24_out) {
    Rng g(seed);
    for (int i = 0; i < n; i++) {
        uint64_t next = SplitMix64(g.state);
        if (m24_out) m24_out[i] = (uint32_t)(next >> 40);
        float f = g.NextUnit(); memcpy(&bits_out[i], &f, 4);
    }
}
YO_API void yo_rng_cs_draws(uint64_t seed, int n, uint32_t *bits_out) {
    RngCs g(seed);
    for (int i = 0; i < n; i++) { float f = g.NextUnit(); memcpy(&bits_out[i], &f, 4); }
}
YO_API float yo_blue_noise(int x, int y, int frameIdx, int channel) { return BlueNoiseSample(x, y, frameIdx, channel); }
YO_API int yo_blue_noise_table(int iy, int ix) { return BlueNoise8x8[iy][ix]; }
YO_API int yo_morton3(int x, int y, int z) { return VolumeGrid::Morton3_3bits(x, y, z); }
YO_API int yo_volume_index_of(int nx, int ny, int nz, int ix, int iy, int iz) {
    VolumeGrid g; g.nx = nx; g.ny = ny; g.nz = nz; g.nbx = (nx + 7) >> 3; g.nby = (ny + 7) >> 3; g.nbz = (nz + 7) >> 3;
    return g.IndexOf(ix, iy, iz);
}
YO_API int yo_ansi256(float r, float g, float b) { return ChexelToAnsi256(ChexelClamp01(Vec3(r, g, b))); }
YO_API int yo_linear_to_srgb8(double c) { return LinearToSrgb8(c); }
YO_API int yo_cube_level(int v) { return ToCubeLevelSrgb((uint8_t)v); }
YO_API int yo_nearest16(float r, float g, float b) { return NearestConsoleColorFrom(ChexelClamp01(Vec3(r, g, b))); }
YO_API void yo_tonemap(float exposure, float r, float g, float b, float *out3) {
    ToneMapper t; Vec3 o = t.ToneMapAndEncode(Vec3(r, g, b), exposure, 2.2f); out3[0] = o.X; out3[1] = o.Y; out3[2] = o.Z;
}
YO_API void yo_cosine_sample(float nx, float ny, float nz, uint64_t seed, float *out3) {
    Rng g(seed); Vec3 d = CosineSampleHemisphere(Vec3(nx, ny, nz), g); out3[0] = d.X; out3[1] = d.Y; out3[2] = d.Z;
}
YO_API void yo_dotnet_sort_floats(float *keys, int *payload, int n) { /* sorts (key,payload) pairs like Array.Sort with a.CompareTo(b) */
    struct KP { float k; int p; };
    std::vector<KP> v(n);
    for (int i = 0; i < n; i++) v[i] = KP{keys[i], payload[i]};
    auto cmp = [](const KP &a, const KP &b) { return FloatCompareTo(a.k, b.k); };
    DotnetSort<KP, decltype(cmp)> s{v.data(), cmp};
    s.Sort(0, n);
    for (int i = 0; i < n; i++) { keys[i] = v[i].k; payload[i] = v[i].p; }
}
YO_API float yo_math(int fn, float x, float y) {
    switch (fn) { case 0: return m_exp(x); case 1: return m_log(x); case 2: return m_pow(x, y); case 3: return m_sin(x); case 4: return m_cos(x); case 5: return m_tan(x); }
    return 0;
}
