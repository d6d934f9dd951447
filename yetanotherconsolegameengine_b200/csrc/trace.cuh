// trace.cuh — K1: the per-pixel path tracing megakernel (sm_100a).
// One thread = one pixel path: ray generation (RaytraceRenderer.MakeJitteredRay :419-437), the per-pixel RNG
// stream (RaytraceSampler.cs:36-80), TraceFull (:448-620) with every Hit routine it reaches (BVH.cs, MeshBVH.cs,
// BoundedObjects.cs, Surfaces.cs, Triangle.cs, VolumeGrid.cs) and the primary G-buffer write.
// Arithmetic follows the reference operation by operation in binary32; the translation unit is compiled with
// --fmad=false so no multiply-add is contracted, and with IEEE division / square root (nvcc defaults).
#pragma once
#include "device_types.h"
#include "../../include/ycge.h"
#include "../../include/ycge_detmath.h"

namespace ycge {

// ------------------------------------------------------------------------------------------------ math helpers
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 vdiv(V3 a, float s) { float inv = 1.0f / s; return mk(a.x * inv, a.y * inv, a.z * inv); } // Vec3.cs:68-71
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                       // (x+y)+z
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 normalized(V3 a) { // Vec3.cs:98-107
    float l2 = a.x * a.x + a.y * a.y + a.z * a.z;
    if (l2 <= 0.0f) return a;
    float inv = 1.0f / sqrtf(l2);
    return mk(a.x * inv, a.y * inv, a.z * inv);
}
__device__ __forceinline__ float clamp01(float v) { if (v < 0.0f) return 0.0f; if (v > 1.0f) return 1.0f; return v; }
__device__ __forceinline__ V3 saturate3(V3 a) { return mk(clamp01(a.x), clamp01(a.y), clamp01(a.z)); }

// MathF.Max / MathF.Min: IEEE 754-2019 maximum/minimum (NaN propagates; +0 > -0)
__device__ __forceinline__ float MaxF(float a, float b) {
    if (a != b) return (a != a) ? a : (b < a ? a : b);
    return (__float_as_int(b) < 0) ? a : b;
}
__device__ __forceinline__ float MinF(float a, float b) {
    if (a != b) return (a != a) ? a : (a < b ? a : b);
    return (__float_as_int(a) < 0) ? a : b;
}
// NaN-propagating max/min where the sign of zero and the NaN payload cannot influence any later comparison
__device__ __forceinline__ float max_nan(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float min_nan(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

#define YCGE_FLT_MAX 3.402823466e+38f
#define YCGE_INF __int_as_float(0x7f800000)

struct RayD { V3 o, d; }; // Ray.cs: Dir is always the normalised direction
__device__ __forceinline__ RayD make_ray(V3 o, V3 d) { RayD r; r.o = o; r.d = normalized(d); return r; }

// ------------------------------------------------------------------------------------------------ RNG
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) { // RaytraceSampler.cs:71-80
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long per_frame_seed(int x, int y, long long frame, unsigned long long salt) { // :56-68, jx = jy = 0
    unsigned long long h = 1469598103934665603ULL;
    h ^= (unsigned long long)(long long)x * 0x9E3779B97F4A7C15ULL; h = splitmix64(h);
    h ^= (unsigned long long)(long long)y * 0xC2B2AE3D27D4EB4FULL; h = splitmix64(h);
    h ^= (unsigned long long)frame * 0x165667B19E3779F9ULL; h = splitmix64(h);
    h ^= 0ULL; h = splitmix64(h);
    h ^= salt; h = splitmix64(h);
    return h;
}
__device__ __forceinline__ float rng_next(unsigned long long &state) { // Rng.NextUnit :47-52
    state = splitmix64(state);
    unsigned int m24 = (unsigned int)(state >> 40);
    return ((float)m24 + 0.5f) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ unsigned long long rngcs_scramble(unsigned long long x) { // Rng.cs:20-28
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 27; x *= 0x94D049BB133111EBULL; x ^= x >> 31;
    return x;
}

__constant__ unsigned char c_blue_noise[64] = { // RaytraceSampler.cs:9-19
    0, 32, 8, 40, 2, 34, 10, 42, 48, 16, 56, 24, 50, 18, 58, 26, 12, 44, 4, 36, 14, 46, 6, 38, 60, 28, 52, 20, 62, 30, 54, 22,
    3, 35, 11, 43, 1, 33, 9, 41, 51, 19, 59, 27, 49, 17, 57, 25, 15, 47, 7, 39, 13, 45, 5, 37, 63, 31, 55, 23, 61, 29, 53, 21};
__device__ __forceinline__ float fracf_(float v) { return v - floorf(v); }

// ------------------------------------------------------------------------------------------------ hit record
struct Hit {
    float t;
    V3 P, N;
    int obj, sub;
    int mat;        // material table index
    int albedo_ov;  // 0 none, 1 wire black, 2 wire white (VolumeGrid.cs:44-45)
    int sr_ov;      // 1: reflectivity overridden by the object (Surfaces.cs:64-66)
    float refl;     // overriding reflectivity
};

struct Mat { V3 albedo; float reflectivity; V3 emission; float transparency; V3 transmission; float ior; };
__device__ __forceinline__ Mat load_material(const DevScene &sc, const Hit &h) {
    const float4 *m = sc.materials + 4 * (size_t)h.mat;
    float4 a = __ldg(m), b = __ldg(m + 1), c = __ldg(m + 2);
    Mat r;
    r.albedo = mk(a.x, a.y, a.z); r.reflectivity = a.w;
    r.emission = mk(b.x, b.y, b.z); r.transparency = b.w;
    r.transmission = mk(c.x, c.y, c.z); r.ior = c.w;
    if (h.albedo_ov == 1) r.albedo = mk(0.0f, 0.0f, 0.0f); else if (h.albedo_ov == 2) r.albedo = mk(1.0f, 1.0f, 1.0f);
    if (h.sr_ov) r.reflectivity = h.refl;
    return r;
}

// per-thread traversal stack in local memory: (reference, tNear) pairs, shared by the top-level and the mesh walk
#define YCGE_STACK 96
struct Stack { int ref[YCGE_STACK]; float tn[YCGE_STACK]; };

template <int MODE> struct Cnt {
    unsigned int rays = 0, top_nodes = 0, mesh_nodes = 0, leaf_refs = 0, tris = 0, prims = 0, dda = 0, overflow = 0;
};
// MODE bit 0: count the reference-defined traversal events (ycge_render_frame_stats); bit 1: LEAN -- the scene holds no voxel grid, no texture
// and no transparent material (decided at ycge_scene_upload), so the DDA, the texture sampler and the deferred-branch stack are compiled out:
// a smaller kernel for the mesh / primitive scenes (same arithmetic for what remains, bit-identical)
#define CNT_INC(c, f) do { if (MODE & 1) (c).f++; } while (0)
#define YCGE_LEAN ((MODE & 2) != 0)

// ------------------------------------------------------------------------------------------------ box tests
// BVH.BoxHitFast (BVH.cs:201-236): swap-ordered slabs, NaN-propagating max/min, clamp to [tMin,tMax].
__device__ __forceinline__ bool box_top(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, V3 o, V3 inv, float tMin, float tMax, float &tNear) {
    float e, x;
    e = (mnx - o.x) * inv.x; x = (mxx - o.x) * inv.x; float tEnterX = e, tExitX = x; if (e > x) { tEnterX = x; tExitX = e; }
    e = (mny - o.y) * inv.y; x = (mxy - o.y) * inv.y; float tEnterY = e, tExitY = x; if (e > x) { tEnterY = x; tExitY = e; }
    e = (mnz - o.z) * inv.z; x = (mxz - o.z) * inv.z; float tEnterZ = e, tExitZ = x; if (e > x) { tEnterZ = x; tExitZ = e; }
    float tEnter = max_nan(tEnterX, max_nan(tEnterY, tEnterZ));
    float tExit = min_nan(tExitX, min_nan(tExitY, tExitZ));
    if (tEnter < tMin) tEnter = tMin;
    if (tExit > tMax) tExit = tMax;
    tNear = tEnter;
    return tExit >= tEnter;
}
// MeshBVH.BoxHitFast (MeshBVH.cs:308-332): sign-indexed slabs, comparisons only. The two early-outs of the
// reference cannot change the result (tMin only grows, tMax only shrinks), so one final test is equivalent.
__device__ __forceinline__ bool box_mesh(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, V3 o, V3 inv, int sx, int sy, int sz,
                                         float tMin, float tMax, float &tNear) {
    float en = ((sx == 0 ? mnx : mxx) - o.x) * inv.x, ex = ((sx == 0 ? mxx : mnx) - o.x) * inv.x;
    if (en > tMin) tMin = en;
    if (ex < tMax) tMax = ex;
    en = ((sy == 0 ? mny : mxy) - o.y) * inv.y; ex = ((sy == 0 ? mxy : mny) - o.y) * inv.y;
    if (en > tMin) tMin = en;
    if (ex < tMax) tMax = ex;
    en = ((sz == 0 ? mnz : mxz) - o.z) * inv.z; ex = ((sz == 0 ? mxz : mnz) - o.z) * inv.z;
    if (en > tMin) tMin = en;
    if (ex < tMax) tMax = ex;
    tNear = tMin;
    return tMax >= tMin;
}

// ------------------------------------------------------------------------------------------------ MeshBVH.Hit (MeshBVH.cs:132-304)
// Measured and dropped (B200, dragon 1080p, trace 1.18 ms): (1) parking a leaf while the lane keeps walking inner nodes
// until its next entry is a leaf too ("postponed leaves": same leaf order, same hits — every entry is re-tested against
// the current `closest` when popped — verified bit-exact), with and without warp votes to enter the triangle loop
// together: 1.30 / 1.33 ms; (1b, round 2) the "while-while" shape -- every lane descends through inner nodes until it holds a leaf, the nearer
// child kept in a register instead of pushed and popped, then the lanes of a warp intersect their leaves together -- bit-identical incl. the
// event counters, 1.13 against 1.12 ms (dragon stand-in), 1.00 against 1.02 (bunny): no gain, not kept; (2) scene_hit as one __noinline__ copy instead of three inlined ones (halves the code):
// 1.20 ms; (3) 16..40 resident warps per SM via launch bounds: no change.  The triangle loop runs with 2-3 of 32 lanes
// active and 21 % of the stall samples are instruction-fetch misses, but the kernel is bound by the dependent node /
// triangle fetch latency of the longest paths in each warp, not by those.
template <int MODE>
__device__ bool mesh_hit(const DevMesh &mesh, const RayD &r, float tMin, float tMax, Stack &st, int sp0, Cnt<MODE> &cnt,
                         float &tOut, int &slotOut) {
    V3 inv = mk(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    int sx = inv.x < 0.0f ? 1 : 0, sy = inv.y < 0.0f ? 1 : 0, sz = inv.z < 0.0f ? 1 : 0;
    float closest = tMax;
    bool any = false;
    int sp = sp0;
    { // first pop: the root's own box
        float tn;
        CNT_INC(cnt, mesh_nodes);
        if (!box_mesh(mesh.root.lo[0], mesh.root.lo[1], mesh.root.lo[2], mesh.root.hi[0], mesh.root.hi[1], mesh.root.hi[2], r.o, inv, sx, sy, sz, tMin, closest, tn))
            return false;
        st.ref[sp] = mesh.root.ref; st.tn[sp] = -YCGE_INF; sp++;
    }
    bool first = true;
    while (sp > sp0) {
        sp--;
        int ref = st.ref[sp];
        float tn = st.tn[sp];
        // pop + re-test of the node's own box against the shrunken `closest` == (closest >= tNear at push)
        if (!first) { CNT_INC(cnt, mesh_nodes); if (!(closest >= tn)) continue; }
        first = false;
        if (ref < 0) {
            int v = ~ref;
            int count = (v >> 26) + 1, start = v & YCGE_LEAF_MAX_START;
            for (int i = 0; i < count; i++) {
                CNT_INC(cnt, leaf_refs);
                CNT_INC(cnt, tris);
                const DevTri *tp = mesh.tris + (start + i);
                float4 t0 = __ldg(&tp->t0), t1 = __ldg(&tp->t1), t2 = __ldg(&tp->t2);
                float ax = t0.x, ay = t0.y, az = t0.z, e1x = t0.w, e1y = t1.x, e1z = t1.y, e2x = t1.z, e2y = t1.w, e2z = t2.x;
                // TriHit :239-304
                float px = r.d.y * e2z - r.d.z * e2y;
                float py = r.d.z * e2x - r.d.x * e2z;
                float pz = r.d.x * e2y - r.d.y * e2x;
                float det = e1x * px + e1y * py + e1z * pz;
                const float Eps = 1e-8f;
                if (det > -Eps && det < Eps) continue;
                float sx_ = r.o.x - ax, sy_ = r.o.y - ay, sz_ = r.o.z - az;
                float uNum = sx_ * px + sy_ * py + sz_ * pz;
                float sgn = det > 0.0f ? 1.0f : -1.0f;
                float detAbs = det * sgn;
                float uNumS = uNum * sgn;
                if (uNumS < 0.0f || uNumS > detAbs) continue;
                float qx = sy_ * e1z - sz_ * e1y;
                float qy = sz_ * e1x - sx_ * e1z;
                float qz = sx_ * e1y - sy_ * e1x;
                float vNum = r.d.x * qx + r.d.y * qy + r.d.z * qz;
                float vNumS = vNum * sgn;
                float uvSumS = uNumS + vNumS;
                if (vNumS < 0.0f || uvSumS > detAbs) continue;
                float tNum = e2x * qx + e2y * qy + e2z * qz;
                float tNumS = tNum * sgn;
                float tMinScaled = tMin * detAbs;
                float tMaxScaled = closest * detAbs;
                if (tNumS < tMinScaled || tNumS > tMaxScaled) continue;
                float invDet = 1.0f / det;
                closest = tNum * invDet;
                any = true;
                slotOut = start + i;
            }
        } else {
            const PairNode *np = mesh.nodes + ref;
            float4 q0 = __ldg(&np->q0), q1 = __ldg(&np->q1), q2 = __ldg(&np->q2), q3 = __ldg(&np->q3);
            int l = __float_as_int(q3.x), rr = __float_as_int(q3.y);
            float lNear = 0.0f, rNear = 0.0f;
            bool hitL = false, hitR = false;
            if (l != YCGE_REF_NONE) hitL = box_mesh(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, r.o, inv, sx, sy, sz, tMin, closest, lNear);
            if (rr != YCGE_REF_NONE) hitR = box_mesh(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, r.o, inv, sx, sy, sz, tMin, closest, rNear);
            if (sp + 2 > YCGE_STACK) { cnt.overflow++; continue; }
            if (hitL & hitR) {
                if (lNear < rNear) { st.ref[sp] = rr; st.tn[sp] = rNear; sp++; st.ref[sp] = l; st.tn[sp] = lNear; sp++; }
                else { st.ref[sp] = l; st.tn[sp] = lNear; sp++; st.ref[sp] = rr; st.tn[sp] = rNear; sp++; }
            } else if (hitL) { st.ref[sp] = l; st.tn[sp] = lNear; sp++; }
            else if (hitR) { st.ref[sp] = rr; st.tn[sp] = rNear; sp++; }
        }
    }
    tOut = closest;
    return any;
}

// ------------------------------------------------------------------------------------------------ VolumeGrid.Hit (VolumeGrid.cs:99-231)
__device__ __forceinline__ int morton3(int x, int y, int z) { // :246-252
    return ((x & 1) << 0) | ((y & 1) << 1) | ((z & 1) << 2) | ((x & 2) << 2) | ((y & 2) << 3) | ((z & 2) << 4) | ((x & 4) << 4) | ((y & 4) << 5) | ((z & 4) << 6);
}
__device__ __forceinline__ bool vol_slab(float ro, float rd, float mn, float mx, float &tEnter, float &tExit, int axis, int &enterAxis) { // :334-355
    if (fabsf(rd) < 1e-12f) { if (ro < mn || ro > mx) return false; return true; }
    float inv = 1.0f / rd;
    float t0 = (mn - ro) * inv, t1 = (mx - ro) * inv;
    if (t0 > t1) { float tmp = t0; t0 = t1; t1 = tmp; }
    if (t0 > tEnter) { tEnter = t0; enterAxis = axis; }
    if (t1 < tExit) tExit = t1;
    return tExit >= tEnter;
}
__device__ __forceinline__ double edge_distance(double v, double v0, double v1) { // :291-297
    double a = v - v0, b = v1 - v;
    if (a < 0.0) a = 0.0;
    if (b < 0.0) b = 0.0;
    return a < b ? a : b;
}
template <int MODE>
__device__ bool volume_hit(const DevVolume &g, const RayD &r, float tMin, float tMax, Cnt<MODE> &cnt, Hit &h) {
    float minX = g.min_corner[0], minY = g.min_corner[1], minZ = g.min_corner[2];
    float sizeX = g.voxel_size[0], sizeY = g.voxel_size[1], sizeZ = g.voxel_size[2];
    int nx = g.nx, ny = g.ny, nz = g.nz;
    float maxX = minX + nx * sizeX, maxY = minY + ny * sizeY, maxZ = minZ + nz * sizeZ;
    int enterAxis = -1;
    float tEnter = -YCGE_INF, tExit = YCGE_INF;
    if (!vol_slab(r.o.x, r.d.x, minX, maxX, tEnter, tExit, 0, enterAxis)) return false;
    if (!vol_slab(r.o.y, r.d.y, minY, maxY, tEnter, tExit, 1, enterAxis)) return false;
    if (!vol_slab(r.o.z, r.d.z, minZ, maxZ, tEnter, tExit, 2, enterAxis)) return false;
    if (!(tExit >= MaxF(0.0f, tEnter))) return false;
    float t = tEnter; if (t < tMin) t = tMin; if (t > tMax || t > tExit) return false;
    t += 1e-6f;
    float ox = r.o.x, oy = r.o.y, oz = r.o.z, dx = r.d.x, dy = r.d.y, dz = r.d.z;
    float px = ox + dx * t, py = oy + dy * t, pz = oz + dz * t;
    int ix = (int)floorf((px - minX) / sizeX); if (ix < 0) ix = 0; else if (ix >= nx) ix = nx - 1;
    int iy = (int)floorf((py - minY) / sizeY); if (iy < 0) iy = 0; else if (iy >= ny) iy = ny - 1;
    int iz = (int)floorf((pz - minZ) / sizeZ); if (iz < 0) iz = 0; else if (iz >= nz) iz = nz - 1;
    int stepX = dx > 0.0f ? 1 : dx < 0.0f ? -1 : 0;
    int stepY = dy > 0.0f ? 1 : dy < 0.0f ? -1 : 0;
    int stepZ = dz > 0.0f ? 1 : dz < 0.0f ? -1 : 0;
    float invDx = stepX == 0 ? 0.0f : 1.0f / dx;
    float invDy = stepY == 0 ? 0.0f : 1.0f / dy;
    float invDz = stepZ == 0 ? 0.0f : 1.0f / dz;
    float nextVx = minX + (stepX > 0 ? (ix + 1) * sizeX : ix * sizeX);
    float nextVy = minY + (stepY > 0 ? (iy + 1) * sizeY : iy * sizeY);
    float nextVz = minZ + (stepZ > 0 ? (iz + 1) * sizeZ : iz * sizeZ);
    float tMaxX = stepX == 0 ? YCGE_INF : (nextVx - ox) * invDx;
    float tMaxY = stepY == 0 ? YCGE_INF : (nextVy - oy) * invDy;
    float tMaxZ = stepZ == 0 ? YCGE_INF : (nextVz - oz) * invDz;
    float tDeltaX = stepX == 0 ? YCGE_INF : fabsf(sizeX * invDx);
    float tDeltaY = stepY == 0 ? YCGE_INF : fabsf(sizeY * invDy);
    float tDeltaZ = stepZ == 0 ? YCGE_INF : fabsf(sizeZ * invDz);
    int lastAxis = enterAxis < 0 ? (tMaxX <= tMaxY && tMaxX <= tMaxZ ? 0 : tMaxY <= tMaxZ ? 1 : 2) : enterAxis;
    float wireMax2 = g.wire_max_distance <= 0.0f ? -1.0f : g.wire_max_distance * g.wire_max_distance;
    float dirLen2 = dx * dx + dy * dy + dz * dz;

    // The voxel itself is fetched only in octants (4^3 blocks) the occupancy byte of the brick marks solid: a ray through air
    // takes the same steps and counts the same cells with one cached byte per brick instead of a fetch per cell.
    int curBrick = -1;
    unsigned occ = 0u;
    while (t <= tExit && t <= tMax) {
        if ((unsigned)ix < (unsigned)nx && (unsigned)iy < (unsigned)ny && (unsigned)iz < (unsigned)nz) {
            CNT_INC(cnt, dda);
            int brick = (((iz >> 3) * g.nby) + (iy >> 3)) * g.nbx + (ix >> 3);
            if (brick != curBrick) { curBrick = brick; occ = __ldg(g.occ + brick); }
            int code = 0;
            if ((occ >> ((((iz >> 2) & 1) << 2) | (((iy >> 2) & 1) << 1) | ((ix >> 2) & 1))) & 1u)
                code = __ldg(g.vox + brick * 512 + morton3(ix & 7, iy & 7, iz & 7));
            if (code > 0) {
                int normalAxis = lastAxis; // never < 0 here (lastAxis is resolved above), VolumeGrid.cs:161-166 is dead
                float hitT = MaxF(t, tMin);
                V3 n = normalAxis == 0 ? mk(stepX > 0 ? -1.0f : 1.0f, 0.0f, 0.0f)
                     : normalAxis == 1 ? mk(0.0f, stepY > 0 ? -1.0f : 1.0f, 0.0f)
                                       : mk(0.0f, 0.0f, stepZ > 0 ? -1.0f : 1.0f);
                V3 hp = mk(ox + dx * hitT, oy + dy * hitT, oz + dz * hitT); // Ray.At
                bool wire = false;
                if (g.wireframe && wireMax2 >= 0.0f) {
                    float dist2 = hitT * hitT * dirLen2;
                    if (dist2 <= wireMax2) { // IsWireOnFace :256-283 (float products widened to double, as in C#)
                        double x0 = (double)(minX + ix * sizeX), x1 = x0 + (double)sizeX;
                        double y0 = (double)(minY + iy * sizeY), y1 = y0 + (double)sizeY;
                        double z0 = (double)(minZ + iz * sizeZ), z1 = z0 + (double)sizeZ;
                        if (normalAxis == 0) {
                            double w = (double)(g.wire_width_frac * MinF(sizeY, sizeZ));
                            wire = edge_distance((double)hp.y, y0, y1) <= w || edge_distance((double)hp.z, z0, z1) <= w;
                        } else if (normalAxis == 1) {
                            double w = (double)(g.wire_width_frac * MinF(sizeX, sizeZ));
                            wire = edge_distance((double)hp.x, x0, x1) <= w || edge_distance((double)hp.z, z0, z1) <= w;
                        } else {
                            double w = (double)(g.wire_width_frac * MinF(sizeX, sizeY));
                            wire = edge_distance((double)hp.x, x0, x1) <= w || edge_distance((double)hp.y, y0, y1) <= w;
                        }
                    }
                }
                h.t = hitT; h.P = hp; h.N = n;
                h.sub = ix + nx * (iy + ny * iz);
                h.mat = code - 1;
                h.albedo_ov = wire ? 1 : 0; // the racy "centre block" highlight (:181-186) needs odd W,H; see DESIGN.md
                h.sr_ov = 0; h.refl = 0.0f;
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

// ------------------------------------------------------------------------------------------------ analytic primitives
__device__ __forceinline__ void matfunc(const DevObject &o, V3 P, Hit &h) { // Scenes.cs:408-428 + Surfaces.cs:64-66
    int m = o.mat_a;
    if (o.checker_scale != 0.0f) {
        int cx = (int)floorf(P.x / o.checker_scale);
        int cz = (int)floorf(P.z / o.checker_scale);
        if (((cx + cz) & 1) != 0) m = o.mat_b;
    }
    h.mat = m; h.albedo_ov = 0; h.sr_ov = o.override_sr; h.refl = o.reflectivity;
}
// XYRect / XZRect / YZRect.Hit (Surfaces.cs:184-214, 256-286, 328-358). axis = index of the constant coordinate.
__device__ __forceinline__ bool rect_hit(const DevObject &o, int axis, float a0, float a1, float b0, float b1, float k, const RayD &r, float tMin, float tMax, Hit &h) {
    float dirK = axis == 2 ? r.d.z : axis == 1 ? r.d.y : r.d.x;
    float oK = axis == 2 ? r.o.z : axis == 1 ? r.o.y : r.o.x;
    float adir = fabsf(dirK);
    float safeDir = copysignf(MaxF(adir, 1e-8f), dirK);
    float t = (k - oK) / safeDir;
    float pa, pb;
    if (axis == 2) { pa = r.o.x + t * r.d.x; pb = r.o.y + t * r.d.y; }
    else if (axis == 1) { pa = r.o.x + t * r.d.x; pb = r.o.z + t * r.d.z; }
    else { pa = r.o.y + t * r.d.y; pb = r.o.z + t * r.d.z; }
    bool ok = adir >= 1e-8f;
    ok &= (t >= tMin) & (t <= tMax);
    ok &= (pa >= a0) & (pa <= a1) & (pb >= b0) & (pb <= b1);
    if (!ok) return false;
    float nk = copysignf(1.0f, -dirK);
    h.t = t;
    if (axis == 2) { h.P = mk(pa, pb, k); h.N = mk(0.0f, 0.0f, nk); }
    else if (axis == 1) { h.P = mk(pa, k, pb); h.N = mk(0.0f, nk, 0.0f); }
    else { h.P = mk(k, pa, pb); h.N = mk(nk, 0.0f, 0.0f); }
    matfunc(o, h.P, h);
    h.sub = 0;
    return true;
}

template <int MODE>
__device__ bool object_hit(const DevScene &sc, int objId, const RayD &r, float tMin, float tMax, Stack &st, int sp, Cnt<MODE> &cnt, Hit &h) {
    CNT_INC(cnt, prims);
    const DevObject &o = sc.objects[objId];
    const int kind = o.kind;
    switch (kind) {
        case YCGE_SPHERE: { // BoundedObjects.cs:31-69
            float Cx = o.p[0], Cy = o.p[1], Cz = o.p[2], Radius = o.p[3];
            float ox = r.o.x - Cx, oy = r.o.y - Cy, oz = r.o.z - Cz;
            float dx = r.d.x, dy = r.d.y, dz = r.d.z;
            float a = dx * dx + dy * dy + dz * dz;
            float halfB = ox * dx + oy * dy + oz * dz;
            float c = ox * ox + oy * oy + oz * oz - Radius * Radius;
            float disc = halfB * halfB - a * c;
            if (disc < 0.0f) return false;
            float s = sqrtf(disc);
            float invA = 1.0f / a;
            float t = (-halfB - s) * invA;
            if (t < tMin || t > tMax) {
                t = (-halfB + s) * invA;
                if (t < tMin || t > tMax) return false;
            }
            float px = r.o.x + t * dx, py = r.o.y + t * dy, pz = r.o.z + t * dz;
            float invR = 1.0f / Radius;
            h.t = t; h.P = mk(px, py, pz);
            h.N = mk((px - Cx) * invR, (py - Cy) * invR, (pz - Cz) * invR);
            h.mat = o.mat_a; h.albedo_ov = 0; h.sr_ov = 0; h.refl = 0.0f; h.sub = 0;
            return true;
        }
        case YCGE_PLANE: { // Surfaces.cs:39-71
            float nx = o.p[3], ny = o.p[4], nz = o.p[5];
            float denom = nx * r.d.x + ny * r.d.y + nz * r.d.z;
            const float Eps = 1e-6f;
            if (denom > -Eps && denom < Eps) return false;
            float t = (o.d[0] - (nx * r.o.x + ny * r.o.y + nz * r.o.z)) / denom;
            if (t < tMin || t > tMax) return false;
            h.t = t; h.P = mk(r.o.x + t * r.d.x, r.o.y + t * r.d.y, r.o.z + t * r.d.z);
            h.N = denom < 0.0f ? mk(nx, ny, nz) : mk(-nx, -ny, -nz);
            matfunc(o, h.P, h);
            h.sub = 0;
            return true;
        }
        case YCGE_DISK: { // Surfaces.cs:108-142
            V3 Normal = mk(o.p[3], o.p[4], o.p[5]);
            float denom = dot3(Normal, r.d);
            float adenom = fabsf(denom);
            float safeDenom = copysignf(MaxF(adenom, 1e-8f), denom);
            float t = (o.d[0] - dot3(Normal, r.o)) / safeDenom;
            float px = r.o.x + t * r.d.x, py = r.o.y + t * r.d.y, pz = r.o.z + t * r.d.z;
            float ddx = px - o.p[0], ddz = pz - o.p[2];
            float rr = ddx * ddx + ddz * ddz;
            bool ok = adenom >= 1e-6f;
            ok &= (t >= tMin) & (t <= tMax);
            ok &= rr <= o.d[1];
            if (!ok) return false;
            h.t = t; h.P = mk(px, py, pz);
            h.N = denom < 0.0f ? Normal : -Normal;
            matfunc(o, h.P, h);
            h.sub = 0;
            return true;
        }
        case YCGE_XYRECT: return rect_hit(o, 2, o.p[0], o.p[1], o.p[2], o.p[3], o.p[4], r, tMin, tMax, h);
        case YCGE_XZRECT: return rect_hit(o, 1, o.p[0], o.p[1], o.p[2], o.p[3], o.p[4], r, tMin, tMax, h);
        case YCGE_YZRECT: return rect_hit(o, 0, o.p[0], o.p[1], o.p[2], o.p[3], o.p[4], r, tMin, tMax, h);
        case YCGE_BOX: { // BoundedObjects.cs:78-115: faces +Z,-Z (XY), +Y,-Y (XZ), +X,-X (YZ) with shrinking closest
            float mnx = o.p[0], mny = o.p[1], mnz = o.p[2], mxx = o.p[3], mxy = o.p[4], mxz = o.p[5];
            bool any = false;
            float closest = tMax;
            Hit tmp;
            if (rect_hit(o, 2, mnx, mxx, mny, mxy, mxz, r, tMin, closest, tmp)) { any = true; closest = tmp.t; h = tmp; h.sub = 0; }
            if (rect_hit(o, 2, mnx, mxx, mny, mxy, mnz, r, tMin, closest, tmp)) { any = true; closest = tmp.t; h = tmp; h.sub = 1; }
            if (rect_hit(o, 1, mnx, mxx, mnz, mxz, mxy, r, tMin, closest, tmp)) { any = true; closest = tmp.t; h = tmp; h.sub = 2; }
            if (rect_hit(o, 1, mnx, mxx, mnz, mxz, mny, r, tMin, closest, tmp)) { any = true; closest = tmp.t; h = tmp; h.sub = 3; }
            if (rect_hit(o, 0, mny, mxy, mnz, mxz, mxx, r, tMin, closest, tmp)) { any = true; closest = tmp.t; h = tmp; h.sub = 4; }
            if (rect_hit(o, 0, mny, mxy, mnz, mxz, mnx, r, tMin, closest, tmp)) { any = true; closest = tmp.t; h = tmp; h.sub = 5; }
            return any;
        }
        case YCGE_CYLINDER_Y: { // BoundedObjects.cs:148-247
            float Cx = o.p[0], Cz = o.p[2], Radius = o.p[3], YMin = o.p[4], YMax = o.p[5];
            bool Capped = o.p[6] != 0.0f;
            float radius2 = o.d[0];
            float ox = r.o.x - Cx, oy = r.o.y, oz = r.o.z - Cz;
            float dx = r.d.x, dy = r.d.y, dz = r.d.z;
            float a = dx * dx + dz * dz;
            float hitT = YCGE_FLT_MAX;
            V3 hitN = mk(0.0f, 0.0f, 0.0f);
            bool hit = false;
            if (a > 1e-12f) {
                float halfB = ox * dx + oz * dz;
                float c = ox * ox + oz * oz - radius2;
                float disc = halfB * halfB - a * c;
                if (disc >= 0.0f) {
                    float s = sqrtf(disc);
                    float invA = 1.0f / a;
                    float t1 = (-halfB - s) * invA;
                    if (t1 > tMin && t1 < tMax) {
                        float y1 = oy + t1 * dy;
                        if (y1 >= YMin && y1 <= YMax) {
                            hitT = t1;
                            hitN = mk((ox + t1 * dx) / Radius, 0.0f, (oz + t1 * dz) / Radius);
                            hit = true;
                        }
                    }
                    if (!hit) {
                        float t2 = (-halfB + s) * invA;
                        if (t2 > tMin && t2 < tMax) {
                            float y2 = oy + t2 * dy;
                            if (y2 >= YMin && y2 <= YMax) {
                                hitT = t2;
                                hitN = mk((ox + t2 * dx) / Radius, 0.0f, (oz + t2 * dz) / Radius);
                                hit = true;
                            }
                        }
                    }
                }
            }
            if (Capped && fabsf(dy) > 1e-8f) {
                float tTop = (YMax - oy) / dy;
                if (tTop > tMin && tTop < tMax) {
                    float rx = ox + tTop * dx, rz = oz + tTop * dz;
                    if (rx * rx + rz * rz <= radius2) { if (tTop < hitT) { hitT = tTop; hitN = mk(0.0f, 1.0f, 0.0f); hit = true; } }
                }
                float tBot = (YMin - oy) / dy;
                if (tBot > tMin && tBot < tMax) {
                    float rx = ox + tBot * dx, rz = oz + tBot * dz;
                    if (rx * rx + rz * rz <= radius2) { if (tBot < hitT) { hitT = tBot; hitN = mk(0.0f, -1.0f, 0.0f); hit = true; } }
                }
            }
            if (!hit) return false;
            h.t = hitT; h.P = mk(r.o.x + hitT * dx, r.o.y + hitT * dy, r.o.z + hitT * dz);
            h.N = dot3(hitN, r.d) < 0.0f ? hitN : -hitN;
            h.mat = o.mat_a; h.albedo_ov = 0; h.sr_ov = 0; h.refl = 0.0f; h.sub = 0;
            return true;
        }
        case YCGE_TRIANGLE: { // Triangle.cs:69-128, the SSE4.1 path: DPPS 0x71 sums (x*x' + y*y') + (z*z' + 0)
            float e1x = o.d[0], e1y = o.d[1], e1z = o.d[2], e2x = o.d[3], e2y = o.d[4], e2z = o.d[5];
            float Dx = r.d.x, Dy = r.d.y, Dz = r.d.z;
            float Sx = r.o.x - o.p[0], Sy = r.o.y - o.p[1], Sz = r.o.z - o.p[2];
            float hx = Dy * e2z - e2y * Dz, hy = Dz * e2x - e2z * Dx, hz = Dx * e2y - e2x * Dy;
            float det = (e1x * hx + e1y * hy) + (e1z * hz + 0.0f);
            if (fabsf(det) < 1e-8f) return false;
            float invDet = 1.0f / det;
            float u = ((Sx * hx + Sy * hy) + (Sz * hz + 0.0f)) * invDet;
            if (u < 0.0f || u > 1.0f) return false;
            float qx = Sy * e1z - e1y * Sz, qy = Sz * e1x - e1z * Sx, qz = Sx * e1y - e1x * Sy;
            float v = ((Dx * qx + Dy * qy) + (Dz * qz + 0.0f)) * invDet;
            if (v < 0.0f || (u + v) > 1.0f) return false;
            float t = ((e2x * qx + e2y * qy) + (e2z * qz + 0.0f)) * invDet;
            if (t < tMin || t > tMax) return false;
            float nx = o.d[6], ny = o.d[7], nz = o.d[8];
            h.t = t; h.P = mk(r.o.x + t * Dx, r.o.y + t * Dy, r.o.z + t * Dz);
            float nd = nx * Dx + ny * Dy + nz * Dz;
            h.N = nd < 0.0f ? mk(nx, ny, nz) : mk(-nx, -ny, -nz);
            h.mat = o.mat_a; h.albedo_ov = 0; h.sr_ov = 0; h.refl = 0.0f; h.sub = 0;
            return true;
        }
        case YCGE_MESH: { // Mesh.cs:26-29 -> MeshBVH.Hit
            const DevMesh &mesh = sc.meshes[o.ref];
            float t; int slot;
            if (!mesh_hit<MODE>(mesh, r, tMin, tMax, st, sp, cnt, t, slot)) return false;
            float4 t2 = __ldg(&mesh.tris[slot].t2);
            float nx = t2.y, ny = t2.z, nz = t2.w;
            h.t = t; h.P = mk(r.o.x + t * r.d.x, r.o.y + t * r.d.y, r.o.z + t * r.d.z);
            float nd = nx * r.d.x + ny * r.d.y + nz * r.d.z;
            h.N = nd < 0.0f ? mk(nx, ny, nz) : mk(-nx, -ny, -nz);
            h.mat = mesh.material; h.albedo_ov = 0; h.sr_ov = 0; h.refl = 0.0f;
            h.sub = slot; // leaf slot; the MeshLoader face index (tri_id[slot]) is looked up by whoever needs it (mesh_face_id)
            return true;
        }
        case YCGE_VOLUME: if (YCGE_LEAN) return false; else return volume_hit<MODE>(sc.volumes[o.ref], r, tMin, tMax, cnt, h);
    }
    return false;
}

// ------------------------------------------------------------------------------------------------ Scene.Hit -> BVH.Hit (BVH.cs:99-198)
template <int MODE>
__device__ bool scene_hit(const DevScene &sc, const RayD &r, float tMin, float tMax, Stack &st, Cnt<MODE> &cnt, Hit &best) {
    cnt.rays++;
    if (sc.root.ref == YCGE_REF_NONE) return false;
    V3 inv = mk(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    float closest = tMax;
    bool any = false;
    int sp = 0;
    {
        float tn;
        CNT_INC(cnt, top_nodes);
        if (!box_top(sc.root.lo[0], sc.root.lo[1], sc.root.lo[2], sc.root.hi[0], sc.root.hi[1], sc.root.hi[2], r.o, inv, tMin, closest, tn)) return false;
        st.ref[0] = sc.root.ref; st.tn[0] = -YCGE_INF; sp = 1;
    }
    bool first = true;
    while (sp > 0) {
        sp--;
        int ref = st.ref[sp];
        float tn = st.tn[sp];
        if (!first) { CNT_INC(cnt, top_nodes); if (!(closest >= tn)) continue; }
        first = false;
        if (ref < 0) {
            int v = ~ref;
            int count = (v >> 26) + 1, start = v & YCGE_LEAF_MAX_START;
            for (int i = 0; i < count; i++) {
                CNT_INC(cnt, leaf_refs);
                int objId = __ldg(sc.leaf_obj + start + i);
                Hit tmp;
                if (object_hit<MODE>(sc, objId, r, tMin, closest, st, sp, cnt, tmp)) {
                    any = true; closest = tmp.t; best = tmp; best.obj = objId;
                }
            }
        } else {
            const PairNode *np = sc.nodes + ref;
            float4 q0 = __ldg(&np->q0), q1 = __ldg(&np->q1), q2 = __ldg(&np->q2), q3 = __ldg(&np->q3);
            int l = __float_as_int(q3.x), rr = __float_as_int(q3.y);
            float lNear = 0.0f, rNear = 0.0f;
            bool hitL = false, hitR = false;
            if (l != YCGE_REF_NONE) hitL = box_top(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, r.o, inv, tMin, closest, lNear);
            if (rr != YCGE_REF_NONE) hitR = box_top(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, r.o, inv, tMin, closest, rNear);
            if (sp + 2 > YCGE_STACK) { cnt.overflow++; continue; }
            if (hitL & hitR) {
                if (lNear < rNear) { st.ref[sp] = rr; st.tn[sp] = rNear; sp++; st.ref[sp] = l; st.tn[sp] = lNear; sp++; }
                else { st.ref[sp] = l; st.tn[sp] = lNear; sp++; st.ref[sp] = rr; st.tn[sp] = rNear; sp++; }
            } else if (hitL) { st.ref[sp] = l; st.tn[sp] = lNear; sp++; }
            else if (hitR) { st.ref[sp] = rr; st.tn[sp] = rNear; sp++; }
        }
    }
    return any;
}

// ------------------------------------------------------------------------------------------------ shading helpers (RaytraceRenderer.cs:737-831)
__device__ __forceinline__ V3 reflect3(V3 v, V3 n) { return v - n * (2.0f * dot3(v, n)); }
__device__ __forceinline__ bool refract3(V3 v, V3 n, float eta, V3 &out) {
    float cosi = -MaxF(-1.0f, MinF(1.0f, dot3(v, n)));
    float k = 1.0f - eta * eta * (1.0f - cosi * cosi);
    if (k < 0.0f) { out = mk(0.0f, 0.0f, 0.0f); return false; }
    out = (v * eta) + (n * (eta * cosi - sqrtf(k)));
    return true;
}
__device__ __forceinline__ float fresnel_schlick(float cosTheta, float etaI, float etaT) {
    float r0 = (etaI - etaT) / (etaI + etaT);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * ycge_powf(1.0f - cosTheta, 5.0f);
}
__device__ V3 oren_nayar(V3 albedo, V3 n, V3 wo, V3 wi, float sigmaRad) {
    const float Pi = 3.14159265358979323846f, InvPi = 1.0f / Pi;
    float cosThetaI = MaxF(0.0f, dot3(n, wi));
    float cosThetaO = MaxF(0.0f, dot3(n, wo));
    if (cosThetaI <= 0.0f || cosThetaO <= 0.0f) return mk(0.0f, 0.0f, 0.0f);
    float sinThetaI = sqrtf(MaxF(0.0f, 1.0f - cosThetaI * cosThetaI));
    float sinThetaO = sqrtf(MaxF(0.0f, 1.0f - cosThetaO * cosThetaO));
    V3 projI = normalized(wi - n * cosThetaI);
    V3 projO = normalized(wo - n * cosThetaO);
    float cosPhiDiff = MaxF(0.0f, dot3(projI, projO));
    float sigma2 = sigmaRad * sigmaRad;
    float A = 1.0f - (sigma2 / (2.0f * (sigma2 + 0.33f)));
    float B = 0.45f * sigma2 / (sigma2 + 0.09f);
    float sinAlpha = MaxF(sinThetaI, sinThetaO);
    float tanBeta = MinF(sinThetaI / MaxF(1e-6f, cosThetaI), sinThetaO / MaxF(1e-6f, cosThetaO));
    float on = (A + B * cosPhiDiff * sinAlpha * tanBeta);
    return saturate3(albedo * (on * InvPi));
}
__device__ V3 cosine_sample_hemisphere(V3 w, unsigned long long &rng) { // RaytraceSampler.cs:83-111
    float u1 = rng_next(rng);
    float u2 = rng_next(rng);
    float r = sqrtf(u1);
    float phi = 6.2831853071795864769f * u2;
    float sn, cs;
    ycge_sincosf(phi, &sn, &cs);
    float x = r * cs, y = r * sn;
    float z = sqrtf(1.0f - u1);
    float wz = w.z;
    if (wz < -0.999999f) {
        V3 u = mk(0.0f, -1.0f, 0.0f), v = mk(-1.0f, 0.0f, 0.0f);
        return u * x + v * y + w * z;
    }
    float a = 1.0f / (1.0f + wz);
    float b = (-w.x * w.y) * a;
    V3 uAxis = mk((float)(1.0 - (double)((w.x * w.x) * a)), b, -w.x);
    V3 vAxis = mk(b, (float)(1.0 - (double)((w.y * w.y) * a)), -w.y);
    return uAxis * x + vAxis * y + w * z;
}

template <int MODE>
__device__ V3 transmittance_to_light(const DevScene &sc, const TraceParams &tp, const RayD &shadow, float maxDist, Stack &st, Cnt<MODE> &cnt) { // :757-798
    Hit block;
    if (!YCGE_LEAN && sc.is_volume_scene) { // Scene.Occluded: a full nearest-hit query with tMin 0.001 (Scene.cs:77-82)
        bool blocked = scene_hit<MODE>(sc, shadow, 0.001f, maxDist, st, cnt, block);
        return blocked ? mk(0.0f, 0.0f, 0.0f) : mk(1.0f, 1.0f, 1.0f);
    }
    float transR = 1.0f, transG = 1.0f, transB = 1.0f;
    float tmin = 0.0f + tp.eps;
    int counter = 0;
    const float cutoff = 1e-6f;
    while (counter < tp.max_refractions && scene_hit<MODE>(sc, shadow, tmin, maxDist, st, cnt, block)) {
        counter++;
        Mat bm = load_material(sc, block);
        if (bm.transparency <= 0.0f) return mk(0.0f, 0.0f, 0.0f);
        float trf = bm.transparency;
        transR *= bm.transmission.x * trf; transG *= bm.transmission.y * trf; transB *= bm.transmission.z * trf;
        if (transR <= cutoff && transG <= cutoff && transB <= cutoff) return mk(0.0f, 0.0f, 0.0f);
        float tHit = block.t;
        if (tHit > maxDist) break;
        tmin = tHit + tp.eps;
    }
    return mk(transR, transG, transB);
}

// Hit.sub of a mesh hit is the leaf slot during traversal; the primitive id of SURVEY 8(c) is the MeshLoader face index.
__device__ __forceinline__ int mesh_face_id(const DevScene &sc, const Hit &h) {
    const DevObject &o = sc.objects[h.obj];
    return o.kind == YCGE_MESH ? __ldg(sc.meshes[o.ref].tri_id + h.sub) : h.sub;
}

// HitRecord.U/V of the accepted hit, evaluated only when the material is textured: the same arithmetic on the same inputs
// as the intersection routine that accepted the hit (rects: from the stored hit point; triangles: the winning triangle
// again), so the values are the reference's bit for bit without carrying two more registers through every traversal.
// Everything is passed BY VALUE (registers): a reference to the kernel's Hit / ray / scene would pin them in local memory
// for the whole path loop (measured: +11 % on the untextured dragon frame).
struct TexRefs { const DevObject *objects; const DevMesh *meshes; const float4 *materials; const DevTexture *textures; int n_textures; };
__device__ void hit_uv(const TexRefs &sc, int obj, int sub, V3 P, V3 ro, V3 rd, float &u, float &v) {
    u = 0.0f; v = 0.0f;
    const DevObject &o = sc.objects[obj];
    const float *p = o.p;
    switch (o.kind) {
        case YCGE_XYRECT: u = (P.x - p[0]) * (1.0f / (p[1] - p[0])); v = (P.y - p[2]) * (1.0f / (p[3] - p[2])); break; // Surfaces.cs:211-212
        case YCGE_XZRECT: u = (P.x - p[0]) * (1.0f / (p[1] - p[0])); v = (P.z - p[2]) * (1.0f / (p[3] - p[2])); break; // :283-284
        case YCGE_YZRECT: u = (P.y - p[0]) * (1.0f / (p[1] - p[0])); v = (P.z - p[2]) * (1.0f / (p[3] - p[2])); break; // :355-356
        case YCGE_BOX: { // BoundedObjects.cs:83-88: faces 0,1 = XYRect, 2,3 = XZRect, 4,5 = YZRect over the box's extents
            float mnx = p[0], mny = p[1], mnz = p[2], mxx = p[3], mxy = p[4], mxz = p[5];
            if (sub < 2) { u = (P.x - mnx) * (1.0f / (mxx - mnx)); v = (P.y - mny) * (1.0f / (mxy - mny)); }
            else if (sub < 4) { u = (P.x - mnx) * (1.0f / (mxx - mnx)); v = (P.z - mnz) * (1.0f / (mxz - mnz)); }
            else { u = (P.y - mny) * (1.0f / (mxy - mny)); v = (P.z - mnz) * (1.0f / (mxz - mnz)); }
            break; }
        case YCGE_TRIANGLE: { // Triangle.cs:69-128
            float e1x = o.d[0], e1y = o.d[1], e1z = o.d[2], e2x = o.d[3], e2y = o.d[4], e2z = o.d[5];
            float Dx = rd.x, Dy = rd.y, Dz = rd.z;
            float Sx = ro.x - p[0], Sy = ro.y - p[1], Sz = ro.z - p[2];
            float hx = Dy * e2z - e2y * Dz, hy = Dz * e2x - e2z * Dx, hz = Dx * e2y - e2x * Dy;
            float det = (e1x * hx + e1y * hy) + (e1z * hz + 0.0f);
            float invDet = 1.0f / det;
            u = ((Sx * hx + Sy * hy) + (Sz * hz + 0.0f)) * invDet;
            float qx = Sy * e1z - e1y * Sz, qy = Sz * e1x - e1z * Sx, qz = Sx * e1y - e1x * Sy;
            v = ((Dx * qx + Dy * qy) + (Dz * qz + 0.0f)) * invDet;
            break; }
        case YCGE_MESH: { // MeshBVH.TriHit :239-304 (sub = leaf slot)
            const DevTri *tp = sc.meshes[o.ref].tris + sub;
            float4 t0 = __ldg(&tp->t0), t1 = __ldg(&tp->t1), t2 = __ldg(&tp->t2);
            float ax = t0.x, ay = t0.y, az = t0.z, e1x = t0.w, e1y = t1.x, e1z = t1.y, e2x = t1.z, e2y = t1.w, e2z = t2.x;
            float px = rd.y * e2z - rd.z * e2y;
            float py = rd.z * e2x - rd.x * e2z;
            float pz = rd.x * e2y - rd.y * e2x;
            float det = e1x * px + e1y * py + e1z * pz;
            float sx_ = ro.x - ax, sy_ = ro.y - ay, sz_ = ro.z - az;
            float uNum = sx_ * px + sy_ * py + sz_ * pz;
            float qx = sy_ * e1z - sz_ * e1y;
            float qy = sz_ * e1x - sx_ * e1z;
            float qz = sx_ * e1y - sy_ * e1x;
            float vNum = rd.x * qx + rd.y * qy + rd.z * qz;
            float invDet = 1.0f / det;
            u = uNum * invDet; v = vNum * invDet;
            break; }
        default: break; // spheres, planes, disks, cylinders, voxels: U = V = 0
    }
}
// Texture.SampleBilinear (Renderer/Texture.cs:143-162): wrap by fraction, (width - 1) scaling, modulo neighbour
__device__ V3 sample_bilinear(const DevTexture &t, float u, float v) {
    if (t.w <= 0 || t.h <= 0 || t.px == nullptr) return mk(1.0f, 1.0f, 1.0f);
    u = u - floorf(u);
    v = v - floorf(v);
    float fx = u * (float)(t.w - 1), fy = v * (float)(t.h - 1);
    int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    int x1 = (x0 + 1) % t.w, y1 = (y0 + 1) % t.h;
    float tx = fx - (float)x0, ty = fy - (float)y0;
    uchar4 q00 = __ldg(t.px + (y0 * t.w + x0)), q10 = __ldg(t.px + (y0 * t.w + x1)), q01 = __ldg(t.px + (y1 * t.w + x0)), q11 = __ldg(t.px + (y1 * t.w + x1));
    V3 c00 = mk((float)q00.x / 255.0f, (float)q00.y / 255.0f, (float)q00.z / 255.0f), c10 = mk((float)q10.x / 255.0f, (float)q10.y / 255.0f, (float)q10.z / 255.0f);
    V3 c01 = mk((float)q01.x / 255.0f, (float)q01.y / 255.0f, (float)q01.z / 255.0f), c11 = mk((float)q11.x / 255.0f, (float)q11.y / 255.0f, (float)q11.z / 255.0f);
    V3 a = c00 * (1.0f - tx) + c10 * tx;
    V3 b = c01 * (1.0f - tx) + c11 * tx;
    V3 c = a * (1.0f - ty) + b * ty;
    return saturate3(c);
}
// SampleAlbedo (RaytraceRenderer.cs:724-735); `albedo` already carries the wireframe override of voxel hits
__device__ __noinline__ float3 sample_albedo(TexRefs sc, float3 albedo, int mat, int obj, int sub, float3 P, float3 ro, float3 rd) {
    float4 m3 = __ldg(sc.materials + 4 * (size_t)mat + 3); // (specular, texture slot, TextureWeight, UVScale)
    int slot = __float_as_int(m3.y);
    if (slot < 0 || slot >= sc.n_textures || m3.z <= 0.0f) return albedo;
    float u, v;
    hit_uv(sc, obj, sub, mk(P.x, P.y, P.z), mk(ro.x, ro.y, ro.z), mk(rd.x, rd.y, rd.z), u, v);
    float tiles = (float)fmax(1e-6, (double)m3.w);
    V3 tex = sample_bilinear(sc.textures[slot], u * tiles, v * tiles);
    float t = m3.z < 0.0f ? 0.0f : (m3.z > 1.0f ? 1.0f : m3.z);
    V3 outAlbedo = mk(albedo.x, albedo.y, albedo.z) * (1.0f - t) + tex * t;
    outAlbedo = saturate3(outAlbedo);
    return make_float3(outAlbedo.x, outAlbedo.y, outAlbedo.z);
}

struct PathItem { RayD ray; V3 beta; int mirror, diffuse; };
#define YCGE_PATH_STACK 16

// ------------------------------------------------------------------------------------------------ the kernel
// Block = 128 threads = 4 warps; a warp covers an 8x4 pixel tile (coherent primary rays, 128-byte row segments
// on every image write); the block covers 16x8 pixels.
#ifndef YCGE_TRACE_MIN_CTAS
#define YCGE_TRACE_MIN_CTAS 8 // measured on B200 (dragon 1080p): 5 CTAs/SM 1.21 ms, 6: 1.22, 8: 1.19, 10: 1.21 — occupancy is not the limiter
#endif
template <int MODE>
__global__ void __launch_bounds__(128, YCGE_TRACE_MIN_CTAS) trace_kernel(DevScene sc, FrameConsts fc, TraceParams tp, ImagePlanes img, int parity, TraceCounters *counters, TraceTotals *totals) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int py = fc.y0 + blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool active = px < fc.W && py < fc.y1;
    Cnt<MODE> cnt;

    if (active) {
        Stack st;
        // ---- MakeJitteredRay :419-437 (camera basis, tan(fov/2) and the per-frame rotations hoisted to the host)
        float base = ((float)c_blue_noise[(py & 7) * 8 + (px & 7)] + 0.5f) * (1.0f / 64.0f);
        float jxBase = fracf_(base + fc.rot0);
        float jyBase = fracf_(base + fc.rot1);
        float jx = fracf_(jxBase + fc.jitter_rot_x) - 0.5f;
        float jy = fracf_(jyBase + fc.jitter_rot_y) - 0.5f;
        float u = (((float)px + 0.5f + jx) / (float)fc.W) * 2.0f - 1.0f;
        float v = 1.0f - (((float)py + 0.5f + jy) / (float)fc.H) * 2.0f;
        V3 fwd = mk(fc.fwd[0], fc.fwd[1], fc.fwd[2]), right = mk(fc.right[0], fc.right[1], fc.right[2]), up = mk(fc.up[0], fc.up[1], fc.up[2]);
        V3 dir = normalized(fwd + right * (u * fc.half_w) + up * (v * fc.half_h));
        RayD primary = make_ray(mk(fc.cam[0], fc.cam[1], fc.cam[2]), dir); // Ray ctor normalises again (Ray.cs:11)
        const size_t pix = (size_t)px + (size_t)py * fc.W;
        if (img.rays) {
            float *rr = img.rays + 6 * pix;
            rr[0] = primary.o.x; rr[1] = primary.o.y; rr[2] = primary.o.z; rr[3] = primary.d.x; rr[4] = primary.d.y; rr[5] = primary.d.z;
        }

        unsigned long long rng = per_frame_seed(px, py, fc.frame, tp.seed_salt);
        if (rng == 0ULL) rng = 0x9E3779B97F4A7C15ULL; // Rng ctor :41-44

        // ---- TraceFull :448-620
        PathItem stack[YCGE_PATH_STACK];
        int sp = 0;
        V3 radiance = mk(0.0f, 0.0f, 0.0f);
        bool primaryHit = false, isSky = false, gbufValid = false;
        V3 gAlb = mk(0.0f, 0.0f, 0.0f), gN = mk(0.0f, 0.0f, 0.0f);
        float gDepth = YCGE_FLT_MAX;
        int gObj = -1, gSub = -1;

        RayD cur = primary;
        V3 beta = mk(1.0f, 1.0f, 1.0f);
        int mirrorDepth = 0, diffuseDepth = 0;
        bool itemPrimary = true;
        bool havePath = true;
        while (havePath) {
            for (;;) {
                Hit rec;
                if (!scene_hit<MODE>(sc, cur, 0.001f, YCGE_FLT_MAX, st, cnt, rec)) {
                    float tbg = 0.5f * (cur.d.y + 1.0f);
                    V3 bb = mk(sc.bg_bottom[0], sc.bg_bottom[1], sc.bg_bottom[2]), bt = mk(sc.bg_top[0], sc.bg_top[1], sc.bg_top[2]);
                    V3 sky = bb * (1.0f - tbg) + bt * tbg;
                    if (itemPrimary && !primaryHit) { isSky = true; if (!gbufValid) gbufValid = true; }
                    radiance = radiance + mk(beta.x * sky.x, beta.y * sky.y, beta.z * sky.z);
                    break;
                }
                Mat m = load_material(sc, rec);
                if (!YCGE_LEAN && sc.n_textures > 0) { // :494,:505 (both calls see the same hit)
                    TexRefs tr = {sc.objects, sc.meshes, sc.materials, sc.textures, sc.n_textures};
                    float3 al = sample_albedo(tr, make_float3(m.albedo.x, m.albedo.y, m.albedo.z), rec.mat, rec.obj, rec.sub, make_float3(rec.P.x, rec.P.y, rec.P.z),
                                              make_float3(cur.o.x, cur.o.y, cur.o.z), make_float3(cur.d.x, cur.d.y, cur.d.z));
                    m.albedo = mk(al.x, al.y, al.z);
                }
                if (itemPrimary) {
                    primaryHit = true; isSky = false;
                    if (!gbufValid) { gAlb = m.albedo; gN = rec.N; gDepth = rec.t; gObj = rec.obj; gSub = mesh_face_id(sc, rec); gbufValid = true; }
                    itemPrimary = false;
                }
                if (m.emission.x != 0.0f || m.emission.y != 0.0f || m.emission.z != 0.0f)
                    radiance = radiance + mk(beta.x * m.emission.x, beta.y * m.emission.y, beta.z * m.emission.z);
                V3 baseAlbedo = m.albedo;
                if (!YCGE_LEAN && m.transparency > 0.0f) {
                    if (mirrorDepth >= tp.max_mirror_bounces) break;
                    V3 n = rec.N, wo = cur.d;
                    bool frontFace = dot3(n, wo) < 0.0f;
                    V3 nl = frontFace ? n : n * -1.0f;
                    float etaI = frontFace ? 1.0f : m.ior;
                    float etaT = frontFace ? m.ior : 1.0f;
                    float eta = etaI / etaT;
                    V3 reflDir = normalized(reflect3(wo, nl));
                    V3 refrDir;
                    bool hasRefract = refract3(wo, nl, eta, refrDir);
                    float cosTheta = fabsf(dot3(nl, wo * -1.0f));
                    float R = fresnel_schlick(cosTheta, etaI, etaT);
                    float Tr = m.transparency < 0.0f ? 0.0f : (m.transparency > 1.0f ? 1.0f : m.transparency);
                    float T = hasRefract ? (1.0f - R) * Tr : 0.0f;
                    { float vv = R + m.reflectivity * (1.0f - R); R = vv < 0.0f ? 0.0f : (vv > 1.0f ? 1.0f : vv); }
                    if (R > 0.0f && sp < YCGE_PATH_STACK) {
                        PathItem it;
                        it.ray = make_ray(rec.P + nl * tp.eps, reflDir);
                        it.beta = mk(beta.x * baseAlbedo.x * R, beta.y * baseAlbedo.y * R, beta.z * baseAlbedo.z * R);
                        it.mirror = mirrorDepth + 1; it.diffuse = diffuseDepth;
                        stack[sp++] = it;
                    }
                    if (T > 0.0f && sp < YCGE_PATH_STACK) {
                        PathItem it;
                        it.ray = make_ray(rec.P - nl * tp.eps, normalized(refrDir));
                        it.beta = mk(beta.x * m.transmission.x * T, beta.y * m.transmission.y * T, beta.z * m.transmission.z * T);
                        it.mirror = mirrorDepth + 1; it.diffuse = diffuseDepth;
                        stack[sp++] = it;
                    }
                    break;
                }
                if (m.reflectivity >= tp.mirror_threshold) {
                    if (mirrorDepth >= tp.max_mirror_bounces) break;
                    V3 reflDir = normalized(reflect3(cur.d, rec.N));
                    cur = make_ray(rec.P + rec.N * tp.eps, reflDir);
                    beta = mk(beta.x * baseAlbedo.x, beta.y * baseAlbedo.y, beta.z * baseAlbedo.z);
                    mirrorDepth++;
                    continue;
                }
                if (sc.ambient_intensity > 0.0f) {
                    V3 a = mk(sc.ambient[0] * sc.ambient_intensity, sc.ambient[1] * sc.ambient_intensity, sc.ambient[2] * sc.ambient_intensity);
                    V3 amb = mk(a.x * baseAlbedo.x, a.y * baseAlbedo.y, a.z * baseAlbedo.z);
                    radiance = radiance + mk(beta.x * amb.x, beta.y * amb.y, beta.z * amb.z);
                }
                V3 woView = normalized(cur.d * -1.0f);
                for (int i = 0; i < sc.n_lights; i++) {
                    const DevLight &L = sc.lights[i];
                    V3 toL = mk(L.pos[0], L.pos[1], L.pos[2]) - rec.P;
                    float dist2 = dot3(toL, toL);
                    float dist = sqrtf(dist2);
                    V3 ldir = vdiv(toL, dist);
                    float nDotL = MaxF(0.0f, dot3(rec.N, ldir));
                    if (nDotL <= 0.0f) continue;
                    RayD shadow = make_ray(rec.P + rec.N * tp.eps, ldir);
                    V3 trans = transmittance_to_light<MODE>(sc, tp, shadow, dist - tp.eps, st, cnt);
                    if (trans.x <= 1e-6f && trans.y <= 1e-6f && trans.z <= 1e-6f) continue;
                    float atten = L.intensity / dist2;
                    V3 fDiffuse = oren_nayar(baseAlbedo, rec.N, woView, ldir, tp.sigma_rad);
                    V3 Li = mk(L.color[0], L.color[1], L.color[2]) * atten;
                    V3 contrib = (fDiffuse * nDotL) * Li;
                    contrib = mk(contrib.x * trans.x, contrib.y * trans.y, contrib.z * trans.z);
                    radiance = radiance + mk(beta.x * contrib.x, beta.y * contrib.y, beta.z * contrib.z);
                }
                if (diffuseDepth < tp.diffuse_bounces) {
                    V3 bounceDir = cosine_sample_hemisphere(rec.N, rng);
                    V3 fON = oren_nayar(baseAlbedo, rec.N, woView, bounceDir, tp.sigma_rad);
                    const float Pi = 3.14159265358979323846f;
                    V3 mult = mk(fON.x * Pi, fON.y * Pi, fON.z * Pi);
                    cur = make_ray(rec.P + rec.N * tp.eps, bounceDir);
                    beta = mk(beta.x * mult.x, beta.y * mult.y, beta.z * mult.z);
                    diffuseDepth++;
                    continue;
                }
                break;
            }
            if (!YCGE_LEAN && sp > 0) {
                sp--;
                cur = stack[sp].ray; beta = stack[sp].beta; mirrorDepth = stack[sp].mirror; diffuseDepth = stack[sp].diffuse;
                itemPrimary = false;
            } else havePath = false;
        }

        // ---- frame planes (RaytraceRenderer.cs:210-215); the normal is stored normalised because every consumer
        // (TAA :327-328, à-trous :663,684) normalises it before use.
        float luma = 0.2126f * radiance.x + 0.7152f * radiance.y + 0.0722f * radiance.z;
        V3 nn = normalized(gN);
        img.cur[pix] = make_float4(radiance.x, radiance.y, radiance.z, luma);
        img.gnd[parity][pix] = make_float4(nn.x, nn.y, nn.z, gDepth);
        img.gas[parity][pix] = make_float4(gAlb.x, gAlb.y, gAlb.z, isSky ? 1.0f : 0.0f);
        img.prim[pix] = make_int2(gObj, gSub);
    }

    // ---- counters: warp-reduce, one atomic per warp
    unsigned int rays = cnt.rays;
    for (int off = 16; off > 0; off >>= 1) rays += __shfl_down_sync(0xffffffffu, rays, off);
    if (lane == 0 && rays) { atomicAdd(&counters->rays, (unsigned long long)rays); atomicAdd(&totals->rays_total, (unsigned long long)rays); }
    if (cnt.overflow) { atomicAdd(&counters->stack_overflow, (unsigned long long)cnt.overflow); if (tp.host_err) *(volatile int *)tp.host_err = 2; }
    if (MODE & 1) {
        unsigned int vals[6] = {cnt.top_nodes, cnt.mesh_nodes, cnt.leaf_refs, cnt.tris, cnt.prims, cnt.dda};
        unsigned long long *dst[6] = {&counters->top_nodes, &counters->mesh_nodes, &counters->leaf_refs, &counters->tris, &counters->prims, &counters->dda};
        for (int k = 0; k < 6; k++) {
            unsigned int x = vals[k];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
            if (lane == 0 && x) atomicAdd(dst[k], (unsigned long long)x);
        }
    }
}

// ------------------------------------------------------------------------------------------------ RNG known-answer kernel
__global__ void rng_kat_kernel(int which, int n, const int *x, const int *y, const long long *frame, int n_draws, unsigned int *out_bits, unsigned long long *out_seed,
                               unsigned long long salt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (which == 0) {
        unsigned long long s = per_frame_seed(x[i], y[i], frame[i], salt);
        out_seed[i] = s;
        if (s == 0ULL) s = 0x9E3779B97F4A7C15ULL;
        for (int k = 0; k < n_draws; k++) out_bits[(size_t)i * n_draws + k] = __float_as_uint(rng_next(s));
    } else { // ConsoleRayTracing.Rng (Rng.cs:3-29)
        unsigned long long seed = ((unsigned long long)(unsigned int)x[i] << 32) | (unsigned int)y[i];
        out_seed[i] = seed;
        unsigned long long state = rngcs_scramble(seed + 0x9E3779B97F4A7C15ULL);
        for (int k = 0; k < n_draws; k++) {
            state += 0x9E3779B97F4A7C15ULL;
            unsigned long long z = rngcs_scramble(state);
            float f = (float)((double)(z >> 11) * (1.0 / 9007199254740992.0));
            out_bits[(size_t)i * n_draws + k] = __float_as_uint(f);
        }
    }
}

} // namespace ycge
