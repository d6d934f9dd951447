// trace_stream.cuh — K1, ray-stream form of the path tracing megakernel (sm_100a).
//
// trace_kernel (trace.cuh) runs one pixel path per thread from start to end; its lanes sit in three different inlined
// copies of the traversal (path ray, shadow ray of a volume scene, shadow ray of the transmittance loop) that the hardware
// runs one after the other, and a CTA's registers stay allocated until its slowest warp ends.  Here the same per-pixel
// arithmetic is cut at every Scene.Hit / Scene.Occluded call (RaytraceRenderer.cs:471, :763, :773) into a per-lane state
// machine run by persistent warps:
//     loop { idle lanes take pixels of the next 8x4 tile;  ONE scene_hit for every lane that has a ray;  each lane consumes
//            its hit and prepares its next ray (path segment or shadow ray) or finishes its pixel }
// Every pixel still performs exactly the reference's operations in the reference's order (its state lives in the lane
// between rays), so the output planes and the traversal event counters are bit-identical to trace_kernel's
// (tests/test_gpu_parity.py::test_trace_kernel_forms_are_bit_identical runs both).
//
// Measured on the B200 (dragon-standin 1080p; trace_kernel 1.18 ms).  `refill_min` = number of idle lanes at which a warp
// hands out new pixels: 1 (every finished lane is refilled at once, the textbook persistent-threads scheme) 1.51 ms,
// 8: 1.38, 16: 1.25, 24: 1.23, 32 (a warp takes a new tile only when all of its pixels are done) 1.09 ms.  Lane occupancy
// is NOT what bounds this kernel: mixing pixels of different tiles — and with them rays of different kinds and regions —
// in one warp costs more in incoherent node / triangle fetches and in rounds that are as long as their longest ray than
// the idle lanes it fills.  The default is therefore 32; what the stream form gains over trace_kernel is the single
// traversal site and the persistent launch.  Registers: 64 (8 CTAs/SM) 1.09 ms, 80 (6 CTAs/SM) 1.21 ms, 51 (10) 1.12 ms.
#pragma once
#include "trace.cuh"

namespace ycge {

enum { YCGE_PH_IDLE = 0, YCGE_PH_PATH = 1, YCGE_PH_SHADOW = 2 };
enum { YCGE_ACT_TRACE = 0, YCGE_ACT_LIGHT_DONE = 1, YCGE_ACT_NEXT_LIGHT = 2, YCGE_ACT_ITEM_DONE = 3 };

#ifndef YCGE_STREAM_MIN_CTAS
#define YCGE_STREAM_MIN_CTAS 8
#endif

template <int MODE>
__global__ void __launch_bounds__(128, YCGE_STREAM_MIN_CTAS) trace_stream_kernel(DevScene sc, FrameConsts fc, TraceParams tp, ImagePlanes img, int parity, TraceCounters *counters,
                                                                                 TraceTotals *totals, int refill_min) {
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int tiles_x = (fc.W + 7) >> 3;
    const unsigned int n_tiles = (unsigned int)tiles_x * (unsigned int)((fc.y1 - fc.y0 + 3) >> 2);
    unsigned int *next_tile = (unsigned int *)&counters->next_tile;
    Cnt<MODE> cnt;
    Stack st;
    PathItem stack[YCGE_PATH_STACK];

    // the warp's pixel pool: tile `pool_tile`, pixels [pool_next, 32) not handed out yet (warp-uniform)
    unsigned int pool_tile = 0;
    int pool_next = 32;
    bool exhausted = false;

    // ---- per-lane state: the locals of TraceFull (:448-620) and of ComputeTransmittanceToLight (:757-798)
    int phase = YCGE_PH_IDLE;
    int px = 0, py = 0;
    unsigned long long rng = 0;
    int sp = 0;
    V3 radiance = mk(0.0f, 0.0f, 0.0f);
    bool primaryHit = false, isSky = false, gbufValid = false, itemPrimary = false;
    V3 gAlb = mk(0.0f, 0.0f, 0.0f), gN = mk(0.0f, 0.0f, 0.0f);
    float gDepth = YCGE_FLT_MAX;
    int gObj = -1, gSub = -1;
    V3 beta = mk(1.0f, 1.0f, 1.0f);
    int mirrorDepth = 0, diffuseDepth = 0;
    RayD ray;                        // the ray of this round (path segment = `currentRay`, or the shadow ray)
    ray.o = mk(0.0f, 0.0f, 0.0f); ray.d = mk(0.0f, 0.0f, 1.0f);
    float tMin = 0.001f, tMax = YCGE_FLT_MAX;
    V3 sP = mk(0.0f, 0.0f, 0.0f), sN = mk(0.0f, 0.0f, 0.0f), sAlb = mk(0.0f, 0.0f, 0.0f), woView = mk(0.0f, 0.0f, 0.0f); // rec.P, rec.N, baseAlbedo, woView of the hit being lit
    int li = 0;                      // light index of the loop :578-603
    float transR = 1.0f, transG = 1.0f, transB = 1.0f, maxDist = 0.0f;
    int counter = 0;

    for (;;) {
        // ---- refill idle lanes
        unsigned int idle = __ballot_sync(FULL, phase == YCGE_PH_IDLE);
        while (idle != 0u && !exhausted && (__popc(idle) >= refill_min || idle == FULL)) {
            if (pool_next >= 32) {
                unsigned int t = 0;
                if (lane == 0) t = atomicAdd(next_tile, 1u);
                t = __shfl_sync(FULL, t, 0);
                if (t >= n_tiles) { exhausted = true; break; }
                pool_tile = t; pool_next = 0;
            }
            const int k = pool_next + __popc(idle & ((1u << lane) - 1u));
            if (phase == YCGE_PH_IDLE && k < 32) {
                px = (int)(pool_tile % (unsigned int)tiles_x) * 8 + (k & 7);
                py = fc.y0 + (int)(pool_tile / (unsigned int)tiles_x) * 4 + (k >> 3);
                if (px < fc.W && py < fc.y1) {
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
                    ray = make_ray(mk(fc.cam[0], fc.cam[1], fc.cam[2]), dir); // Ray ctor normalises again (Ray.cs:11)
                    if (img.rays) {
                        float *rr = img.rays + 6 * ((size_t)px + (size_t)py * fc.W);
                        rr[0] = ray.o.x; rr[1] = ray.o.y; rr[2] = ray.o.z; rr[3] = ray.d.x; rr[4] = ray.d.y; rr[5] = ray.d.z;
                    }
                    rng = per_frame_seed(px, py, fc.frame, tp.seed_salt);
                    if (rng == 0ULL) rng = 0x9E3779B97F4A7C15ULL; // Rng ctor :41-44
                    sp = 0;
                    radiance = mk(0.0f, 0.0f, 0.0f);
                    primaryHit = false; isSky = false; gbufValid = false; itemPrimary = true;
                    gAlb = mk(0.0f, 0.0f, 0.0f); gN = mk(0.0f, 0.0f, 0.0f); gDepth = YCGE_FLT_MAX; gObj = -1; gSub = -1;
                    beta = mk(1.0f, 1.0f, 1.0f);
                    mirrorDepth = 0; diffuseDepth = 0;
                    tMin = 0.001f; tMax = YCGE_FLT_MAX;
                    phase = YCGE_PH_PATH;
                }
            }
            pool_next += min(__popc(idle), 32 - pool_next);
            idle = __ballot_sync(FULL, phase == YCGE_PH_IDLE);
        }
        if (__ballot_sync(FULL, phase != YCGE_PH_IDLE) == 0u) break;

        // ---- one Scene.Hit / Scene.Occluded for every lane that carries a ray
        Hit rec;
        bool hit = false;
        if (phase != YCGE_PH_IDLE) hit = scene_hit<MODE>(sc, ray, tMin, tMax, st, cnt, rec);
        if (phase == YCGE_PH_IDLE) continue;

        int act = YCGE_ACT_TRACE;
        if (phase == YCGE_PH_PATH) { // the body of TraceFull's inner loop :471-616
            if (!hit) {
                float tbg = 0.5f * (ray.d.y + 1.0f);
                V3 bb = mk(sc.bg_bottom[0], sc.bg_bottom[1], sc.bg_bottom[2]), bt = mk(sc.bg_top[0], sc.bg_top[1], sc.bg_top[2]);
                V3 sky = bb * (1.0f - tbg) + bt * tbg;
                if (itemPrimary && !primaryHit) { isSky = true; if (!gbufValid) gbufValid = true; }
                radiance = radiance + mk(beta.x * sky.x, beta.y * sky.y, beta.z * sky.z);
                act = YCGE_ACT_ITEM_DONE;
            } else {
                Mat m = load_material(sc, rec);
                if (!YCGE_LEAN && sc.n_textures > 0) { // :494,:505 (both calls see the same hit)
                    TexRefs tr = {sc.objects, sc.meshes, sc.materials, sc.textures, sc.n_textures};
                    float3 al = sample_albedo(tr, make_float3(m.albedo.x, m.albedo.y, m.albedo.z), rec.mat, rec.obj, rec.sub, make_float3(rec.P.x, rec.P.y, rec.P.z),
                                              make_float3(ray.o.x, ray.o.y, ray.o.z), make_float3(ray.d.x, ray.d.y, ray.d.z));
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
                    if (mirrorDepth < tp.max_mirror_bounces) {
                        V3 n = rec.N, wo = ray.d;
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
                    }
                    act = YCGE_ACT_ITEM_DONE;
                } else if (m.reflectivity >= tp.mirror_threshold) {
                    if (mirrorDepth >= tp.max_mirror_bounces) act = YCGE_ACT_ITEM_DONE;
                    else {
                        V3 reflDir = normalized(reflect3(ray.d, rec.N));
                        ray = make_ray(rec.P + rec.N * tp.eps, reflDir);
                        beta = mk(beta.x * baseAlbedo.x, beta.y * baseAlbedo.y, beta.z * baseAlbedo.z);
                        mirrorDepth++;
                        tMin = 0.001f; tMax = YCGE_FLT_MAX;
                    }
                } else {
                    if (sc.ambient_intensity > 0.0f) {
                        V3 a = mk(sc.ambient[0] * sc.ambient_intensity, sc.ambient[1] * sc.ambient_intensity, sc.ambient[2] * sc.ambient_intensity);
                        V3 amb = mk(a.x * baseAlbedo.x, a.y * baseAlbedo.y, a.z * baseAlbedo.z);
                        radiance = radiance + mk(beta.x * amb.x, beta.y * amb.y, beta.z * amb.z);
                    }
                    woView = normalized(ray.d * -1.0f);
                    sP = rec.P; sN = rec.N; sAlb = baseAlbedo;
                    li = 0;
                    act = YCGE_ACT_NEXT_LIGHT;
                }
            }
        } else { // the shadow ray of light `li` came back: one turn of ComputeTransmittanceToLight's loop :757-798
            if (!YCGE_LEAN && sc.is_volume_scene) { // Scene.Occluded: a full nearest-hit query with tMin 0.001 (Scene.cs:77-82)
                float tv = hit ? 0.0f : 1.0f;
                transR = tv; transG = tv; transB = tv;
                act = YCGE_ACT_LIGHT_DONE;
            } else if (!hit) act = YCGE_ACT_LIGHT_DONE;
            else {
                counter++;
                Mat bm = load_material(sc, rec);
                const float cutoff = 1e-6f;
                if (bm.transparency <= 0.0f) { transR = 0.0f; transG = 0.0f; transB = 0.0f; act = YCGE_ACT_LIGHT_DONE; }
                else {
                    float trf = bm.transparency;
                    transR *= bm.transmission.x * trf; transG *= bm.transmission.y * trf; transB *= bm.transmission.z * trf;
                    if (transR <= cutoff && transG <= cutoff && transB <= cutoff) { transR = 0.0f; transG = 0.0f; transB = 0.0f; act = YCGE_ACT_LIGHT_DONE; }
                    else if (rec.t > maxDist) act = YCGE_ACT_LIGHT_DONE;
                    else {
                        tMin = rec.t + tp.eps;
                        act = counter < tp.max_refractions ? YCGE_ACT_TRACE : YCGE_ACT_LIGHT_DONE;
                    }
                }
            }
        }

        // ---- the light loop :578-603 and the diffuse bounce :604-615, resumed where the lane left it
        while (act == YCGE_ACT_LIGHT_DONE || act == YCGE_ACT_NEXT_LIGHT) {
            if (act == YCGE_ACT_LIGHT_DONE) {
                if (!(transR <= 1e-6f && transG <= 1e-6f && transB <= 1e-6f)) {
                    const DevLight &L = sc.lights[li];
                    V3 toL = mk(L.pos[0], L.pos[1], L.pos[2]) - sP;
                    float dist2 = dot3(toL, toL);
                    float dist = sqrtf(dist2);
                    V3 ldir = vdiv(toL, dist);
                    float nDotL = MaxF(0.0f, dot3(sN, ldir));
                    float atten = L.intensity / dist2;
                    V3 fDiffuse = oren_nayar(sAlb, sN, woView, ldir, tp.sigma_rad);
                    V3 Li = mk(L.color[0], L.color[1], L.color[2]) * atten;
                    V3 contrib = (fDiffuse * nDotL) * Li;
                    contrib = mk(contrib.x * transR, contrib.y * transG, contrib.z * transB);
                    radiance = radiance + mk(beta.x * contrib.x, beta.y * contrib.y, beta.z * contrib.z);
                }
                li++;
                act = YCGE_ACT_NEXT_LIGHT;
            }
            if (li >= sc.n_lights) {
                if (diffuseDepth < tp.diffuse_bounces) {
                    V3 bounceDir = cosine_sample_hemisphere(sN, rng);
                    V3 fON = oren_nayar(sAlb, sN, woView, bounceDir, tp.sigma_rad);
                    const float Pi = 3.14159265358979323846f;
                    V3 mult = mk(fON.x * Pi, fON.y * Pi, fON.z * Pi);
                    ray = make_ray(sP + sN * tp.eps, bounceDir);
                    beta = mk(beta.x * mult.x, beta.y * mult.y, beta.z * mult.z);
                    diffuseDepth++;
                    tMin = 0.001f; tMax = YCGE_FLT_MAX;
                    phase = YCGE_PH_PATH;
                    act = YCGE_ACT_TRACE;
                } else act = YCGE_ACT_ITEM_DONE;
                break;
            }
            const DevLight &L = sc.lights[li];
            V3 toL = mk(L.pos[0], L.pos[1], L.pos[2]) - sP;
            float dist2 = dot3(toL, toL);
            float dist = sqrtf(dist2);
            V3 ldir = vdiv(toL, dist);
            float nDotL = MaxF(0.0f, dot3(sN, ldir));
            if (nDotL <= 0.0f) { li++; continue; }
            ray = make_ray(sP + sN * tp.eps, ldir);
            maxDist = dist - tp.eps;
            transR = 1.0f; transG = 1.0f; transB = 1.0f;
            counter = 0;
            tMax = maxDist;
            if (!YCGE_LEAN && sc.is_volume_scene) { tMin = 0.001f; phase = YCGE_PH_SHADOW; act = YCGE_ACT_TRACE; }
            else if (counter < tp.max_refractions) { tMin = 0.0f + tp.eps; phase = YCGE_PH_SHADOW; act = YCGE_ACT_TRACE; }
            else act = YCGE_ACT_LIGHT_DONE; // max_refractions == 0: the loop :773 never runs, full transmittance
        }

        if (act == YCGE_ACT_ITEM_DONE) {
            if (!YCGE_LEAN && sp > 0) { // :463-468: next deferred item (reflection / refraction branch)
                sp--;
                ray = stack[sp].ray; beta = stack[sp].beta; mirrorDepth = stack[sp].mirror; diffuseDepth = stack[sp].diffuse;
                itemPrimary = false;
                tMin = 0.001f; tMax = YCGE_FLT_MAX;
                phase = YCGE_PH_PATH;
            } else {
                // ---- frame planes (RaytraceRenderer.cs:210-215); the normal is stored normalised because every consumer
                // (TAA :327-328, à-trous :663,684) normalises it before use.
                const size_t pix = (size_t)px + (size_t)py * fc.W;
                float luma = 0.2126f * radiance.x + 0.7152f * radiance.y + 0.0722f * radiance.z;
                V3 nn = normalized(gN);
                img.cur[pix] = make_float4(radiance.x, radiance.y, radiance.z, luma);
                img.gnd[parity][pix] = make_float4(nn.x, nn.y, nn.z, gDepth);
                img.gas[parity][pix] = make_float4(gAlb.x, gAlb.y, gAlb.z, isSky ? 1.0f : 0.0f);
                img.prim[pix] = make_int2(gObj, gSub);
                phase = YCGE_PH_IDLE;
            }
        }
    }

    // ---- counters: warp-reduce, one atomic per warp
    unsigned int rays = cnt.rays;
    for (int off = 16; off > 0; off >>= 1) rays += __shfl_down_sync(FULL, rays, off);
    if (lane == 0 && rays) { atomicAdd(&counters->rays, (unsigned long long)rays); atomicAdd(&totals->rays_total, (unsigned long long)rays); }
    if (cnt.overflow) { atomicAdd(&counters->stack_overflow, (unsigned long long)cnt.overflow); if (tp.host_err) *(volatile int *)tp.host_err = 2; }
    if (MODE & 1) {
        unsigned int vals[6] = {cnt.top_nodes, cnt.mesh_nodes, cnt.leaf_refs, cnt.tris, cnt.prims, cnt.dda};
        unsigned long long *dst[6] = {&counters->top_nodes, &counters->mesh_nodes, &counters->leaf_refs, &counters->tris, &counters->prims, &counters->dda};
        for (int k = 0; k < 6; k++) {
            unsigned int x = vals[k];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(FULL, x, off);
            if (lane == 0 && x) atomicAdd(dst[k], (unsigned long long)x);
        }
    }
}

} // namespace ycge
