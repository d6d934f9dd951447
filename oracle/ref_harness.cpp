// ref_harness.cpp — C entry points over oracle/_ref/ref_generated.hpp, i.e. over the reference's own C# text rewritten into C++
// by oracle/ref_transpile.py (test infrastructure; see that script).  Nothing here computes: it moves planes in and out.
#include "_ref/ref_generated.hpp"
#include "../include/ycge.h"

#define RH_API extern "C" __attribute__((visibility("default")))
using namespace refcs;

namespace {
struct Handle {
    RendererRef r;
    Fast2D<Chexel> target;
    std::unique_ptr<Framebuffer> fb;
    std::unique_ptr<TemporalAA> taa; // RaytraceRenderer.taa (:56, :96): asked for the reset decision (:171), told the camera (:266)
    int W = 0, H = 0;
};
template <class T> Fast2D<T> plane(int w, int h) { return Fast2D<T>(w, h); }
}

// ---- the tail of TryFlipAndBlit (RaytraceRenderer.cs:218-264): TAA, a-trous, exposure, cells
RH_API void *ref_renderer_create(int fb_w, int fb_h, int ss, int proc_count) {
    auto *h = new Handle();
    h->r.ss = ss; h->r.fbW = fb_w; h->r.fbH = fb_h; h->r.procCount = proc_count;
    h->W = fb_w * ss; h->H = fb_h * 2 * ss; // hiW, hiH (RaytraceRenderer.cs:86-87)
    h->target = Fast2D<Chexel>(fb_w, fb_h);
    h->fb.reset(new Framebuffer(fb_w, fb_h));
    h->taa.reset(new TemporalAA(h->W, h->H, h->r.taaAlpha, RendererRef::MotionTransReset, RendererRef::MotionRotReset)); // :96
    h->r.taa = h->taa.get();
    return h;
}
// RaytraceRenderer.Resize (:110-138) as the reference wrote it: planes and history reallocated, taa.Resize, taaHistoryValid = false;
// the tone mapper (aeExposure) and the frame counter stay.  The trace half of the harness is resized by ref_trace_resize.
RH_API int ref_renderer_resize(void *hh, int fb_w, int fb_h, int ss) {
    Handle &h = *(Handle *)hh;
    try {
        h.fb.reset(new Framebuffer(fb_w, fb_h));
        h.r.Resize(h.fb.get(), ss);
        h.W = h.r.hiW; h.H = h.r.hiH;
        h.target = h.r.frameBuffer; // `var target = frameBuffer;` (:181)
        return 0;
    } catch (...) { return -1; }
}
RH_API void ref_renderer_destroy(void *hh) { delete (Handle *)hh; }
// TryFlipAndBlit :171 (without `|| scene.HasDynamicTextures`) and :266; Resize :128
RH_API int ref_taa_should_reset(void *hh, const float *cam3, float yaw, float pitch) { return ((Handle *)hh)->taa->ShouldResetHistory(Vec3(cam3[0], cam3[1], cam3[2]), yaw, pitch) ? 1 : 0; }
RH_API void ref_taa_commit_camera(void *hh, const float *cam3, float yaw, float pitch) { ((Handle *)hh)->taa->CommitCamera(Vec3(cam3[0], cam3[1], cam3[2]), yaw, pitch); }
RH_API void ref_taa_resize(void *hh) { Handle &h = *(Handle *)hh; h.taa->Resize(h.W, h.H); }
// one frame: the trace stage's planes in (row-major W x H: hdr rgb, albedo rgb, raw normal xyz, depth, sky 0/1), everything after out
RH_API int ref_post_frame(void *hh, const float *hdr3, const float *albedo3, const float *normal3, const float *depth, const uint8_t *sky, int reset_history,
                          float *taa3, float *den3, float *exposure2, uint16_t *glyph, uint8_t *fg16, uint8_t *bg16, uint8_t *fg_ansi, uint8_t *bg_ansi, float *fg3, float *bg3) {
    Handle &h = *(Handle *)hh;
    const int W = h.W, H = h.H;
    try {
        Fast2D<Vec3> cur(W, H);
        h.r.gAlbedo = Fast2D<Vec3>(W, H); h.r.gNormal = Fast2D<Vec3>(W, H); h.r.gDepth = Fast2D<float>(W, H); h.r.skyMask = Fast2D<bool>(W, H);
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            const size_t p = (size_t)x + (size_t)y * W;
            cur[x, y] = Vec3(hdr3[3 * p], hdr3[3 * p + 1], hdr3[3 * p + 2]);
            h.r.gAlbedo[x, y] = Vec3(albedo3[3 * p], albedo3[3 * p + 1], albedo3[3 * p + 2]);
            h.r.gNormal[x, y] = Vec3(normal3[3 * p], normal3[3 * p + 1], normal3[3 * p + 2]);
            h.r.gDepth[x, y] = depth[p];
            h.r.skyMask[x, y] = sky[p] != 0;
        }
        h.r.PostTail(cur, reset_history != 0, h.target, *h.fb);
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            const size_t p = (size_t)x + (size_t)y * W;
            const Vec3 t = h.r.taaHistory[x, y], d = h.r.lastDenoised[x, y];
            taa3[3 * p] = t.X; taa3[3 * p + 1] = t.Y; taa3[3 * p + 2] = t.Z;
            den3[3 * p] = d.X; den3[3 * p + 1] = d.Y; den3[3 * p + 2] = d.Z;
        }
        exposure2[0] = h.r.toneMapper.aeExposure; exposure2[1] = h.r.toneMapper.effectiveExposure;
        for (int cy = 0; cy < h.r.fbH; cy++) for (int cx = 0; cx < h.r.fbW; cx++) {
            const size_t c = (size_t)cx + (size_t)cy * h.r.fbW;
            const Chexel ch = h.target[cx, cy];
            glyph[c] = (uint16_t)ch.Char;
            fg16[c] = (uint8_t)(int)ch.ForegroundColor.color_16; bg16[c] = (uint8_t)(int)ch.BackgroundColor.color_16;
            fg_ansi[c] = (uint8_t)AnsiRef::ChexelToAnsi256(ch.ForegroundColor); bg_ansi[c] = (uint8_t)AnsiRef::ChexelToAnsi256(ch.BackgroundColor);
            fg3[3 * c] = ch.ForegroundColor.color_f32.X; fg3[3 * c + 1] = ch.ForegroundColor.color_f32.Y; fg3[3 * c + 2] = ch.ForegroundColor.color_f32.Z;
            bg3[3 * c] = ch.BackgroundColor.color_f32.X; bg3[3 * c + 1] = ch.BackgroundColor.color_f32.Y; bg3[3 * c + 2] = ch.BackgroundColor.color_f32.Z;
        }
        return 0;
    } catch (...) { return -1; }
}

// ---- RaytraceSampler.cs
RH_API uint64_t ref_per_frame_seed(int x, int y, int64_t frame, int jx, int jy, uint64_t salt) { return RaytraceSampler::PerFrameSeed(x, y, frame, jx, jy, salt); }
RH_API uint64_t ref_splitmix64(uint64_t z) { return RaytraceSampler::SplitMix64(z); }
RH_API void ref_rng_draws(uint64_t seed, int n, float *out) { RaytraceSampler::Rng rng(seed); for (int i = 0; i < n; i++) out[i] = rng.NextUnit(); }
RH_API void ref_rng_cs_draws(uint64_t seed, int n, uint32_t *bits_out) { rngcs::Rng g(seed); for (int i = 0; i < n; i++) { float f = g.NextUnit(); std::memcpy(&bits_out[i], &f, 4); } } // Rng.cs
RH_API float ref_blue_noise(int x, int y, int frame_idx, int channel) { return RaytraceSampler::BlueNoiseSample(x, y, frame_idx, channel); }
RH_API void ref_cosine_sample(float nx, float ny, float nz, uint64_t seed, float *out3) {
    RaytraceSampler::Rng rng(seed);
    const Vec3 d = RaytraceSampler::CosineSampleHemisphere(Vec3(nx, ny, nz), rng);
    out3[0] = d.X; out3[1] = d.Y; out3[2] = d.Z;
}
// ---- BSDF helpers of RaytraceRenderer.cs
RH_API float ref_fresnel_schlick(float cos_theta, float eta_i, float eta_t) { return RendererRef::FresnelSchlick(cos_theta, eta_i, eta_t); }
RH_API int ref_refract(const float *v3, const float *n3, float eta, float *out3) {
    Vec3 o;
    const bool ok = RendererRef::Refract(Vec3(v3[0], v3[1], v3[2]), Vec3(n3[0], n3[1], n3[2]), eta, o);
    out3[0] = o.X; out3[1] = o.Y; out3[2] = o.Z;
    return ok ? 1 : 0;
}
RH_API void ref_reflect(const float *v3, const float *n3, float *out3) {
    const Vec3 o = RendererRef::Reflect(Vec3(v3[0], v3[1], v3[2]), Vec3(n3[0], n3[1], n3[2]));
    out3[0] = o.X; out3[1] = o.Y; out3[2] = o.Z;
}
RH_API void ref_oren_nayar(const float *albedo3, const float *n3, const float *wo3, const float *wi3, float sigma, float *out3) {
    const Vec3 o = RendererRef::OrenNayarBRDF(Vec3(albedo3[0], albedo3[1], albedo3[2]), Vec3(n3[0], n3[1], n3[2]), Vec3(wo3[0], wo3[1], wo3[2]), Vec3(wi3[0], wi3[1], wi3[2]), sigma);
    out3[0] = o.X; out3[1] = o.Y; out3[2] = o.Z;
}
// ---- quantisers
RH_API int ref_ansi256(float r, float g, float b) { return AnsiRef::ChexelToAnsi256(ChexelColor(Vec3(r, g, b))); }
RH_API int ref_nearest16(float r, float g, float b) { return (int)ChexelColor(Vec3(r, g, b)).color_16; }
RH_API int ref_linear_to_srgb8(double c) { return AnsiRef::LinearToSrgb8(c); }
// Texture.SampleBilinear (Texture.cs:108-163, static image) and RaytraceRenderer.SampleAlbedo (:724-735) over a w x h image of RGBA words, row 0 first
RH_API void ref_texture_sample(int w, int hh, const uint32_t *rgba, float u, float v, float *out3) {
    Texture t; t.width = w; t.height = hh; t.pixels.assign((const int *)rgba, (const int *)rgba + (size_t)w * hh);
    const Vec3 c = t.SampleBilinear(u, v);
    out3[0] = c.X; out3[1] = c.Y; out3[2] = c.Z;
}
RH_API void ref_sample_albedo(const float *albedo3, double texture_weight, double uv_scale, int w, int hh, const uint32_t *rgba, float u, float v, float *out3) {
    Texture t; t.width = w; t.height = hh; t.pixels.assign((const int *)rgba, (const int *)rgba + (size_t)w * hh);
    Material m; m.Albedo = Vec3(albedo3[0], albedo3[1], albedo3[2]); m.DiffuseTexture = rgba ? &t : nullptr; m.TextureWeight = texture_weight; m.UVScale = uv_scale;
    const Vec3 c = RendererRef::SampleAlbedo(m, Vec3(0.0f, 0.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), u, v);
    out3[0] = c.X; out3[1] = c.Y; out3[2] = c.Z;
}
// The island generator as the reference wrote it: BuildMinecraftLike's WorldConfig (VolumeScenes.cs:579-591: chunks of 32, worldMin (-size/2, 0, -size/2),
// unit voxels, seed 0) for a world of size x height x size, then the three passes of WorldManager.GenerateAndSaveWorld (heights + rivers, voxel fill,
// flora).  ids / metas: [x][y][z], size * height * size ints each.
// path != NULL: the reference's own VG01 writer (WorldManager.cs:607-631) writes the world there as well; ids / metas may be NULL then.
RH_API int ref_generate_island(int world_size, int world_height, int *ids, int *metas, const char *path) {
    try {
        const int chunk = 32;
        WorldConfig cfg(chunk, world_size / chunk, world_height / chunk, world_size / chunk, 8, Vec3(-world_size / 2, 0, -world_size / 2), Vec3(1, 1, 1), 0);
        Array3<Cell2> cells = WorldGenRef::GenerateCells(cfg, path ? String(path) : String());
        const size_t n = (size_t)cells.n0 * cells.n1 * cells.n2;
        if (ids && metas) for (size_t i = 0; i < n; i++) { ids[i] = (*cells.p)[i].Item1; metas[i] = (*cells.p)[i].Item2; }
        return 0;
    } catch (...) { return -1; }
}
RH_API int ref_map_attributes(int fg16, int bg16) { return (int)Win32Ref::MapAttributes((ConsoleColor)fg16, (ConsoleColor)bg16); } // Win32TerminalRenderer.cs:109-112
// ANSITerminalRenderer.Render (:86-153) over one Framebuffer of fb_w x fb_h cells (row-major glyph / fg rgb / bg rgb) placed at (vx, vy) on a
// console of console_w x console_h cells; the renderer believes the console to be known_w x known_h (differs -> the resize prologue).
// Returns the number of bytes the reference would write (or -needed when `cap` is too small).
RH_API int ref_ansi_render(int fb_w, int fb_h, const uint16_t *glyph, const float *fg3, const float *bg3, int vx, int vy, int console_w, int console_h, int known_w, int known_h,
                           uint8_t *out, int cap) {
    try {
        Framebuffer fb(fb_w, fb_h);
        fb.ViewportX = vx; fb.ViewportY = vy;
        for (int y = 0; y < fb_h; y++) for (int x = 0; x < fb_w; x++) {
            const size_t c = (size_t)x + (size_t)y * fb_w;
            fb.SetChexel(x, y, Chexel((char16_t)glyph[c], Vec3(fg3[3 * c], fg3[3 * c + 1], fg3[3 * c + 2]), Vec3(bg3[3 * c], bg3[3 * c + 1], bg3[3 * c + 2])));
        }
        AnsiRenderRef r;
        r.frameBuffers.Add(&fb);
        Console::WindowWidth = console_w; Console::WindowHeight = console_h + 1; // consoleHeight = Console.WindowHeight - 1 (:32, :89)
        r.consoleWidth = known_w; r.consoleHeight = known_h;
        r.Render();
        if ((int)r.flushed.size() > cap) return -(int)r.flushed.size();
        std::memcpy(out, r.flushed.data(), r.flushed.size());
        return (int)r.flushed.size();
    } catch (...) { return -1; }
}

// ---- the analytic primitives: objects built by the reference's own constructors from the flat scene description
// (include/ycge.h: ycge_object.p = the public fields after the ctor ran), rays against the reference's own Hit methods.
// Materials: MS = 16 floats (albedo3, specular, reflectivity, emission3, transparency, ior, transmission3, texture id or -1, texture weight,
// uv scale); a texture id names an image handed over before by ref_set_textures (static images: `new Texture(path)`, Texture.cs:25-49).
namespace {
constexpr int MS = 16;
std::vector<std::unique_ptr<Texture>> g_textures;
Material mat_of(const float *m) {
    Material r(Vec3(m[0], m[1], m[2]), (double)m[3], (double)m[4], Vec3(m[5], m[6], m[7]), (double)m[8], (double)m[9], Vec3(m[10], m[11], m[12]));
    const int tex = (int)m[13];
    if (tex >= 0 && tex < (int)g_textures.size()) { r.DiffuseTexture = g_textures[(size_t)tex].get(); r.TextureWeight = (double)m[14]; r.UVScale = (double)m[15]; }
    return r;
}
Hittable *make_prim(int kind, const float *p, const float *ma, const float *mb, float checker_scale, float spec, float refl) {
    const Material A = mat_of(ma);
    std::function<Material(Vec3, Vec3, float)> fn = [A](Vec3, Vec3, float) { return A; }; // `(pos, n, u) => m`, as the scene factories write it
    if (checker_scale != 0.0f) fn = ScenesRef::Checker(Vec3(ma[0], ma[1], ma[2]), Vec3(mb[0], mb[1], mb[2]), checker_scale);
    switch (kind) {
        case 0: return new Sphere(Vec3(p[0], p[1], p[2]), p[3], A);
        case 1: return new Plane(Vec3(p[0], p[1], p[2]), Vec3(p[3], p[4], p[5]), fn, spec, refl);
        case 2: return new Disk(Vec3(p[0], p[1], p[2]), Vec3(p[3], p[4], p[5]), p[6], fn, spec, refl);
        case 3: return new XYRect(p[0], p[1], p[2], p[3], p[4], fn, spec, refl);
        case 4: return new XZRect(p[0], p[1], p[2], p[3], p[4], fn, spec, refl);
        case 5: return new YZRect(p[0], p[1], p[2], p[3], p[4], fn, spec, refl);
        case 6: return new Box(Vec3(p[0], p[1], p[2]), Vec3(p[3], p[4], p[5]), fn, spec, refl);
        case 7: return new CylinderY(Vec3(p[0], p[1], p[2]), p[3], p[4], p[5], p[6] != 0.0f, A);
        case 8: return new Triangle(Vec3(p[0], p[1], p[2]), Vec3(p[3], p[4], p[5]), Vec3(p[6], p[7], p[8]), A);
    }
    return nullptr;
}
}
// the scene's static images for the materials of the objects created AFTER this call (kept until the next call): w x h RGBA words, row 0 first
RH_API void ref_set_textures(int n, const int *w, const int *hh, const uint32_t *const *rgba) {
    g_textures.clear();
    for (int i = 0; i < n; i++) {
        std::unique_ptr<Texture> t(new Texture());
        t->width = w[i]; t->height = hh[i];
        t->pixels.assign((const int *)rgba[i], (const int *)rgba[i] + (size_t)w[i] * hh[i]);
        g_textures.push_back(std::move(t));
    }
}
// nearest hit over the objects IN ORDER with a shrinking tMax (what a leaf of BVH.Hit does with its items, BVH.cs:160-178)
RH_API int ref_objects_hit(int n_obj, const int *kind, const float *p12, const float *mat_a13, const float *mat_b13, const float *checker_scale, const float *spec, const float *refl,
                           int n, const float *rays6, float t_min, float t_max, int *id_out, float *t_out, float *n_out3, float *p_out3, float *mat_out5) {
    std::vector<Hittable *> objs;
    for (int k = 0; k < n_obj; k++) {
        objs.push_back(make_prim(kind[k], p12 + 12 * k, mat_a13 + MS * k, mat_b13 + MS * k, checker_scale[k], spec[k], refl[k]));
        if (!objs.back()) return -1;
    }
    for (int i = 0; i < n; i++) {
        const float *q = rays6 + 6 * i;
        const Ray r(Vec3(q[0], q[1], q[2]), Vec3(q[3], q[4], q[5]));
        HitRecord rec, tmp;
        float closest = t_max;
        int id = -1;
        for (int k = 0; k < n_obj; k++)
            if (objs[k]->Hit(r, t_min, closest, tmp, 0.0f, 0.0f)) { closest = tmp.T; rec = tmp; id = k; }
        id_out[i] = id;
        t_out[i] = id >= 0 ? rec.T : 0.0f;
        n_out3[3 * i] = rec.N.X; n_out3[3 * i + 1] = rec.N.Y; n_out3[3 * i + 2] = rec.N.Z;
        p_out3[3 * i] = rec.P.X; p_out3[3 * i + 1] = rec.P.Y; p_out3[3 * i + 2] = rec.P.Z;
        mat_out5[5 * i] = rec.Mat.Albedo.X; mat_out5[5 * i + 1] = rec.Mat.Albedo.Y; mat_out5[5 * i + 2] = rec.Mat.Albedo.Z;
        mat_out5[5 * i + 3] = (float)rec.Mat.Specular; mat_out5[5 * i + 4] = (float)rec.Mat.Reflectivity;
    }
    for (Hittable *h : objs) delete h;
    return 0;
}

// ---- MeshBVH.cs: the reference's own constructor (SoA, binned-SAH BuildRecursive with its partition and Array.Sort fallback) and Hit
RH_API void *ref_mesh_build(int n_tris, const float *abc9, const float *mat13) {
    try {
        const Material m = mat_of(mat13);
        std::vector<Triangle *> tris;
        for (int i = 0; i < n_tris; i++) {
            const float *t = abc9 + 9 * (size_t)i;
            tris.push_back(new Triangle(Vec3(t[0], t[1], t[2]), Vec3(t[3], t[4], t[5]), Vec3(t[6], t[7], t[8]), m));
        }
        MeshBVH *b = new MeshBVH(tris);
        for (Triangle *t : tris) delete t;
        return b;
    } catch (...) { return nullptr; }
}
RH_API void ref_mesh_destroy(void *h) { delete (MeshBVH *)h; }
// MeshScenes.AddMeshAutoGround (:173-184) as the reference wrote it: TryReadObjBoundsNormalized(path) -> the ground translate, then
// MeshLoader.FromObj(path, mat, scale, translate, normalize: true, targetSize: 1) -> the triangles A, B, C (9 floats each) the scene gets.
// Returns the triangle count (out_abc9 may be NULL to ask for it), -1 on a reference exception.
RH_API int ref_mesh_from_obj(const char *path, float scale, const float *target3, float *out_abc9, int cap_tris, float *translate3) {
    try {
        const Vec3 tr = MeshScenesRef::AutoGroundTranslate(String(path), scale, Vec3(target3[0], target3[1], target3[2]));
        if (translate3) { translate3[0] = tr.X; translate3[1] = tr.Y; translate3[2] = tr.Z; }
        List<Triangle *> tris = MeshLoaderRef::FromObj(String(path), Material(), scale, tr, true, 1.0f);
        const int n = tris.Count();
        if (out_abc9) for (int i = 0; i < n && i < cap_tris; i++) {
            const Triangle &t = *tris[i];
            const float v[9] = {t.A.X, t.A.Y, t.A.Z, t.B.X, t.B.Y, t.B.Z, t.C.X, t.C.Y, t.C.Z};
            std::memcpy(out_abc9 + 9 * (size_t)i, v, sizeof v);
        }
        for (int i = 0; i < n; i++) delete tris[i];
        return n;
    } catch (...) { return -1; }
}
RH_API void ref_mesh_info(void *h, int *n_nodes, int *root, int *n_leaf) { MeshBVH &b = *(MeshBVH *)h; *n_nodes = b.nodeCountUsed; *root = b.rootIndex; *n_leaf = (int)b.leafTriIndex.size(); }
RH_API void ref_mesh_tree(void *h, float *boxes6, int *lrsc4, int *leaf) {
    MeshBVH &b = *(MeshBVH *)h;
    for (int i = 0; i < b.nodeCountUsed; i++) {
        float *q = boxes6 + 6 * (size_t)i;
        q[0] = b.nodeMinX[i]; q[1] = b.nodeMinY[i]; q[2] = b.nodeMinZ[i]; q[3] = b.nodeMaxX[i]; q[4] = b.nodeMaxY[i]; q[5] = b.nodeMaxZ[i];
        int *w = lrsc4 + 4 * (size_t)i;
        w[0] = b.nodeLeft[i]; w[1] = b.nodeRight[i]; w[2] = b.nodeStart[i]; w[3] = b.nodeCount[i];
    }
    for (size_t i = 0; i < b.leafTriIndex.size(); i++) leaf[i] = b.leafTriIndex[i];
}
RH_API void ref_mesh_hit(void *h, int n, const float *rays6, float t_min, float t_max, uint8_t *hit, float *t_out, float *n_out3) {
    MeshBVH &b = *(MeshBVH *)h;
    for (int i = 0; i < n; i++) {
        const float *q = rays6 + 6 * i;
        const Ray r(Vec3(q[0], q[1], q[2]), Vec3(q[3], q[4], q[5]));
        HitRecord rec;
        hit[i] = b.Hit(r, t_min, t_max, rec, 0.0f, 0.0f) ? 1 : 0;
        t_out[i] = rec.T;
        n_out3[3 * i] = rec.N.X; n_out3[3 * i + 1] = rec.N.Y; n_out3[3 * i + 2] = rec.N.Z;
    }
}

// ---- the whole trace stage: the reference's Scene (transpiled Objects / Lights / BVH), its constructors, and the verbatim head of
// TryFlipAndBlit (frame counter, jitter rotations, MakeJitteredRay, PerFrameSeed, TraceFull per pixel)
namespace {
struct TraceHandle {
    SceneRef scene;
    RendererRef r;
    std::vector<Hittable *> owned;
    std::unique_ptr<TemporalAA> taa; // only so that the reference's Resize finds its `taa` (the decisions are taken on the post handle's)
};
}
RH_API void *ref_trace_create(int fb_w, int fb_h, int ss, float fov_deg, int n_obj, const int *kind, const float *p12, const float *mat_a13, const float *mat_b13,
                              const float *checker_scale, const float *spec, const float *refl, const int *mesh_tris, const float *const *mesh_abc9, const float *mesh_mat13,
                              int n_lights, const float *lights7, const float *bg_top3, const float *bg_bottom3, const float *ambient4,
                              int n_vols, const ycge_volume *vols, int n_mats, const float *scene_mats13, int is_volume_scene) {
    try {
        auto *h = new TraceHandle();
        h->scene.IsVolumeScene = is_volume_scene != 0;
        int mi = 0, vi = 0;
        for (int k = 0; k < n_obj; k++) {
            Hittable *o = nullptr;
            if (kind[k] == 9) { // Mesh: MeshLoader's triangles through the reference's Mesh / MeshBVH constructors
                const Material m = mat_of(mesh_mat13 + MS * mi);
                std::vector<Triangle *> tris;
                for (int i = 0; i < mesh_tris[mi]; i++) {
                    const float *t = mesh_abc9[mi] + 9 * (size_t)i;
                    tris.push_back(new Triangle(Vec3(t[0], t[1], t[2]), Vec3(t[3], t[4], t[5]), Vec3(t[6], t[7], t[8]), m));
                }
                o = new Mesh(tris, Vec3(0.0f, 0.0f, 0.0f), Vec3(0.0f, 0.0f, 0.0f)); // BoundsMin / BoundsMax are informational (Mesh.cs:9-10)
                for (Triangle *t : tris) delete t;
                mi++;
            } else if (kind[k] == 10) { // VolumeGrid: the fields its constructor derives (VolumeGrid.cs:55-93), from the flat description
                if (vi >= n_vols) { delete h; return nullptr; }
                const ycge_volume &v = vols[vi++];
                VolumeGrid *g = new VolumeGrid();
                g->nx = v.nx; g->ny = v.ny; g->nz = v.nz;
                g->nbx = (v.nx + 7) >> 3; g->nby = (v.ny + 7) >> 3; g->nbz = (v.nz + 7) >> 3;
                g->brickCount = g->nbx * g->nby * g->nbz; g->capacity = g->brickCount * 512;
                g->matPtr = const_cast<int *>(v.mat); g->metaPtr = const_cast<int *>(v.meta);
                g->minCorner = Vec3(v.min_corner[0], v.min_corner[1], v.min_corner[2]);
                g->voxelSize = Vec3(MathF::Max(1e-6f, v.voxel_size[0]), MathF::Max(1e-6f, v.voxel_size[1]), MathF::Max(1e-6f, v.voxel_size[2]));
                g->wireframe = v.wireframe != 0; g->wireWidthFrac = v.wire_width_frac; g->wireMaxDistance = v.wire_max_distance;
                std::vector<int> table(v.palette, v.palette + (size_t)v.palette_n_ids * (v.palette_meta_levels < 1 ? 1 : v.palette_meta_levels));
                std::vector<Material> mats;
                for (int m = 0; m < n_mats; m++) mats.push_back(mat_of(scene_mats13 + MS * m));
                const int n_ids = v.palette_n_ids, levels = v.palette_meta_levels < 1 ? 1 : v.palette_meta_levels, def = v.palette_default;
                g->materialLookup = [table, mats, n_ids, levels, def](int id, int meta) { // ycge.h: the palette as data
                    if (id >= n_ids || id < 0) return mats[(size_t)def];
                    const int m = meta < 0 ? 0 : (meta >= levels ? levels - 1 : meta);
                    return mats[(size_t)table[(size_t)id * levels + m]];
                };
                o = g;
            } else o = make_prim(kind[k], p12 + 12 * k, mat_a13 + MS * k, mat_b13 + MS * k, checker_scale[k], spec[k], refl[k]);
            if (!o) { delete h; return nullptr; }
            h->owned.push_back(o);
            h->scene.Objects.Add(o);
        }
        for (int i = 0; i < n_lights; i++) {
            const float *l = lights7 + 7 * i;
            h->scene.Lights.Add(PointLight(Vec3(l[0], l[1], l[2]), Vec3(l[3], l[4], l[5]), l[6]));
        }
        h->scene.BackgroundTop = Vec3(bg_top3[0], bg_top3[1], bg_top3[2]);
        h->scene.BackgroundBottom = Vec3(bg_bottom3[0], bg_bottom3[1], bg_bottom3[2]);
        h->scene.Ambient = AmbientLight(Vec3(ambient4[0], ambient4[1], ambient4[2]), ambient4[3]);
        h->scene.RebuildBVH();
        RendererRef &r = h->r;
        r.scene = &h->scene;
        r.ss = ss; r.fbW = fb_w; r.fbH = fb_h; r.procCount = refcs::ref_threads() > 1 ? refcs::ref_threads() : 3; r.fovDeg = fov_deg;
        r.hiW = fb_w * ss; r.hiH = fb_h * 2 * ss; // RaytraceRenderer.cs:86-87
        r.rays = Fast2D<Ray>(r.hiW, r.hiH);
        r.gAlbedo = Fast2D<Vec3>(r.hiW, r.hiH); r.gNormal = Fast2D<Vec3>(r.hiW, r.hiH); r.gDepth = Fast2D<float>(r.hiW, r.hiH); r.skyMask = Fast2D<bool>(r.hiW, r.hiH);
        r.frameBuffer = Fast2D<Chexel>(fb_w, fb_h);
        return h;
    } catch (...) { return nullptr; }
}
RH_API void ref_trace_destroy(void *hh) { TraceHandle *h = (TraceHandle *)hh; for (Hittable *o : h->owned) delete o; delete h; }
RH_API int ref_trace_resize(void *hh, int fb_w, int fb_h, int ss) { // the same Resize on the trace half: rays and G-buffer planes at the new size, frame counter kept
    TraceHandle &h = *(TraceHandle *)hh;
    try {
        if (!h.taa) { h.taa.reset(new TemporalAA(h.r.hiW, h.r.hiH, h.r.taaAlpha, RendererRef::MotionTransReset, RendererRef::MotionRotReset)); h.r.taa = h.taa.get(); }
        Framebuffer fb(fb_w, fb_h);
        h.r.Resize(&fb, ss);
        return 0;
    } catch (...) { return -1; }
}
// one frame of the trace stage; planes out (row-major hiW x hiH)
RH_API int ref_trace_frame(void *hh, const float *cam3, float yaw, float pitch, float *rays6, float *hdr3, float *albedo3, float *normal3, float *depth, uint8_t *sky) {
    TraceHandle &h = *(TraceHandle *)hh;
    try {
        h.r.TraceStage(Vec3(cam3[0], cam3[1], cam3[2]), yaw, pitch);
        const int W = h.r.hiW, H = h.r.hiH;
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            const size_t p = (size_t)x + (size_t)y * W;
            const Ray ry = h.r.rays[x, y];
            rays6[6 * p] = ry.Origin.X; rays6[6 * p + 1] = ry.Origin.Y; rays6[6 * p + 2] = ry.Origin.Z; rays6[6 * p + 3] = ry.Dir.X; rays6[6 * p + 4] = ry.Dir.Y; rays6[6 * p + 5] = ry.Dir.Z;
            const Vec3 c = h.r.currentHdr[x, y], a = h.r.gAlbedo[x, y], n = h.r.gNormal[x, y];
            hdr3[3 * p] = c.X; hdr3[3 * p + 1] = c.Y; hdr3[3 * p + 2] = c.Z;
            albedo3[3 * p] = a.X; albedo3[3 * p + 1] = a.Y; albedo3[3 * p + 2] = a.Z;
            normal3[3 * p] = n.X; normal3[3 * p + 1] = n.Y; normal3[3 * p + 2] = n.Z;
            depth[p] = h.r.gDepth[x, y];
            sky[p] = h.r.skyMask[x, y] ? 1 : 0;
        }
        return 0;
    } catch (...) { return -1; }
}

// worker threads of FixedThreadFor / PixelThreadPool (the reference uses Environment.ProcessorCount); 1 = serial
RH_API void ref_set_threads(int n) { refcs::ref_threads() = n < 1 ? 1 : n; }
