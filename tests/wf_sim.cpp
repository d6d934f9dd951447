// wf_sim.cpp — CPU simulator of the wavefront kernel's SCHEDULE (test infrastructure, built and run by
// tests/test_wavefront_layout.py with g++; no GPU involved).
//
// It runs an in-place 5x5 stride-2 filter with a non-linear, value-dependent weight twice over random planes:
//   (1) literally: row-major, in place, taps clamped to the image (the reference's loop, RaytraceRenderer.cs:651-719);
//   (2) the way atrous_wave_kernel does: per pixel 26 records laid out by wavefront_layout.h (old taps as finished terms,
//       new taps as a history address), bands in ticket order, chains in lock step i = t - L r - cx, every new value
//       read from the band's history rings, the two rows above a band committed from "global memory" by the halo warp:
//       the pixels of "step s" (3 resp. 6 pixels ahead of the band's first row) anywhere between AHEAD steps early and
//       just in time -- the simulator runs both extremes (last argument: 0 = just in time, 1 = as early as allowed).
// Every history read is checked against a tag: it must hold exactly the pixel the tap names, written in an EARLIER step,
// and no entry may be overwritten in a step in which it is read; a halo commit must find its pixel already produced by a
// band with a lower ticket.  The two results must be bit-identical.  Exit code 0 and "ok" on success.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../yetanotherconsolegameengine_b200/csrc/wavefront_layout.h"

struct Px { float r, g, b, l; };
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float wfun(float d) { return 1.0f / (1.0f + 3.0f * d + d * d); } // stands in for exp(-d / phi)
static inline float luma(float r, float g, float b) { return 0.2126f * r + 0.7152f * g + 0.0722f * b; }
static const float KW[5] = {1.f / 16.f, 1.f / 4.f, 3.f / 8.f, 1.f / 4.f, 1.f / 16.f};

struct Rec { float x, y, z; int32_t code; float w; }; // code >= 0: new tap, history entry; < 0: finished term (x, y, z, w)

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: wf_sim W H y0 y1 [seed] [early]\n"); return 2; }
    const int early_halo = argc > 6 ? atoi(argv[6]) : 0;
    const int W = atoi(argv[1]), H = atoi(argv[2]), y0 = atoi(argv[3]), y1 = atoi(argv[4]);
    uint32_t seed = argc > 5 ? (uint32_t)atoi(argv[5]) : 1u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (float)((seed >> 8) & 0xFFFF) / 65536.0f; };
    std::vector<Px> old_((size_t)W * H);
    std::vector<float> guide((size_t)W * H);
    std::vector<uint8_t> sky((size_t)W * H);
    for (size_t p = 0; p < old_.size(); p++) {
        old_[p].r = rnd() * 2.0f; old_[p].g = rnd(); old_[p].b = rnd() * 0.5f; old_[p].l = luma(old_[p].r, old_[p].g, old_[p].b);
        guide[p] = rnd();
        sky[p] = rnd() < 0.07f;
    }
    // ---- (1) the literal in-place pass over rows [y0, y1); rows above y0 are "already new" (another tile's output)
    std::vector<Px> ref = old_;
    for (int y = 0; y < y0; y++) for (int x = 0; x < W; x++) { Px &p = ref[(size_t)y * W + x]; p.r *= 0.5f; p.g *= 0.75f; p.l = luma(p.r, p.g, p.b); }
    const std::vector<Px> above = ref; // what the pass finds in the rows above y0
    for (int y = y0; y < y1; y++) for (int x = 0; x < W; x++) {
        const size_t pix = (size_t)y * W + x;
        const Px c0 = ref[pix];
        if (sky[pix]) continue;
        float wsum = 0, ax = 0, ay = 0, az = 0;
        for (int ky = -2; ky <= 2; ky++) for (int kx = -2; kx <= 2; kx++) {
            const int sy = clampi(y + 2 * ky, 0, H - 1), sx = clampi(x + 2 * kx, 0, W - 1);
            const size_t sp = (size_t)sy * W + sx;
            if (sky[sp] != sky[pix]) continue;
            const Px c = ref[sp];
            const float w = KW[kx + 2] * KW[ky + 2] * wfun(fabsf(c.l - c0.l)) * wfun(fabsf(guide[sp] - guide[pix]));
            ax = ax + c.r * w; ay = ay + c.g * w; az = az + c.b * w; wsum += w;
        }
        Px o = c0;
        if (wsum > 1e-8f) { const float inv = 1.0f / wsum; o.r = ax * inv; o.g = ay * inv; o.b = az * inv; }
        o.l = luma(o.r, o.g, o.b);
        ref[pix] = o;
    }
    // ---- (2) records
    const WfGeom g = wf_geom(W, H, y0, y1);
    std::vector<Rec> rec(wf_record_count(g));
    std::vector<uint8_t> rec_written(rec.size(), 0);
    for (int y = y0; y < y1; y++) for (int x = 0; x < W; x++) {
        const size_t pix = (size_t)y * W + x;
        const WfPlace pl = wf_place(g, x, y);
        const Px c0 = old_[pix];
        for (int ky = -2; ky <= 2; ky++) for (int kx = -2; kx <= 2; kx++) {
            const int sy = clampi(y + 2 * ky, 0, H - 1), sx = clampi(x + 2 * kx, 0, W - 1);
            const size_t sp = (size_t)sy * W + sx;
            const bool is_new = sy < y || (sy == y && sx < x);
            const bool skip = sky[pix] || sky[sp] != sky[pix];
            Rec r;
            if (skip) { r.x = r.y = r.z = 0; r.w = 0; r.code = -1; }
            else if (is_new) { r.x = wfun(fabsf(guide[sp] - guide[pix])); r.y = r.z = r.w = 0; r.code = wf_history_entry(pl.yb0, sx, sy); }
            else {
                const Px c = old_[sp];
                const float w = KW[kx + 2] * KW[ky + 2] * wfun(fabsf(c.l - c0.l)) * wfun(fabsf(guide[sp] - guide[pix]));
                r.x = c.r * w; r.y = c.g * w; r.z = c.b * w; r.w = w; r.code = -1;
            }
            const size_t ri = wf_record_index(g, pl, (ky + 2) * 5 + (kx + 2));
            if (ri >= rec.size() || rec_written[ri]) { printf("record index collision / overflow at (%d,%d)\n", x, y); return 1; }
            rec[ri] = r; rec_written[ri] = 1;
        }
        Rec c; c.x = c0.r; c.y = c0.g; c.z = c0.b; c.w = c0.l; c.code = 0;
        const size_t ri = wf_record_index(g, pl, 25);
        if (ri >= rec.size() || rec_written[ri]) { printf("centre record collision at (%d,%d)\n", x, y); return 1; }
        rec[ri] = c; rec_written[ri] = 1;
    }
    // ---- (2) the schedule
    std::vector<Px> out = above;
    std::vector<uint8_t> produced((size_t)W * H, 0);
    for (int y = 0; y < y0; y++) for (int x = 0; x < W; x++) produced[(size_t)y * W + x] = 1;
    long violations = 0;
    struct Ent { Px v; int tag_x, tag_y, t; };
    for (int warp = 0; warp < g.n_warps; warp++) {
        const int cy = warp & 1, b = warp >> 1;
        if (b >= g.nb[cy]) continue;
        const int yb0 = g.yf[cy] + 2 * YCGE_WF_ROWS * b;
        std::vector<Ent> hist(YCGE_WF_HISTORY_ENTRIES, Ent{Px{0, 0, 0, 0}, -1, -1, -1000000});
        int halo_next = -YCGE_WF_LEAD;
        for (int t = -YCGE_WF_LEAD - (early_halo ? YCGE_WF_AHEAD : 0); t < g.nt; t++) {
            std::vector<int> read_set;
            struct Wr { int e; Ent v; };
            std::vector<Wr> writes;
            // halo warp: rows h = 0, 1 (virtual rows r = -2, -1), both column parities; the pixels of step s enter the history
            // while the band is in step t = s (just in time) or t = s - AHEAD (as early as the kernel allows)
            const int s_hi = early_halo ? t + YCGE_WF_AHEAD : t;
            for (; halo_next <= s_hi && halo_next < g.nt; halo_next++) {
                const int sc = halo_next;
                for (int h = 0; h < 2; h++) for (int cxh = 0; cxh < 2; cxh++) {
                    const int hy = wf_halo_row(yb0, h);
                    if (hy >= yb0) continue; // not a halo: the band's own row (top of the image)
                    const int ih = sc + YCGE_WF_L * (2 - h) - cxh;
                    if (ih < 0 || ih >= g.ws[cxh]) continue;
                    const int x = 2 * ih + cxh;
                    if (!produced[(size_t)hy * W + x]) { if (violations++ < 10) printf("halo (%d,%d) not produced before warp %d step %d\n", x, hy, warp, t); }
                    writes.push_back(Wr{(h * 2 + cxh) * YCGE_WF_RING + (ih & (YCGE_WF_RING - 1)), Ent{out[(size_t)hy * W + x], x, hy, t}});
                }
            }
            for (int c = 0; c < YCGE_WF_CHAINS; c++) {
                const int r = c >> 1, cx = c & 1, i = t - YCGE_WF_L * r - cx, y = yb0 + 2 * r, x = 2 * i + cx;
                if (y >= y1 || i < 0 || i >= g.ws[cx]) continue;
                WfPlace pl; pl.warp = warp; pl.step = t; pl.chain = c; pl.yb0 = yb0;
                const Rec cen = rec[wf_record_index(g, pl, 25)];
                float acc[4] = {0, 0, 0, 0};
                for (int k = 0; k < 25; k++) {
                    const Rec rc = rec[wf_record_index(g, pl, k)];
                    float term[4];
                    if (rc.code >= 0) {
                        const int ky = k / 5 - 2, kx = k % 5 - 2;
                        const int sy = clampi(y + 2 * ky, 0, H - 1), sx = clampi(x + 2 * kx, 0, W - 1);
                        if (rc.code >= YCGE_WF_HISTORY_ENTRIES) { printf("history address out of range\n"); return 1; }
                        const Ent &e = hist[rc.code];
                        if (e.tag_x != sx || e.tag_y != sy || e.t >= t) {
                            if (violations++ < 10) printf("pixel (%d,%d) tap %d wants (%d,%d) at step %d, history holds (%d,%d) written at %d\n", x, y, k, sx, sy, t, e.tag_x, e.tag_y, e.t);
                        }
                        read_set.push_back(rc.code);
                        const Px tv = e.v;
                        const float w = KW[kx + 2] * KW[ky + 2] * wfun(fabsf(tv.l - cen.w)) * rc.x;
                        term[0] = tv.r * w; term[1] = tv.g * w; term[2] = tv.b * w; term[3] = w;
                    } else { term[0] = rc.x; term[1] = rc.y; term[2] = rc.z; term[3] = rc.w; }
                    for (int q = 0; q < 4; q++) acc[q] = acc[q] + term[q];
                }
                Px o; o.r = cen.x; o.g = cen.y; o.b = cen.z;
                if (acc[3] > 1e-8f) { const float inv = 1.0f / acc[3]; o.r = acc[0] * inv; o.g = acc[1] * inv; o.b = acc[2] * inv; }
                o.l = luma(o.r, o.g, o.b);
                if (sky[(size_t)y * W + x]) o = Px{cen.x, cen.y, cen.z, cen.w}; // the literal loop leaves a sky pixel alone
                writes.push_back(Wr{((r + 2) * 2 + cx) * YCGE_WF_RING + (i & (YCGE_WF_RING - 1)), Ent{o, x, y, t}});
            }
            for (const Wr &w : writes) {
                for (int e : read_set) if (e == w.e) { if (violations++ < 10) printf("history entry %d overwritten in step %d while it is read\n", e, t); }
                hist[w.e] = w.v;
                if (w.v.tag_y >= yb0) { out[(size_t)w.v.tag_y * W + w.v.tag_x] = w.v.v; produced[(size_t)w.v.tag_y * W + w.v.tag_x] = 1; }
            }
        }
    }
    long differ = 0, missing = 0;
    for (int y = y0; y < y1; y++) for (int x = 0; x < W; x++) {
        const size_t p = (size_t)y * W + x;
        if (!produced[p]) missing++;
        if (memcmp(&out[p], &ref[p], sizeof(Px)) != 0) { if (differ++ < 5) printf("pixel (%d,%d) differs: %g %g %g %g vs %g %g %g %g\n", x, y, out[p].r, out[p].g, out[p].b, out[p].l, ref[p].r, ref[p].g, ref[p].b, ref[p].l); }
    }
    printf("%dx%d rows [%d,%d): warps %d, steps %d, violations %ld, missing %ld, differing %ld -> %s\n", W, H, y0, y1, g.n_warps, g.nt, violations, missing, differ,
           (violations || missing || differ) ? "FAIL" : "ok");
    return (violations || missing || differ) ? 1 : 0;
}
