// wavefront_layout.h — geometry and addressing of the systolic wavefront form of the in-place à-trous iteration
// (RaytraceRenderer.cs:651-719 with cur == dst, stride 2).  Plain integer functions for host and device: the pre-pass
// kernel, the wavefront kernel, the host launcher and the CPU schedule simulator of the test suite all derive their
// addresses from here.
//
// The pass filters in place in row-major order, so pixel (x, y) reads ALREADY FILTERED values at taps that precede it
// (rows above; same row to the left) and unfiltered ones elsewhere.  With stride 2 a pixel's taps lie at x +- 2, 4 and
// y +- 2, 4 (then clamped to the image): a pixel row holds two independent CHAINS (x even / x odd), rows of equal parity
// depend on each other, and a chain may run L = 3 pixels behind the chain above it ((x + 4, y - 2) must be done).
//
// One WARP owns a BAND: 4 rows of one row parity x 2 chains = 8 chains, 4 lanes per chain (a QUAD; lane q evaluates the
// taps 3q .. 3q+2, every lane of the quad then adds the 25 terms in the reference's order).  All
// chains of a band advance in LOCK STEP, one pixel per step: chain (row r, column parity cx) is at pixel index
//     i = t - L * r - cx                      (x = 2 i + cx)
// in step t.  With that schedule every filtered value a pixel needs from its own band — the two rows above, the pixels to
// its left, and the clamped border pixels, which belong to the SIBLING chain of a row — was produced in an earlier step,
// by construction: inside a band the wavefront needs no flag, no poll and no memory round trip; values travel through a
// small shared-memory history ring.  Only the two rows above a band come from another warp (through L2, prefetched).
#pragma once

#if defined(__CUDACC__)
#define YWF_HD __host__ __device__ __forceinline__
#else
#define YWF_HD static inline
#endif

#define YCGE_WF_L 3        // steps a row runs behind the row of equal parity above it
#define YCGE_WF_ROWS 4     // rows per band
#define YCGE_WF_CHAINS 8   // chains per band
#define YCGE_WF_SLOTS 26   // records per pixel: 25 taps + the centre
#define YCGE_WF_RING 32    // history entries per chain: a band reads values up to 2 L + 3 = 9 steps old, and the rows above it are brought in up to AHEAD steps early
#define YCGE_WF_AHEAD 16   // steps the halo warp may run ahead of its band
#define YCGE_WF_DEPTH 8    // steps of records in flight in shared memory (a power of two)
#define YCGE_WF_CLUSTER 8  // bands per thread-block cluster of the OPT-IN cluster form (hand-off through distributed shared memory; measured slower, see wavefront.cuh)
#define YCGE_WF_LEAD 6     // steps the halo warp starts before step 0, >= 2 L

struct WfGeom {
    int W, H, y0, y1; // the pass covers pixel rows [y0, y1) of a W x H image
    int yf[2];        // first row of parity cy in [y0, y1)
    int hs[2];        // rows of parity cy
    int nb[2];        // bands of parity cy
    int ws[2];        // pixels of column parity cx in a row
    int nt;           // steps a band takes (t = 0 .. nt - 1)
    int n_warps;      // bands in ticket order: g = 2 * b + cy
};

YWF_HD WfGeom wf_geom(int W, int H, int y0, int y1) {
    WfGeom g;
    g.W = W; g.H = H; g.y0 = y0; g.y1 = y1;
    int nbmax = 0;
    for (int cy = 0; cy < 2; cy++) {
        g.yf[cy] = y0 + ((cy - y0) & 1);
        g.hs[cy] = g.yf[cy] < y1 ? (y1 - g.yf[cy] + 1) / 2 : 0;
        g.nb[cy] = (g.hs[cy] + YCGE_WF_ROWS - 1) / YCGE_WF_ROWS;
        if (g.nb[cy] > nbmax) nbmax = g.nb[cy];
        g.ws[cy] = (W - cy + 1) / 2;
    }
    g.nt = g.ws[0] + YCGE_WF_L * (YCGE_WF_ROWS - 1) + 2;
    g.n_warps = 2 * nbmax;
    return g;
}
YWF_HD size_t wf_record_count(const WfGeom &g) { return (size_t)g.n_warps * (size_t)g.nt * (YCGE_WF_SLOTS * YCGE_WF_CHAINS); }

// where pixel (x, y) of the pass is processed
struct WfPlace { int warp, step, chain, yb0; };
YWF_HD WfPlace wf_place(const WfGeom &g, int x, int y) {
    const int cy = y & 1, j = (y - g.yf[cy]) >> 1, b = j / YCGE_WF_ROWS, r = j - b * YCGE_WF_ROWS, cx = x & 1;
    WfPlace p;
    p.warp = 2 * b + cy;
    p.step = (x >> 1) + YCGE_WF_L * r + cx;
    p.chain = 2 * r + cx;
    p.yb0 = g.yf[cy] + 2 * YCGE_WF_ROWS * b;
    return p;
}
// The 26 x 8 records of one (band, step) are contiguous (3 328 bytes) and travel to shared memory verbatim (cp.async), so
// the position inside the block is chosen for the shared-memory banks: the four lanes of a quad read the slots 3q + j of
// their chain at the same time, which a plain [slot][chain] layout would put on the same banks; rotating the chain
// index by 2 * (slot / 3) gives the eight lanes of a quarter warp eight different 16-byte columns.
YWF_HD int wf_record_column(int slot, int chain) { return (chain + 2 * (slot / 3)) & (YCGE_WF_CHAINS - 1); }
YWF_HD size_t wf_record_index(const WfGeom &g, const WfPlace &p, int slot) {
    return (((size_t)p.warp * (size_t)g.nt + (size_t)p.step) * YCGE_WF_SLOTS + (size_t)slot) * YCGE_WF_CHAINS + (size_t)wf_record_column(slot, p.chain);
}
// the two rows above a band that another warp (or an earlier launch, or a peer GPU) produces; needed iff < yb0
YWF_HD int wf_halo_row(int yb0, int h) { const int y = yb0 - 4 + 2 * h; return y < 0 ? 0 : y; }
// history entry (in float4 units) of the filtered value of pixel (sx, sy) as seen from a band whose first row is yb0
YWF_HD int wf_history_entry(int yb0, int sx, int sy) {
    const int hrow = sy >= yb0 ? 2 + ((sy - yb0) >> 1) : (sy == wf_halo_row(yb0, 1) ? 1 : 0);
    return (hrow * 2 + (sx & 1)) * YCGE_WF_RING + ((sx >> 1) & (YCGE_WF_RING - 1));
}
#define YCGE_WF_HISTORY_ENTRIES ((YCGE_WF_ROWS + 2) * 2 * YCGE_WF_RING)
