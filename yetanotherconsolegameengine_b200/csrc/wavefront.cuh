// wavefront.cuh — K3'': the in-place à-trous iteration (RaytraceRenderer.cs:651-719 with cur == dst, stride 2) as a
// systolic wavefront (sm_100a).  Geometry, schedule and addressing: wavefront_layout.h (read that first).
//
//   atrous_wave_pre_kernel   one thread per pixel, fully parallel: 26 records per pixel, written where the band that owns
//                            the pixel will stream them from — a tap that FOLLOWS the pixel in row-major order (unfiltered
//                            input) becomes its finished weighted term, a tap that PRECEDES it (filtered value, not known
//                            yet) becomes its three guide weights plus the ADDRESS of that value in the band's history
//                            rings.  Clamping at the image border, rows 0 / H-1 (where taps of other kernel rows fold
//                            onto the pixel's own row) and sky are resolved here, per tap: the wavefront kernel has no
//                            border cases.
//   atrous_wave_kernel       one warp per band of 4 rows x 2 chains, 4 lanes per chain (lane = r, g, b, weight); all chains
//                            of a band advance one pixel per step in lock step (i = t - 3 r - cx), so that everything a
//                            pixel needs from its own band is in the shared-memory history by construction: no flags, no
//                            polling, no memory round trip inside a band.  Per step a lane
//                              1. evaluates the colour weight exp(-|dlum| / cPhi) of 3 of the 12 filtered taps and the product
//                                 wBase*wc*wn*wz*wa in the reference's order (:699), shuffles them round the quad;
//                              2. adds its channel's 25 terms in the reference's ky-major / kx order;
//                              3. normalises, shuffles r, g, b round the quad for the luma, publishes (history + L2).
//                            The records arrive by cp.async 7 steps ahead; the two rows above the band (another warp's, an
//                            earlier launch's or a peer GPU's output) are read from L2 two steps ahead of their commit to
//                            the history and validated against the all-ones sentinel the output buffer is pre-filled with.
//   Bands take tickets in dispatch order; a band only ever waits for lower tickets, i.e. for warps that are running or
//   done, so the kernel needs no co-residency guarantee and shares the GPU with other frames' kernels.  Every poll has a
//   bound: a value that does not arrive sets *err and the frame fails with YCGE_ERR_CUDA instead of hanging the GPU.
#pragma once
#include "post.cuh"
#include "wavefront_layout.h"

namespace ycge {

__device__ __forceinline__ float wf_kw(int k) { return k == 0 ? 3.f / 8.f : ((k == 1 || k == -1) ? 1.f / 4.f : 1.f / 16.f); }

// shared-memory history: every entry is stored twice, at e and e + RING, so that the five entries i - 2 .. i + 2 of a row
// are contiguous from ((i - 2) & 15) without a wrap: the pipelined path reads all filtered taps at immediate offsets from
// one moving pointer.  A second array of the same shape holds 1.0f: the lane that sums the WEIGHT channel reads its "colour"
// from there (term.w = 1 * w), with the same addresses as the colour lanes.
#define YCGE_WF_HROW_BYTES (2 * 2 * YCGE_WF_RING * 16)                          // one history row: 2 chains x 32 entries x 16 B
#define YCGE_WF_HIST_FLOATS ((YCGE_WF_ROWS + 2) * 2 * 2 * YCGE_WF_RING * 4)
__host__ __device__ __forceinline__ int wf_history_offset(int entry) { return ((entry >> 4) * (2 * YCGE_WF_RING) + (entry & (YCGE_WF_RING - 1))) * 16; } // logical entry -> byte offset

struct WavePreArgs {
    const float4 *old_, *gnd, *gas;
    float4 *rec;
    WfGeom g;
    EdgeDiv e;
};

template <bool FAST> __global__ void __launch_bounds__(256) atrous_wave_pre_kernel(WavePreArgs a) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = a.g.y0 + blockIdx.y * 8 + threadIdx.y;
    const int W = a.g.W, H = a.g.H;
    if (x >= W || y >= a.g.y1) return;
    const size_t pix = (size_t)x + (size_t)y * W;
    const float4 c0 = __ldg(&a.old_[pix]);
    const float4 as0 = __ldg(&a.gas[pix]);
    const float4 nd0 = __ldg(&a.gnd[pix]);
    const bool sky0 = as0.w != 0.0f; // sky centre: every tap contributes nothing, the wavefront then falls back to c0 (:659)
    const WfPlace pl = wf_place(a.g, x, y);
    float4 *out = a.rec + wf_record_index(a.g, pl, 0);
#pragma unroll 1
    for (int ky = -2; ky <= 2; ky++) {
        const int sy = clampi(y + ky * 2, 0, H - 1);
        const float wy = wf_kw(ky);
#pragma unroll
        for (int kx = -2; kx <= 2; kx++) {
            const int sx = clampi(x + kx * 2, 0, W - 1);
            const size_t sp = (size_t)sx + (size_t)sy * W;
            const bool is_new = (sy < y) || (sy == y && sx < x);
            const float4 as = __ldg(&a.gas[sp]);
            const bool skip = sky0 || as.w != as0.w;
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f); // a skipped tap adds +0 to sums that are never -0: exact
            if (!skip) {
                const float4 c = __ldg(&a.old_[sp]);
                const float4 nd = __ldg(&a.gnd[sp]);
                const float wBase = wf_kw(kx) * wy;
                const float dl = fabsf(c.w - c0.w);
                const float dn = MaxF(0.0f, 1.0f - (nd0.x * nd.x + nd0.y * nd.y + nd0.z * nd.z));
                const float dz = fabsf(nd.w - nd0.w);
                const float da = fabsf(as.x - as0.x) + fabsf(as.y - as0.y) + fabsf(as.z - as0.z);
                const float wn = edge_weight_vote<FAST>(dn, a.e.dn, a.e.rn);
                const float wz = exp_nonpos(neg_div<FAST>(dz, a.e.dz, a.e.rz));
                const float wa = edge_weight_vote<FAST>(da, a.e.da, a.e.ra);
                if (is_new) v = make_float4(wn, wz, wa, __int_as_float(wf_history_offset(wf_history_entry(pl.yb0, sx, sy)) | (int)0x80000000)); // sign bit = "filtered tap"
                else {
                    const float wc = exp_nonpos(neg_div<FAST>(dl, a.e.dc, a.e.rc));
                    const float wght = wBase * wc * wn * wz * wa;
                    v = make_float4(c.x * wght, c.y * wght, c.z * wght, wght == wght ? wght : __int_as_float(0x7FC00000)); // a finished term: never negative (a NaN keeps the sign clear)
                }
            }
            out[(size_t)((ky + 2) * 5 + (kx + 2)) * YCGE_WF_CHAINS] = v;
        }
    }
    out[(size_t)25 * YCGE_WF_CHAINS] = c0;
}

struct WaveArgs {
    const float4 *rec;
    float4 *new_;      // pass output, pre-filled with the sentinel on the rows this pass (or a peer) produces
    WfGeom g;
    float dc, rc;      // max(1e-6, cPhi) and its reciprocal
    unsigned int *ticket;
    unsigned int ticket_base;
    int *err;          // mapped host memory: set to 1 when a poll gives up
    unsigned long long *trace; // development aid: globaltimer at the first and the last step of every band, or NULL
    // multi-GPU: rows [peer_y0, peer_y1) are ALSO stored into the output buffer of the rank below (see post.cuh)
    float4 *peer_new;
    int peer_y0, peer_y1;
    const int *ready;
    int frame;
};

__device__ __forceinline__ float ld_relaxed_f32(const float *p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_f32(float *p, float v) { asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed_sys_f32(float *p, float v) { asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
#define YCGE_WF_POLL_LIMIT (1 << 21) // ~1 s of L2 round trips

template <bool FAST, bool PEER> __global__ void __launch_bounds__(YCGE_WF_ROWS * 32) atrous_wave_kernel(WaveArgs a) {
    __shared__ __align__(16) float4 s_rec[YCGE_WF_DEPTH][YCGE_WF_SLOTS * YCGE_WF_CHAINS];
    __shared__ __align__(16) float s_hist[2][YCGE_WF_HIST_FLOATS]; // [0]: filtered values (r, g, b, luma); [1]: ones
    __shared__ int s_ticket;
    const unsigned int FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, r = tid >> 5, cx = lane >> 4, s = lane & 15, q = s & 3, c = 2 * r + cx;
    const int hbase = lane & 16; // first lane of this chain's half warp
    if (tid == 0) s_ticket = (int)(atomicAdd(a.ticket, 1u) - a.ticket_base);
    for (int k = tid; k < YCGE_WF_HIST_FLOATS; k += YCGE_WF_ROWS * 32) { s_hist[0][k] = 0.0f; s_hist[1][k] = 1.0f; }
    __syncthreads();
    const int band = s_ticket;
    const WfGeom &g = a.g;
    const int cy = band & 1, b = band >> 1;
    if (band >= g.n_warps || b >= g.nb[cy]) return;
    const int yb0 = g.yf[cy] + 2 * YCGE_WF_ROWS * b;
    const int y = yb0 + 2 * r;
    const bool row_ok = y < g.y1;
    const bool row_reg = wf_row_regular(g, y);
    const int ws_c = g.ws[cx];
    const bool writer = s < 4; // lanes 0..3 of a half warp = channels r, g, b, weight of the chain
    // the slot whose colour weight this lane evaluates: 0..11, and on the last row 15, 16, 20, 21 (lanes 12..15)
    const int my_slot = s < 12 ? s : (s == 12 ? 15 : (s == 13 ? 16 : (s == 14 ? 20 : 21)));
    const int my_ky = my_slot / 5 - 2, my_kx = my_slot % 5 - 2;
    const float my_wB = wf_kw(my_kx) * wf_kw(my_ky);
    const int my_d = (s == 9 || s == 11 || s >= 12) ? 0 : 1; // pipelined path: slots 9 and 11 belong to this step's pixel, the others to the next one
    const int my_rowoff = ((r + my_ky + 2) * 2 + cx) * (2 * YCGE_WF_RING) * 4 + 3; // float index of the luma of entry 0 of my slot's history row (rows above / own)
    float *hist_f = s_hist[0];
    const float *chan_f = (q == 3 ? s_hist[1] : s_hist[0] + q); // where this lane's channel of a history entry lives
    const int own_row_f = ((r + 2) * 2 + cx) * (2 * YCGE_WF_RING) * 4;
    const int geo_row_f = (r * 2 + cx) * (2 * YCGE_WF_RING) * 4;    // history row of (y - 4): the pipelined path's window origin
    float *new_f = reinterpret_cast<float *>(a.new_);
    float *out_row = new_f + (size_t)y * g.W * 4;
    float *peer_row = nullptr;
    if (PEER) {
        if (a.peer_new && row_ok && y >= a.peer_y0 && y < a.peer_y1) peer_row = reinterpret_cast<float *>(a.peer_new) + (size_t)y * g.W * 4;
        if (a.peer_new && tid == 0) { // the rank below must have reset its buffer for this frame before anything is stored into it
            int v, n = 0;
            do { asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(a.ready) : "memory"); } while (v < a.frame && ++n < YCGE_WF_POLL_LIMIT);
            if (v < a.frame) *(volatile int *)a.err = 1;
        }
    }
    // halo loader: warp 0, lanes 0..15 = (row h above the band, column parity, channel)
    const int lh = (lane >> 3) & 1, lcx = (lane >> 2) & 1;
    const int hy = wf_halo_row(yb0, lh);
    const bool loader = r == 0 && lane < 16 && hy < yb0; // at the top of the image the rows above fold onto the band's own rows: no halo
    const float *halo_row = new_f + (size_t)hy * g.W * 4 + q;
    const int ws_l = g.ws[lcx];
    auto halo_ptr = [&](int t) -> const float * { // the word this lane commits in step t, or NULL
        const int ih = t + YCGE_WF_L * (2 - lh) - lcx;
        return (loader && ih >= 0 && ih < ws_l) ? halo_row + (size_t)(2 * ih + lcx) * 4 : nullptr;
    };
    float pf[YCGE_WF_PF];
#pragma unroll
    for (int k = 0; k < YCGE_WF_PF; k++) { const float *p = halo_ptr(-YCGE_WF_LEAD + k); pf[k] = p ? ld_relaxed_f32(p) : 0.0f; }
    const float4 *rec_g = a.rec + (size_t)band * (size_t)g.nt * (YCGE_WF_SLOTS * YCGE_WF_CHAINS);
    float P = 0.0f, t10 = 0.0f; // pipelined path: ordered partial sum of slots 0..8 and the term of slot 10 of the NEXT pixel
    bool pipe_valid = false;    // ... valid for the pixel of the coming step
    const int RS = YCGE_WF_CHAINS * 4;          // floats between consecutive slots of a chain's records
    const int HR = YCGE_WF_HROW_BYTES / 4;      // floats per history row
    // normalise (:706-714), luma, publish: shared by both paths
    auto publish = [&](float acc, const float *rec_t, int i, bool store) {
        const float wsum = __shfl_sync(FULL, acc, hbase | 3);
        const float inv = rcp_rn<FAST>(wsum);
        const float resq = wsum > 1e-8f ? acc * inv : rec_t[25 * RS + q];
        const float rr = __shfl_sync(FULL, resq, hbase), gg = __shfl_sync(FULL, resq, hbase | 1), bb = __shfl_sync(FULL, resq, hbase | 2);
        const float outv = q == 3 ? luma3(rr, gg, bb) : resq;
        if (store && writer) {
            float *h = hist_f + own_row_f + (i & (YCGE_WF_RING - 1)) * 4 + q;
            h[0] = outv; h[YCGE_WF_RING * 4] = outv;
            const size_t xo = (size_t)(2 * i + cx) * 4 + q;
            st_relaxed_f32(out_row + xo, outv);
            if (PEER && peer_row) st_relaxed_sys_f32(peer_row + xo, outv);
        }
    };
    __syncthreads();

#pragma unroll 1
    for (int t = -YCGE_WF_LEAD; t < g.nt; t++) {
        if (a.trace && tid == 0 && t == 0) { unsigned long long tm; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm)); a.trace[2 * band] = tm; }
        // ---- halo commit: the word loaded PF steps ago must be valid by now (else poll), then it enters the history
        {
            const float *p = halo_ptr(t);
            float v = pf[0];
#pragma unroll
            for (int k = 0; k + 1 < YCGE_WF_PF; k++) pf[k] = pf[k + 1];
            const float *pn = halo_ptr(t + YCGE_WF_PF);
            pf[YCGE_WF_PF - 1] = pn ? ld_relaxed_f32(pn) : 0.0f;
            bool ok = !p || __float_as_uint(v) != YCGE_SENTINEL;
            if (!__all_sync(FULL, ok)) {
                int n = 0;
                while (!ok && ++n < YCGE_WF_POLL_LIMIT) { v = ld_relaxed_f32(p); ok = __float_as_uint(v) != YCGE_SENTINEL; }
                if (!ok) *(volatile int *)a.err = 1;
                __syncwarp();
            }
            if (p) {
                const int ih = t + YCGE_WF_L * (2 - lh) - lcx;
                float *h = hist_f + ((lh * 2 + lcx) * (2 * YCGE_WF_RING) + (ih & (YCGE_WF_RING - 1))) * 4 + q;
                h[0] = v; h[YCGE_WF_RING * 4] = v;
            }
        }
        const int i0 = t - YCGE_WF_L * r, i = i0 - cx;
        const bool active = row_ok && i >= 0 && i < ws_c;
        const bool reg_next = row_reg && wf_step_regular(g, i0 + 1);
        const bool lean = pipe_valid; // this step's pixels were prepared in the step before
        const float *rec_t = reinterpret_cast<const float *>(s_rec[t & (YCGE_WF_DEPTH - 1)]) + c * 4;        // this step's records of my chain
        if (lean || reg_next) {
            // ---- pipelined path, ONE basic block so that its three dependency chains interleave:
            // (1) one colour weight per lane.  Lanes 9 and 11: this step's pixel (values of the step before); the other lanes:
            //     the NEXT step's pixel, whose filtered taps but 9 and 11 are in the history already;
            // (2) this step's pixel: the partial sum prepared one step ago + slots 9, 10, 11 + the 13 unfiltered taps, publish;
            // (3) the next step's pixel: ordered partial sum of slots 0..8, term of slot 10.
            const float *rec_n = reinterpret_cast<const float *>(s_rec[(t + 1) & (YCGE_WF_DEPTH - 1)]) + c * 4;
            const float *rec_m = my_d ? rec_n : rec_t;
            const float4 R = *reinterpret_cast<const float4 *>(rec_m + my_slot * RS);
            const float c0w = rec_m[25 * RS + 3];
            const float lt = hist_f[my_rowoff + ((i + my_d + my_kx) & (YCGE_WF_RING - 1)) * 4];
            const float wc = exp_nonpos(neg_div<FAST>(fabsf(lt - c0w), a.dc, a.rc));
            const float Wm = my_wB * wc * R.x * R.y * R.z; // the reference's product order wBase*wc*wn*wz*wa (:699)
            // the next pixel may take this path if none of its 12 filtered taps is skipped (sky edge): all their records are "filtered tap"
            const bool clean = __float_as_int(rec_n[my_slot * RS + 3]) < 0;
            const bool next_valid = reg_next && __all_sync(FULL, clean || s >= 12);
            const float *win = chan_f + geo_row_f + ((i - 2) & (YCGE_WF_RING - 1)) * 4; // entry (i - 2) of the row two above; slot k at + (ky + 2) rows + (kx + 2) entries
            const float *wnx = chan_f + geo_row_f + ((i - 1) & (YCGE_WF_RING - 1)) * 4; // the same for the next pixel
            float acc = P;
            acc = acc + win[HR * 1 + 4 * 4] * __shfl_sync(FULL, Wm, hbase | 9);
            acc = acc + t10;
            acc = acc + win[HR * 2 + 1 * 4] * __shfl_sync(FULL, Wm, hbase | 11);
#pragma unroll
            for (int k = 12; k < 25; k++) acc = acc + rec_t[k * RS + q];
            publish(acc, rec_t, i, lean);
            float p = 0.0f;
#pragma unroll
            for (int k = 0; k < 9; k++) p = p + wnx[HR * (k / 5) + (k % 5) * 4] * __shfl_sync(FULL, Wm, hbase | k);
            P = p;
            t10 = wnx[HR * 2] * __shfl_sync(FULL, Wm, hbase | 10);
            pipe_valid = next_valid;
        } else pipe_valid = false;
        if (!lean && __any_sync(FULL, active)) {
            // ---- generic path (image border columns, rows 0..3 and H-1, sky edges): every filtered tap at the address the
            // pre-pass recorded, everything of this step's pixel evaluated now
            const float c0w = rec_t[25 * RS + 3];
            const float4 R = *reinterpret_cast<const float4 *>(rec_t + my_slot * RS);
            const int code = __float_as_int(R.w);
            const float lt = hist_f[min(code & 0x3FF0, YCGE_WF_HIST_FLOATS * 4 - 16) / 4 + 3];
            const float wc = exp_nonpos(neg_div<FAST>(fabsf(lt - c0w), a.dc, a.rc));
            const float Wm = my_wB * wc * R.x * R.y * R.z;
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 25; k++) {
                const int src_lane = k < 12 ? k : (k == 15 ? 12 : (k == 16 ? 13 : (k == 20 ? 14 : 15)));
                const bool may_be_new = k < 12 || k == 15 || k == 16 || k == 20 || k == 21;
                const float *rk = rec_t + k * RS;
                if (may_be_new) {
                    const float Wk = __shfl_sync(FULL, Wm, hbase | src_lane);
                    const int ck = __float_as_int(rk[3]);
                    const bool is_new = ck < 0;
                    const float v = is_new ? chan_f[min(ck & 0x3FF0, YCGE_WF_HIST_FLOATS * 4 - 16) / 4] : rk[q];
                    acc = acc + (is_new ? v * Wk : v);
                } else acc = acc + rk[q];
            }
            publish(acc, rec_t, i, active);
        }
        // ---- records of step t + DEPTH - 1 into the ring slot step t - 1 has left; the groups of steps t + 1 and t + 2 must have landed
        {
            const int ts = min(max(t + YCGE_WF_DEPTH - 1, 0), g.nt - 1); // out of range: a harmless reload
            const float4 *src = rec_g + (size_t)ts * (YCGE_WF_SLOTS * YCGE_WF_CHAINS);
            float4 *dst = s_rec[(t + YCGE_WF_DEPTH - 1) & (YCGE_WF_DEPTH - 1)];
#pragma unroll
            for (int p = tid; p < YCGE_WF_SLOTS * YCGE_WF_CHAINS; p += YCGE_WF_ROWS * 32) cp_async16(dst + p, src + p);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(YCGE_WF_DEPTH - 3) : "memory");
        }
        __syncthreads();
    }
    if (a.trace && tid == 0) { unsigned long long tm; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm)); a.trace[2 * band + 1] = tm; }
}

} // namespace ycge
