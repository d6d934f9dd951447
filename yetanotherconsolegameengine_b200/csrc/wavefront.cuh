// wavefront.cuh — K3'': the in-place à-trous iteration (RaytraceRenderer.cs:651-719 with cur == dst, stride 2) as a
// systolic wavefront (sm_100a).  Geometry, schedule and addressing: wavefront_layout.h (read that first).
//
//   atrous_wave_pre_kernel   one thread per pixel, fully parallel: 26 records per pixel, written where the band that owns
//                            the pixel will stream them from — a tap that FOLLOWS the pixel in row-major order (unfiltered
//                            input) becomes its finished weighted term, a tap that PRECEDES it (filtered value, not known
//                            yet) becomes its three guide weights plus the ADDRESS of that value in the band's history
//                            rings.  Clamping at the image border, rows 0 / H-1 (where taps of other kernel rows fold
//                            onto the pixel's own row) and sky are resolved here, per tap: the wavefront kernel has no
//                            border cases and ONE code path.
//   atrous_wave_kernel       ONE WARP per band of 4 rows x 2 chains, a quad of lanes per chain; all chains of a band advance
//                            one pixel per step in lock step (i = t - 3 r - cx), so that everything a pixel needs from its
//                            own band is in the shared-memory history by construction: no flags, no polling, no memory
//                            round trip, no block-wide barrier inside a band (only __syncwarp).  Per step
//                              1. lane q of a quad turns the records of slots 3q .. 3q+2 into terms: for a filtered tap the
//                                 colour weight exp(-|dlum| / cPhi) from the history and the product wBase*wc*wn*wz*wa in the
//                                 reference's order (:699); the term replaces the record in shared memory;
//                              2. every lane of the quad adds the 25 terms of its chain in the reference's ky-major / kx order
//                                 (packed FADD2), normalises (:706-714), takes the luma;
//                              3. lane 0 of the quad publishes the pixel: history ring + L2 (+ the peer GPU's buffer).
//                            The records arrive by cp.async DEPTH - 1 steps ahead; the two rows above the band (another warp's,
//                            an earlier launch's or a peer GPU's output) are read from L2 two steps ahead of their commit to
//                            the history and validated against the all-ones sentinel the output buffer is pre-filled with.
//   Bands take tickets in dispatch order; a band only ever waits for lower tickets, i.e. for warps that are running or
//   done, so the kernel needs no co-residency guarantee and shares the GPU with other frames' kernels.  Every poll has a
//   bound: a value that does not arrive sets *err and the frame fails with YCGE_ERR_CUDA instead of hanging the GPU.
#pragma once
#include "post.cuh"
#include "wavefront_layout.h"

namespace ycge {

__device__ __forceinline__ float wf_kw(int k) { return k == 0 ? 3.f / 8.f : ((k == 1 || k == -1) ? 1.f / 4.f : 1.f / 16.f); }

// shared-memory history: (ROWS + 2) rows x 2 chains x RING entries of (r, g, b, luma); a record names an entry by its byte offset
__host__ __device__ __forceinline__ int wf_history_offset(int entry) { return entry * 16; } // logical entry -> byte offset
#define YCGE_WF_BLOCK (YCGE_WF_SLOTS * YCGE_WF_CHAINS) // float4 records of one (band, step)

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
    float4 *out = a.rec + ((size_t)pl.warp * (size_t)a.g.nt + (size_t)pl.step) * YCGE_WF_BLOCK; // the (band, step) block
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
            const int slot = (ky + 2) * 5 + (kx + 2);
            out[slot * YCGE_WF_CHAINS + wf_record_column(slot, pl.chain)] = v;
        }
    }
    out[25 * YCGE_WF_CHAINS + wf_record_column(25, pl.chain)] = c0;
}

struct WaveArgs {
    const float4 *rec;
    float4 *new_;      // pass output, pre-filled with the sentinel on the rows this pass (or a peer) produces
    WfGeom g;
    float dc, rc;      // max(1e-6, cPhi) and its reciprocal
    unsigned int *ticket;
    unsigned int ticket_base;
    int *err;          // mapped host memory: set to 1 when a poll gives up
    unsigned long long *trace; // development aid: per band 32 words: globaltimer at the first step, after the last one, at every 64th step; or NULL
    // multi-GPU: rows [peer_y0, peer_y1) are ALSO stored into the output buffer of the rank below (see post.cuh)
    float4 *peer_new;
    int peer_y0, peer_y1;
    const int *ready;
    int frame;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
#define YCGE_WF_POLL_LIMIT (1 << 21) // polls of the peer's `ready` word (system-scope loads over NVLink: seconds)
// A band that waits for rows another kernel or another GPU has yet to produce gives up after this many nanoseconds WITHOUT
// ANY PROGRESS (globaltimer): ranks driven by separate host processes may start their kernels far apart -- a count of polls
// (an earlier form: 2^21 polls of 0.15 us) expired during such a gap on 4 GPUs and the frame was silently wrong.
#define YCGE_WF_WAIT_NS 20000000000ull
__device__ __forceinline__ unsigned long long wf_now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }


// one record -> one term (see the header).  R = the record, hv = the history entry it names (anything for a finished term),
// lum0 = luma of the centre, dc, rc = max(1e-6, cPhi) and its reciprocal
__device__ __forceinline__ float4 wf_term_wc(const float4 R, const float4 hv, const float wc, const float wB) {
    const float w = wB * wc * R.x * R.y * R.z; // the reference's product order wBase*wc*wn*wz*wa (:699)
    float4 t; // opaque select (every lane runs the exp chain: a compiler-made branch around it would split the block)
    asm("{ .reg .pred p; setp.lt.s32 p, %8, 0; selp.f32 %0, %4, %9, p; selp.f32 %1, %5, %10, p; selp.f32 %2, %6, %11, p; selp.f32 %3, %7, %12, p; }"
        : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
        : "f"(hv.x * w), "f"(hv.y * w), "f"(hv.z * w), "f"(w), "r"(__float_as_int(R.w)), "f"(R.x), "f"(R.y), "f"(R.z), "f"(R.w));
    return t;
}
template <bool FAST> __device__ __forceinline__ float wf_wc(const float4 hv, const float lum0, const float dc, const float rc) { return ycge_expf_nonpos(neg_div<FAST>(fabsf(hv.w - lum0), dc, rc)); }
template <bool FAST> __device__ __forceinline__ float4 wf_term(const float4 R, const float4 hv, const float lum0, const float wB, const float dc, const float rc) {
    const float wc = wf_wc<FAST>(hv, lum0, dc, rc);
    const float w = wB * wc * R.x * R.y * R.z; // the reference's product order wBase*wc*wn*wz*wa (:699)
    float4 t; // opaque select (every lane runs the exp chain: a compiler-made branch around it would split the block)
    asm("{ .reg .pred p; setp.lt.s32 p, %8, 0; selp.f32 %0, %4, %9, p; selp.f32 %1, %5, %10, p; selp.f32 %2, %6, %11, p; selp.f32 %3, %7, %12, p; }"
        : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
        : "f"(hv.x * w), "f"(hv.y * w), "f"(hv.z * w), "f"(w), "r"(__float_as_int(R.w)), "f"(R.x), "f"(R.y), "f"(R.z), "f"(R.w));
    return t;
}
#define YCGE_WF_HIST_MASK 0x1FF0 // a record names a history entry by its byte offset: sign bit | offset; a finished term reads any entry
static_assert(YCGE_WF_HISTORY_ENTRIES * 16 <= YCGE_WF_HIST_MASK + 16, "history must cover every offset the mask lets through");

template <bool V> struct WfTag { static constexpr bool value = V; };
// ---- thread-block clusters: consecutive bands of one row parity sit in one cluster and hand their last two rows over
// through distributed shared memory instead of an L2 round trip: the band stores each pixel of its rows 2 and 3, tagged with
// its index, into a staging ring of the band below, whose halo warp polls that ring in its OWN shared memory.  No fence
// anywhere (a release / acquire pair at cluster scope compiles to MEMBAR.ALL.GPU resp. CCTL.IVALL per step: measured 2x
// slower than no clusters): a staging entry is the all-ones sentinel until all four words of the pixel have landed and
// carries the pixel index instead of the luma (recomputed on arrival), so it validates itself.
// MEASURED (B200, 1080p, bit-identical, tests green): the hand-off inside a cluster takes 0.4 us instead of 1.8 us, but every
// step of every band gets slower (0.39 -> 0.45 us: the remote stores go through the generic path, and a cluster is packed
// onto few SMs of one GPC, so two ACTIVE bands share an SM's shared-memory pipeline where the plain launch pairs bands
// that are 148 tickets apart and never active together): 1.34 ms with clusters of 8, 1.36 of 4, 1.41 of 2 against 1.27 ms
// without.  The form is kept opt-in (YCGE_WAVE_CLUSTER=8) and off by default.
__device__ __forceinline__ unsigned int cluster_rank() { unsigned int r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned int map_to_rank(unsigned int sa, unsigned int rank) { unsigned int ra; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sa), "r"(rank)); return ra; }
__device__ __forceinline__ int ld_remote_s32(unsigned int ra) { int v; asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(ra) : "memory"); return v; }
__device__ __forceinline__ void st_remote_f4(unsigned int ra, float4 v) { asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(ra), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
__device__ __forceinline__ void st_remote_s32(unsigned int ra, int v) { asm volatile("st.relaxed.cluster.shared::cluster.s32 [%0], %1;" ::"r"(ra), "r"(v) : "memory"); }
__device__ __forceinline__ float4 lds_volatile_f4(unsigned int sa) { float4 v; asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa) : "memory"); return v; }
#define YCGE_WF_SPIN_LIMIT (1u << 31) // shared-memory polls of a band's warp, all waits together (a minute): its halo warp gives up long before
#define YCGE_WF_PUSH_AHEAD 24 // steps a band may run ahead of the band below it in its cluster (whose 32-entry rings it writes)
#ifndef YCGE_WF_UNROLL
#define YCGE_WF_UNROLL 1 // steps per loop iteration; 2 and 8 (= DEPTH: every ring slot a constant) were measured: no difference, the step is bound by its dependency chain
#endif
// volatile shared-memory words by shared-window address (kept in a register by the caller: see keep_reg)
__device__ __forceinline__ int lds_volatile(unsigned int sa) { int v; asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(sa) : "memory"); return v; }
__device__ __forceinline__ void sts_volatile(unsigned int sa, int v) { asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(sa), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned int keep_reg(unsigned int v) { asm volatile("" : "+r"(v)); return v; } // the compiler would otherwise rebuild a shared-window address from %cluster_ctarank at every use

// The HALO WARP of a band (warp 1 of its CTA): brings the two rows above the band from L2 into the band's history, in the
// order and at the addresses the schedule defines (in "step s" the four loaders -- row h above the band x column parity --
// commit one pixel each), as early as the producer and the history rings allow, and announces in *ready the last step
// whose pixels are complete.  The band's own warp never touches global memory for them: it only compares *ready with its
// step counter.  Here: lane = 4 j + loader polls the pixel of step base + j with a STRONG load (a weak one, cp.async
// included, may be served from a stale copy of the line in the near L2 partition: measured -- once a band had read a
// line before its producer wrote it, every later request for it returned the sentinel), the longest complete prefix of
// the eight steps is committed, the window moves on.  YCGE_WF_POLLS such windows are in flight, each in its own
// registers (the loop is unrolled over them), re-issued as soon as they are examined: a pixel is seen a quarter of an L2
// round trip after it became visible, on average, and no instruction ever touches a value that is still in flight except
// the one that has waited longest.
#define YCGE_WF_POLLS 4
__device__ __forceinline__ void wf_halo_warp(const WaveArgs &a, const WfGeom &g, const int yb0, const int lane, float4 *s_hist, const unsigned int ready, const unsigned int runner_t, const int nt, const int trace_band) {
    const unsigned int FULL = 0xffffffffu;
    const int L = lane & 3, j = lane >> 2, lh = L >> 1, lcx = L & 1;
    const int hy = wf_halo_row(yb0, lh);
    const bool loader = hy < yb0; // at the top of the image the rows above fold onto the band's own rows: no halo
    const int tl0 = lcx - YCGE_WF_L * (2 - lh); // loader L commits pixel ih = s - tl0 in step s
    const unsigned int n_halo = loader ? (unsigned int)g.ws[lcx] : 0u;
    const float4 *halo_p = a.new_ + (size_t)hy * g.W + lcx; // + 2 ih
    float4 *halo_hist = s_hist + (lh * 2 + lcx) * YCGE_WF_RING;
    int base = -YCGE_WF_LEAD, idle = 0;
    unsigned long long t_progress = wf_now_ns();
    float4 v[YCGE_WF_POLLS];
    int bs[YCGE_WF_POLLS];
    auto issue = [&](float4 &vv, int &b0) {
        b0 = base;
        const int ih = base + j - tl0;
        vv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if ((unsigned int)ih < n_halo) vv = ld_relaxed_f4(halo_p + 2 * ih);
    };
#pragma unroll
    for (int k = 0; k < YCGE_WF_POLLS; k++) issue(v[k], bs[k]);
    for (;;) {
#pragma unroll
        for (int k = 0; k < YCGE_WF_POLLS; k++) {
            // the window issued longest ago: lane (j, L) holds the pixel loader L commits in step bs + j
            const int s = bs[k] + j, ih = s - tl0;
            const bool mine = (unsigned int)ih < n_halo;
            const bool ok = !mine || f4_valid(v[k]);
            unsigned int m = __ballot_sync(FULL, ok); // bit 4 j + L
            m &= m >> 1; m &= m >> 2;                 // bit 4 j: step bs + j complete
            const int shift = base - bs[k];           // steps of this window that are committed already
            // steps the history rings have room for: the band's own warp is in step rt and still reads pixels down to rt - 6
            const int rt = lds_volatile(runner_t);
            const int room = min(min(rt + YCGE_WF_AHEAD, nt - 1) - base + 1, 8 - shift);
            int n = 0;
            if (shift < 8) {
                const unsigned int miss = (~m & 0x11111111u) >> (4 * shift); // bit 4 i: step base + i incomplete
                n = miss ? (__ffs((int)miss) - 1) >> 2 : 8;
                n = min(n, room);
            }
            if (mine && s >= base && s < base + n) halo_hist[ih & (YCGE_WF_RING - 1)] = v[k];
            if (n > 0) {
                __syncwarp();
                __threadfence_block();
#ifdef YCGE_WF_TRACE
                if (a.trace && lane == 0 && base <= 500 && base + n > 500) { unsigned long long tm; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm)); a.trace[32 * trace_band + 21] = tm; } // development aid
#endif
                base += n;
                if (lane == 0) sts_volatile(ready, base - 1);
                idle = 0;
                if (base >= nt) return;
            } else if (room <= 0) { __nanosleep(100); idle = 0; } // the band's own warp has to move first
            else if (shift < 8 && (++idle & 1023) == 0 && (idle == 1024 ? (t_progress = wf_now_ns(), false) : wf_now_ns() - t_progress > YCGE_WF_WAIT_NS)) { // the producer is gone: fail the frame, let the band run out
                if (lane == 0) { *(volatile int *)a.err = 1; sts_volatile(ready, nt); }
                return;
            }
            issue(v[k], bs[k]);
        }
    }
}

// The halo warp of a band whose rows above come through the staging ring (the band above sits in the same cluster): the same
// window of eight steps, polled in the band's own shared memory.  An entry is valid when none of its words is the sentinel
// and its tag is the pixel index expected at that place; it is committed with its luma recomputed (:269-272, the same
// three products and two sums as the producer's) and set back to the sentinel for the pixel 32 places on.
__device__ __forceinline__ void wf_halo_warp_staged(const WaveArgs &a, const WfGeom &g, const int lane, float4 *s_hist, float4 *s_stage, const unsigned int ready, const unsigned int runner_t, const int nt) {
    const unsigned int FULL = 0xffffffffu;
    const int L = lane & 3, j = lane >> 2, lh = L >> 1, lcx = L & 1;
    const int tl0 = lcx - YCGE_WF_L * (2 - lh);
    const unsigned int n_halo = (unsigned int)g.ws[lcx];
    float4 *halo_hist = s_hist + (lh * 2 + lcx) * YCGE_WF_RING;
    float4 *stage = s_stage + L * YCGE_WF_RING;
    const unsigned int stage_sa = (unsigned int)__cvta_generic_to_shared(stage);
    const float sv = __uint_as_float(YCGE_SENTINEL);
    int base = -YCGE_WF_LEAD, idle = 0;
    unsigned long long t_progress = 0;
    while (base < nt) {
        const int rt = lds_volatile(runner_t);
        const int room = min(min(rt + YCGE_WF_AHEAD, nt - 1) - base + 1, 8);
        const int ih = base + j - tl0;
        const bool mine = (unsigned int)ih < n_halo && j < room;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (mine) v = lds_volatile_f4(stage_sa + (unsigned int)(ih & (YCGE_WF_RING - 1)) * 16u);
        const bool ok = !mine || (f4_valid(v) && __float_as_int(v.w) == ih);
        unsigned int m = __ballot_sync(FULL, ok);
        m &= m >> 1; m &= m >> 2;
        const unsigned int miss = ~m & 0x11111111u;
        const int n = min(room, miss ? (__ffs((int)miss) - 1) >> 2 : 8);
        if (mine && j < n) {
            halo_hist[ih & (YCGE_WF_RING - 1)] = make_float4(v.x, v.y, v.z, luma3(v.x, v.y, v.z));
            stage[ih & (YCGE_WF_RING - 1)] = make_float4(sv, sv, sv, sv);
        }
        if (n > 0) {
            __syncwarp();
            __threadfence_block();
            base += n;
            if (lane == 0) sts_volatile(ready, base - 1);
            idle = 0;
        } else if (room <= 0) __nanosleep(100);
        else if ((++idle & 1023) == 0 && (idle == 1024 ? (t_progress = wf_now_ns(), false) : wf_now_ns() - t_progress > YCGE_WF_WAIT_NS)) { // the band above is gone: fail the frame, let the band run out
            if (lane == 0) { *(volatile int *)a.err = 1; sts_volatile(ready, nt); }
            return;
        } else __nanosleep(20);
    }
}

// CL = CTAs per cluster (1: no cluster).  A cluster takes ONE ticket: ticket k = row parity k & 1, bands (k >> 1) * CL + rank.
template <bool FAST, bool PEER, int CL> __global__ void __launch_bounds__(64) atrous_wave_kernel(WaveArgs a) {
    __shared__ __align__(128) float4 s_rec[YCGE_WF_DEPTH][YCGE_WF_BLOCK];
    __shared__ __align__(16) float4 s_hist[(YCGE_WF_HIST_MASK + 16) / 16];
    __shared__ __align__(16) float4 s_stage[4 * YCGE_WF_RING]; // [row h above x column parity][pixel & 31]: (r, g, b, pixel index) from the band above, or the sentinel
    __shared__ int s_band, s_ready, s_runner_t, s_below_t;
    const int lane = threadIdx.x & 31, c = lane >> 2, q = lane & 3, r = c >> 1, cx = c & 1;
    const unsigned int rank = CL > 1 ? cluster_rank() : 0u;
    if (threadIdx.x == 0) { if (rank == 0) s_band = (int)(atomicAdd(a.ticket, 1u) - a.ticket_base); s_ready = -YCGE_WF_LEAD - 1; s_runner_t = 0; s_below_t = 0; }
    for (int k = threadIdx.x; k < (YCGE_WF_HIST_MASK + 16) / 16; k += 64) s_hist[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (CL > 1) for (int k = threadIdx.x; k < 4 * YCGE_WF_RING; k += 64) { const float sv = __uint_as_float(YCGE_SENTINEL); s_stage[k] = make_float4(sv, sv, sv, sv); }
    __syncthreads();
    int ticket = s_band;
    if (CL > 1) { cluster_sync_all(); ticket = ld_remote_s32(map_to_rank((unsigned int)__cvta_generic_to_shared(&s_band), 0u)); }
    const WfGeom &g = a.g;
    const int cy = ticket & 1, b = (ticket >> 1) * CL + (int)rank, band = 2 * b + cy;
    const bool valid = band < g.n_warps && b < g.nb[cy];
    const int yb0 = g.yf[cy] + 2 * YCGE_WF_ROWS * b;
    const int nt = g.nt;
    const unsigned int ready_s = keep_reg((unsigned int)__cvta_generic_to_shared(&s_ready)), runner_s = keep_reg((unsigned int)__cvta_generic_to_shared(&s_runner_t));
    const bool pushed = CL > 1 && rank > 0 && valid;                                 // the band above (same cluster) stores my two halo rows into my history
    const bool pushes = CL > 1 && rank + 1 < CL && valid && b + 1 < g.nb[cy];        // ... and I do that for the band below
    if (!valid || threadIdx.x >= 32) {
        if (valid && !pushed) wf_halo_warp(a, g, yb0, lane, s_hist, ready_s, runner_s, nt, band);
        if (valid && pushed) wf_halo_warp_staged(a, g, lane, s_hist, s_stage, ready_s, runner_s, nt);
        if (CL > 1) cluster_sync_all(); // nobody leaves while a neighbour may still store into its shared memory
        return;
    }
    const int y = yb0 + 2 * r;
    const bool row_ok = y < g.y1;
    const bool last_row_band = yb0 + 2 * (YCGE_WF_ROWS - 1) >= g.H - 1 && g.y1 == g.H; // row H-1 folds the kernel rows below onto itself: slots 15, 16, 20, 21 may be filtered taps
    const char *hist_b = reinterpret_cast<const char *>(s_hist);
    // my three slots 3q + j, their kernel weights and their byte offsets in a block
    float wB[3];
    int off[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int k = 3 * q + j;
        wB[j] = wf_kw(k % 5 - 2) * wf_kw(k / 5 - 2);
        off[j] = (k * YCGE_WF_CHAINS + wf_record_column(k, c)) * 16;
    }
    // Row H-1 folds the kernel rows below onto itself: its slots 15, 20 (resp. 16, 21) are filtered taps with the SOURCE PIXEL
    // of slot 10 (resp. 11) -- same centre, hence the same colour weight, guide weights and history entry, only the kernel
    // weight differs.  Lane 3 of that chain's quad (it owns slots 10 and 11) writes those four terms as well.
    const bool fold_lane = last_row_band && q == 3 && y == g.H - 1;
    int off_fold[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const int sl = k == 0 ? 15 : (k == 1 ? 20 : (k == 2 ? 16 : 21)); off_fold[k] = (sl * YCGE_WF_CHAINS + wf_record_column(sl, c)) * 16; }
    const float wB_fold[4] = {wf_kw(-2) * wf_kw(1), wf_kw(-2) * wf_kw(2), wf_kw(-1) * wf_kw(1), wf_kw(-1) * wf_kw(2)}; // slots 15, 20, 16, 21
    const int off_c0 = (25 * YCGE_WF_CHAINS + wf_record_column(25, c)) * 16;
    int col[4]; // byte offset of my chain's column in the slots k with (k / 3) & 3 == m
#pragma unroll
    for (int m = 0; m < 4; m++) col[m] = ((c + 2 * m) & (YCGE_WF_CHAINS - 1)) * 16;
    // publishing lane of the quad: pixel i = t - t0 of the chain while 0 <= i < n_pub
    const int t0 = YCGE_WF_L * r + cx;
    const unsigned int n_pub = (q == 0 && row_ok) ? (unsigned int)g.ws[cx] : 0u;
    float4 *out_p = a.new_ + (size_t)y * g.W + cx; // + 2 i
    float4 *own_hist = s_hist + ((r + 2) * 2 + cx) * YCGE_WF_RING;
    // cluster hand-off: rows 2 and 3 of this band are the rows h = 0, 1 above the next one
    const unsigned int below_s = keep_reg((unsigned int)__cvta_generic_to_shared(&s_below_t));
    unsigned int rem_hist = 0, rem_above_t = 0;
    if (CL > 1) {
        if (pushes) {
            rem_hist = map_to_rank((unsigned int)__cvta_generic_to_shared(s_stage + ((r >= 2 ? r - 2 : 0) * 2 + cx) * YCGE_WF_RING), rank + 1);
        }
        if (pushed) rem_above_t = map_to_rank(below_s, rank - 1);
    }
    const bool push_lane = pushes && r >= 2;
    float4 *peer_p = out_p;
    if (PEER) {
        if (a.peer_new && row_ok && y >= a.peer_y0 && y < a.peer_y1) peer_p = a.peer_new + (size_t)y * g.W + cx; // else: this row once more (no branch in the body)
        if (a.peer_new && lane == 0) { // the rank below must have reset its buffer for this frame before anything is stored into it
            int v, n = 0;
            do { asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(a.ready) : "memory"); } while (v < a.frame && ++n < YCGE_WF_POLL_LIMIT);
            if (v < a.frame) *(volatile int *)a.err = 1;
        }
        __syncwarp();
    }
    const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    // record ring: the block of step ts into ring slot ts & (DEPTH - 1); one commit group per call.  Past the last step the
    // last block is loaded once more (into a slot nobody reads): no branch in the body
    const float4 *rec_g = a.rec + (size_t)band * (size_t)nt * YCGE_WF_BLOCK + lane;
    const unsigned int ring_s = keep_reg((unsigned int)__cvta_generic_to_shared(&s_rec[0][0]) + lane * 16); // shared-window address of my 16 bytes of ring slot 0
    auto fetch = [&](int ts) {
        const float4 *src = rec_g + (size_t)min(ts, nt - 1) * YCGE_WF_BLOCK;
        const unsigned int dst = ring_s + (unsigned int)(ts & (YCGE_WF_DEPTH - 1)) * (YCGE_WF_BLOCK * 16);
#pragma unroll
        for (int p = 0; p < YCGE_WF_BLOCK; p += 32)
            if (p + 32 <= YCGE_WF_BLOCK || p + lane < YCGE_WF_BLOCK) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + p * 16), "l"(src + p) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll 1
    for (int ts = 0; ts < YCGE_WF_DEPTH - 1; ts++) fetch(ts);
    asm volatile("cp.async.wait_group %0;" ::"n"(YCGE_WF_DEPTH - 3) : "memory"); // the blocks of steps 0 and 1 have landed (mine: the warp barrier covers the other lanes')
    __syncwarp();
    const float dc = a.dc, rc = a.rc;
    // the records of my slots and the centre of the coming step, loaded one step ahead (their block has landed by then)
    float4 R[3], c0;
    {
        const char *blk = reinterpret_cast<const char *>(s_rec[0]);
        c0 = *reinterpret_cast<const float4 *>(blk + off_c0);
#pragma unroll
        for (int j = 0; j < 3; j++) R[j] = *reinterpret_cast<const float4 *>(blk + off[j]);
    }
    int rdy = lds_volatile(ready_s); // what the halo warp had announced a step ago: enough in the steady state, re-read otherwise
    int below = 0;                   // the step the band below (same cluster) has reached
    unsigned int spins = 0;          // every wait of this warp counts: past the limit nothing waits any more and the frame fails
    if (a.trace && lane == 0) {
        unsigned long long tm; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm)); a.trace[32 * band] = tm;
        unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); a.trace[32 * band + 23] = smid; // development aid: where the band runs
    }

    auto run = [&](auto last_tag) { // two copies of the loop: the band that holds row H-1 also writes that row's folded slots
    constexpr bool LASTROW = decltype(last_tag)::value;
    constexpr int U = LASTROW ? 1 : YCGE_WF_UNROLL; // steps per loop iteration: ring slots become constants, one back edge per U steps
#pragma unroll 1
    for (int tb = 0; tb < nt; tb += U) {
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int t = tb + u;
        if (U > 1 && t >= nt) break;
        // step t reads the halo pixels committed in the steps up to t - 1 (the common case falls through)
        while (rdy < t - 1 && spins < YCGE_WF_SPIN_LIMIT) { rdy = lds_volatile(ready_s); spins++; }
        if (CL > 1) {
            while (pushes && t - below > YCGE_WF_PUSH_AHEAD && spins < YCGE_WF_SPIN_LIMIT) { below = lds_volatile(below_s); spins++; } // the band below still reads what I would overwrite
            if (pushed && lane == 0) st_remote_s32(rem_above_t, t);
        }
#ifdef YCGE_WF_TRACE
        if (a.trace && lane == 0 && t == 501) { unsigned long long tm; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm)); a.trace[32 * band + 22] = tm; }
        if (a.trace && lane == 0 && t == 513) { unsigned long long tm; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm)); a.trace[32 * band + 20] = tm; } // step 512 (row 3 pixel 503, row 2 pixel 506) is published
        if (a.trace && lane == 0 && (t & 63) == 63 && (t >> 6) < 18) { unsigned long long tm; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm)); a.trace[32 * band + 2 + (t >> 6)] = tm; } // development aid
#endif
        if (lane == 0) sts_volatile(runner_s, t);
        fetch(t + YCGE_WF_DEPTH - 1); // into the ring slot step t - 1 has left
        char *blk = reinterpret_cast<char *>(s_rec[t & (YCGE_WF_DEPTH - 1)]);
        // ---- 1. my slots: record -> term, in place: three independent dependency chains
        float4 hv[3];
#pragma unroll
        for (int j = 0; j < 3; j++) hv[j] = *reinterpret_cast<const float4 *>(hist_b + (__float_as_int(R[j].w) & YCGE_WF_HIST_MASK));
        if (LASTROW) {
            float wc[3];
#pragma unroll
            for (int j = 0; j < 3; j++) wc[j] = wf_wc<FAST>(hv[j], c0.w, dc, rc);
#pragma unroll
            for (int k = 0; k < 4; k++) { // the folded slots of row H-1, from slot 10 (k < 2) or 11
                const int j = 1 + (k >> 1);
                const float4 tf = wf_term_wc(R[j], hv[j], wc[j], wB_fold[k]);
                if (fold_lane && __float_as_int(R[j].w) < 0) *reinterpret_cast<float4 *>(blk + off_fold[k]) = tf;
            }
#pragma unroll
            for (int j = 0; j < 3; j++) R[j] = wf_term_wc(R[j], hv[j], wc[j], wB[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 3; j++) R[j] = wf_term<FAST>(R[j], hv[j], c0.w, wB[j], dc, rc);
        }
#pragma unroll
        for (int j = 0; j < 3; j++) *reinterpret_cast<float4 *>(blk + off[j]) = R[j];
        __syncwarp();
        rdy = lds_volatile(ready_s);
        if (CL > 1) below = lds_volatile(below_s);
        // the records of the NEXT step (its block landed a step ago): their addresses into the history are ready when this step ends
        float4 c0n;
        {
            const char *nblk = reinterpret_cast<const char *>(s_rec[(t + 1) & (YCGE_WF_DEPTH - 1)]);
            c0n = *reinterpret_cast<const float4 *>(nblk + off_c0);
#pragma unroll
            for (int j = 0; j < 3; j++) R[j] = *reinterpret_cast<const float4 *>(nblk + off[j]);
        }
        // ---- 2. the 25 terms in the reference's order, normalise (:706-714), luma
        float4 acc = zero4;
#pragma unroll
        for (int k = 0; k < 25; k++) acc = add4_rn(acc, *reinterpret_cast<const float4 *>(blk + k * (YCGE_WF_CHAINS * 16) + col[(k / 3) & 3]));
        const float inv = rcp_rn<FAST>(acc.w);
        const bool okw = acc.w > 1e-8f;
        const float rr = okw ? acc.x * inv : c0.x, gg = okw ? acc.y * inv : c0.y, bb = okw ? acc.z * inv : c0.z;
        const float4 res = make_float4(rr, gg, bb, luma3(rr, gg, bb));
        // ---- 3. publish
        const int i = t - t0;
        if ((unsigned int)i < n_pub) {
            own_hist[i & (YCGE_WF_RING - 1)] = res;
            st_relaxed_f4(out_p + 2 * i, res);
            if (PEER) st_relaxed_sys_f4(peer_p + 2 * i, res);
            if (CL > 1 && push_lane) st_remote_f4(rem_hist + (unsigned int)(i & (YCGE_WF_RING - 1)) * 16u, make_float4(res.x, res.y, res.z, __int_as_float(i)));
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(YCGE_WF_DEPTH - 3) : "memory"); // the block of step t + 2 has landed
        __syncwarp();
        c0 = c0n;
    }
    }
    };
    if (last_row_band) run(WfTag<true>{}); else run(WfTag<false>{});
    if (spins >= YCGE_WF_SPIN_LIMIT && lane == 0) *(volatile int *)a.err = 1;
    if (lane == 0) sts_volatile(runner_s, nt + YCGE_WF_AHEAD); // lets the halo warp run out
    if (CL > 1 && lane == 0) {
        if (pushed) st_remote_s32(rem_above_t, nt + 2 * YCGE_WF_PUSH_AHEAD);
    }
    if (a.trace && lane == 0) { unsigned long long tm; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm)); a.trace[32 * band + 1] = tm; }
    if (CL > 1) cluster_sync_all();
}

} // namespace ycge
