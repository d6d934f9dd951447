// ycge_lib.cu — the C ABI of include/ycge.h: context, scene flattening into HBM, per-frame launch sequence.
// Product code: no CPU fallback.  Every entry point fails loudly (negative status + message) when CUDA is not usable.
#include "post.cuh"
#include "wavefront.cuh"
#include "trace_stream.cuh"
#include "bvh_build.hpp"
#include "bvh_device.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

using namespace ycge;

namespace {

thread_local std::string tl_error;

template <class T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count) {
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const std::vector<T> &v, cudaStream_t s) {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        e = cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return e;
        return cudaStreamSynchronize(s);
    }
};

struct MeshStore {
    DevBuf<PairNode> nodes;
    DevBuf<DevTri> tris;
    DevBuf<int> tri_id;
    TreeRoot root;
    ycge_material material;
    int n_tris = 0, n_pairs = 0;
    unsigned int sort_fallbacks = 0;
    // a device build in flight (ycge_mesh_build_device): root record and counters arrive in pinned memory behind `pending`
    struct Pending { TreeRoot root; DbCounters cnt; } *pending_host = nullptr;
    cudaEvent_t pending = nullptr;
    ~MeshStore() { if (pending) cudaEventDestroy(pending); if (pending_host) cudaFreeHost(pending_host); }
};
struct TextureStore {
    DevBuf<uchar4> px;
    int w = 0, h = 0;
};
struct VolumeStore {
    DevBuf<uint8_t> vox;
    DevBuf<uint8_t> occ; // occupancy byte per brick (voxel_occupancy_kernel)
    DevBuf<int> raw_mat, raw_meta; // kept until the scene (and so the material indices) is known
    std::vector<int> palette;
    int n_ids = 0, levels = 1, def = 0;
    bool packed = false;
    DevVolume dv;
};

} // namespace

struct YcgeGroup;
struct ycge_ctx {
    std::unique_ptr<YcgeGroup> group; // several GPUs behind this context (ycge_multi.inl); every other member then describes the frame only
    ~ycge_ctx();
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    ycge_params P;
    int fbW = 0, fbH = 0, ss = 1, W = 0, H = 0;
    int tile_row0 = 0, tile_rows = 0;
    bool sharded = false;

    // image planes
    DevBuf<float4> cur, gnd0, gnd1, gas0, gas1, hist, sa, sb;
    DevBuf<unsigned long long> chain_trace;
    DevBuf<unsigned char> db_scratch; // device BVH build (ycge_mesh_build_device)
    // FRONT-only frames (ycge_frame_front): the trace of a frame depends on nothing of the frame before it -- only its TAA does
    // -- so it runs on its own stream (one per frame parity) into its own radiance plane, guide set (three of them) and
    // counters, ordered by events: trace(f) waits for TAA(f-2), TAA(f) for trace(f) and, by stream order, TAA(f-1)
    bool front_ahead = true, front_ahead_force = false, front_ahead_ready = false;
    cudaStream_t trace_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev_trace_done[2] = {nullptr, nullptr}, ev_taa_done[2] = {nullptr, nullptr};
    DevBuf<float4> cur_b, gnd2, gas2;
    DevBuf<TraceCounters> counters_b;
    TraceCounters *counters_last = nullptr; // the counters of the frame traced last
    DevBuf<unsigned char> ansi;       // device ANSI byte stream (ycge_ansi_emit)
    DevBuf<unsigned int> ansi_rows;   // [rows] lengths, [rows] offsets, [1] total
    DevBuf<float4> pre; // in-place à-trous pass: 25 planes of per-tap precomputed terms / guide weights
    DevBuf<int2> prim;
    DevBuf<float> rays_dbg;
    DevBuf<float> logs;
    DevBuf<ExposureState> expo;
    DevBuf<ycge_cell> cells;
    DevBuf<TraceCounters> counters;
    DevBuf<TraceTotals> totals;
    int sw = 0, sh = 0;
    const float4 *denoised = nullptr;

    // scene
    std::map<int, std::unique_ptr<MeshStore>> meshes;
    std::map<int, std::unique_ptr<VolumeStore>> volumes;
    std::map<int, std::unique_ptr<TextureStore>> textures;
    DevBuf<DevTexture> s_textures;
    DevBuf<PairNode> s_nodes;
    DevBuf<int> s_leaf;
    DevBuf<DevObject> s_objects;
    DevBuf<float4> s_materials;
    DevBuf<DevMesh> s_meshes;
    DevBuf<DevVolume> s_volumes;
    DevBuf<DevLight> s_lights;
    DevScene ds;
    bool have_scene = false;

    // renderer state (RaytraceRenderer.cs:24-29,69; TemporalAA.cs:12-16; ToneMapper.cs:13)
    long long frame_counter = 0;
    float cam[3] = {0.0f, 1.0f, 0.0f};
    float yaw = 0.0f, pitch = 0.0f, fov = 45.0f;
    float last_cam[3] = {NAN, NAN, NAN}, last_yaw = NAN, last_pitch = NAN;
    bool taa_valid = false, force_reset = false;
    float snap_cam[3] = {0, 0, 0}, snap_yaw = 0, snap_pitch = 0; // snapshot of the frame in flight
    bool frame_open = false;
    bool want_stats = false;
    bool debug_rays = false;
    float ansi_th[5] = {0, 0, 0, 0, 0};
    int inplace_ctas_per_launch = 0, inplace_ctas_static = 0;
    bool trace_lean = false; // the scene holds no voxel grid, no texture and no transparent material: the smaller kernel variant (trace.cuh MODE bit 1)
    int trace_variant = 1;  // 0: one thread per pixel path (trace_kernel), 1: ray stream with lane refill (trace_stream_kernel)
    int stream_ctas = 0, stream_ctas_stats = 0;
    // resumable à-trous state of the frame in flight (a sharded tile pauses before every in-place pass so that the caller
    // can move the boundary rows between ranks)
    struct Denoise {
        float4 *phys[3] = {nullptr, nullptr, nullptr}; // logical buffers of the reference: 0 = src (TAA history), 1 = scratchA, 2 = scratchB
        int cur_id = 0, dst_id = 1, it = 0, K = 0, parity = 0, ty0 = 0, ty1 = 0;
        int halo_after[10] = {0};
        bool pending = false; // an in-place pass is prepared (pre-pass + sentinel done) and waits for ycge_frame_inplace
        bool early_reset = false; // the sentinel fill + ready signal of iteration 1 were issued at the start of the frame
        int pa = 0, pb = 0;   // its row range
    } dn;
    EdgeDiv edge_div;      // max(1e-6, phi) and reciprocals (RaytraceRenderer.cs:694-697)
    // frame pipelining across ranks (ycge_frame_stash / ycge_frame_finish_stashed): the tile's denoised rows and the
    // exposure samples of a frame are parked so that the next frames can start before the all-reduce of this one
    struct Stash { DevBuf<float4> den; DevBuf<float> logs; };
    std::vector<std::unique_ptr<Stash>> stash;
    bool chain_timed = false;
    cudaStream_t aux = nullptr; // side stream: the peer-storing rows of the wavefront run as a concurrent kernel
    cudaEvent_t e_fork = nullptr, e_join = nullptr;
    // peer hand-off (ycge_peer_attach)
    DevBuf<int> flags;                 // [0]: "the rank below has reset its buffer for frame N" (written by that rank)
    float4 *below_sa = nullptr, *below_sb = nullptr;
    int *above_flags = nullptr;
    bool peers = false, has_above = false, has_below = false;
    void *ipc_opened[3] = {nullptr, nullptr, nullptr};
    bool fast_div = false; // the FMA division sequence was verified against IEEE division for these four divisors
    volatile int *wave_err_host = nullptr; // pinned, mapped: set by a wavefront kernel whose poll gave up (a value that never arrived)
    int *wave_err_dev = nullptr;
    int wave_pad_smem = 72 * 1024;      // see the launch of atrous_wave_kernel
    int wave_cluster = 1; // bands per thread-block cluster of the systolic form: 1 = no clusters (default); YCGE_WAVE_CLUSTER=8 in the environment opts in
    bool use_wave = true;              // stride-2 in-place pass: true = systolic bands (wavefront.cuh, 1.33 ms at 1080p); false: one warp per chain (post.cuh, 2.2 ms); bit-identical
    DevBuf<unsigned int> tickets;      // [0]: plain wavefront launches, [1]: peer-storing launches (dispatch-order tickets)
    std::vector<unsigned int> ticket_bases = std::vector<unsigned int>(132, 0u); // host mirror of the counters
    unsigned int ticket_base_of(int peer, int slot) const { return ticket_bases[peer ? 1 : 2 * slot + 2]; }
    void ticket_advance(int peer, int slot, unsigned int n) { ticket_bases[peer ? 1 : 2 * slot + 2] += n; }

    // Frame pipelining on one GPU (ycge_pipeline_config, used by ycge_render_frames_async and ycge_submit_frame).
    // Between consecutive frames the path has exactly three dependencies: the TAA history + guides (frame N+1's TAA
    // reads what frame N's wrote), the exposure scalar (ToneMapper.aeExposure) and the output cells.  The à-trous
    // passes of frame N -- among them the latency-bound wavefront -- feed nothing of frame N+1's trace / TAA, because the
    // reference never writes the denoised image back into the history (RaytraceRenderer.cs:221-224).  So a frame is
    // cut into FRONT (trace, TAA, à-trous pass 0; serial on the ctx's stream), BACK (pre-pass, wavefront, later
    // passes, exposure samples; on the slot's own stream, concurrent with the next frames' fronts and backs) and FINISH
    // (ordered exposure sum, cells, optional D2H; on the slot's stream but chained frame to frame by an event).
    // Slot k = frame % n_slots owns the scratch pair, the pre-records, the samples and a guide set.
    struct Slot {
        DevBuf<float4> sa, sb, pre, gnd, gas;
        DevBuf<float> logs;
        cudaStream_t st = nullptr;
        cudaEvent_t front_done = nullptr, fin_done = nullptr, host_done = nullptr;
        ~Slot() { if (st) cudaStreamDestroy(st); if (front_done) cudaEventDestroy(front_done); if (fin_done) cudaEventDestroy(fin_done); if (host_done) cudaEventDestroy(host_done); }
    };
    // Frame-parallel BACK over ranks (ycge_back_*): a rank receives whole TAA'd frames (history + guides, assembled from every
    // rank's FRONT tile) into a back slot and runs the à-trous passes, the exposure samples and the FINISH of that frame.
    struct BackSlot {
        DevBuf<float4> hist, gnd, gas, sa, sb, pre;
        DevBuf<float> logs;
        DevBuf<ycge_cell> cells;
        const float4 *denoised = nullptr;
    };
    std::vector<std::unique_ptr<BackSlot>> back_slots;
    struct IoOverride { const float4 *gnd, *gas; DevBuf<float4> *pre; float *logs; size_t n_logs; cudaStream_t stream; int ticket_slot; };
    const IoOverride *io = nullptr;          // set for the duration of ycge_back_denoise
    std::vector<std::unique_ptr<Slot>> slots; // slots[k-1] for k >= 1; slot 0 is the ctx's own buffers and (unpipelined) stream
    cudaStream_t st0 = nullptr;               // slot 0's back stream when pipelining
    cudaEvent_t front_done0 = nullptr, fin_done0 = nullptr, host_done0 = nullptr;
    int n_slots = 1;
    bool pipelined = false;                   // set for the duration of a pipelined submission
    int cur_slot = 0, cur_gset = 0, last_gset = 0;
    cudaEvent_t last_fin = nullptr;           // FINISH of the most recently submitted pipelined frame
    cudaStream_t back = nullptr;              // stream of the frame in flight after its front part
    ycge_cell *host_out = nullptr;            // ycge_submit_frame: destination of the frame being submitted
    int host_stride = 0;
    std::vector<std::pair<long long, cudaEvent_t>> in_flight; // submitted frames whose cells have not been waited for

    // timing
    cudaEvent_t ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // [7],[8] bracket the wavefront kernel
    int launches_last = 0;
    std::string err;
};

namespace {

int fail(ycge_ctx *ctx, int code, const std::string &msg) {
    tl_error = msg;
    if (ctx) ctx->err = msg;
    return code;
}
// No C++ exception may cross the C ABI (a P/Invoke or ctypes host would abort): every exported function that allocates is a
// function-try-block ending in YCGE_CATCH.
#define YCGE_CATCH                                                                                                          \
    catch (const std::bad_alloc &) { return fail(nullptr, YCGE_ERR_LIMIT, "out of host memory"); }                          \
    catch (const std::exception &e) { return fail(nullptr, YCGE_ERR_INVALID, std::string("exception: ") + e.what()); }      \
    catch (...) { return fail(nullptr, YCGE_ERR_INVALID, "unknown exception"); }
int wave_check(ycge_ctx *c) { // after a wait: did a kernel of the frame(s) just finished raise the mapped "failed" flag?
    if (c->wave_err_host && *c->wave_err_host) {
        const int e = *c->wave_err_host;
        *c->wave_err_host = 0;
        if (e & 2) return fail(c, YCGE_ERR_LIMIT, "traversal stack overflow (tree deeper than the device stack): the frame is not valid");
        return fail(c, YCGE_ERR_CUDA, "in-place a-trous wavefront: a filtered value it waited for never arrived (poll limit reached; a peer or an earlier launch failed)");
    }
    return 0;
}
#define CK(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(ctx, YCGE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

// ---- slots (frame pipelining) and guide sets ------------------------------------------------------------------
struct SlotView { float4 *sa, *sb; DevBuf<float4> *pre; float *logs; size_t n_logs; cudaStream_t st; cudaEvent_t front_done, fin_done, host_done; };
SlotView slot_view(ycge_ctx *c, int k) {
    if (k == 0) return SlotView{c->sa.p, c->sb.p, &c->pre, c->logs.p, c->logs.n, c->st0, c->front_done0, c->fin_done0, c->host_done0};
    ycge_ctx::Slot &s = *c->slots[k - 1];
    return SlotView{s.sa.p, s.sb.p, &s.pre, s.logs.p, s.logs.n, s.st, s.front_done, s.fin_done, s.host_done};
}
int n_gsets(const ycge_ctx *c) { return c->front_ahead_ready ? 3 : std::max(2, c->n_slots); }
float4 *gnd_of(ycge_ctx *c, int g) { return g == 0 ? c->gnd0.p : (g == 1 ? c->gnd1.p : (c->front_ahead_ready ? c->gnd2.p : c->slots[g - 1]->gnd.p)); }
float4 *gas_of(ycge_ctx *c, int g) { return g == 0 ? c->gas0.p : (g == 1 ? c->gas1.p : (c->front_ahead_ready ? c->gas2.p : c->slots[g - 1]->gas.p)); }
// every wait for "the context's stream" also covers the trace streams of FRONT-only frames
cudaError_t sync_ctx_streams(ycge_ctx *c) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) if (c->trace_stream[k]) e = cudaStreamSynchronize(c->trace_stream[k]);
    return e;
}

// ---- reference SoA tree -> pair nodes (device_types.h) ------------------------------------------------------
struct TreeView {
    int n_nodes, root, n_leaf;
    const float *min_x, *min_y, *min_z, *max_x, *max_y, *max_z;
    const int32_t *left, *right, *start, *count, *leaf;
};
TreeView view_of(const FlatTree &t) {
    return TreeView{t.n_nodes(), t.root, (int)t.leaf_index.size(), t.min_x.data(), t.min_y.data(), t.min_z.data(), t.max_x.data(), t.max_y.data(), t.max_z.data(),
                    t.left.data(), t.right.data(), t.start.data(), t.count.data(), t.leaf_index.data()};
}
TreeView view_of(const ycge_bvh &b) {
    return TreeView{b.n_nodes, b.root, b.n_leaf_refs, b.min_x, b.min_y, b.min_z, b.max_x, b.max_y, b.max_z, b.left, b.right, b.start, b.count, b.leaf_index};
}
int flatten(ycge_ctx *ctx, const TreeView &t, std::vector<PairNode> &out, TreeRoot &root) {
    out.clear();
    memset(&root, 0, sizeof root);
    root.ref = YCGE_REF_NONE;
    if (t.root < 0 || t.n_nodes == 0) return 0;
    std::vector<int> pair_of(t.n_nodes, -1);
    int n_pairs = 0;
    for (int i = 0; i < t.n_nodes; i++) if (t.count[i] <= 0) pair_of[i] = n_pairs++;
    auto ref_of = [&](int i, int &ref) -> bool {
        if (i < 0) { ref = YCGE_REF_NONE; return true; }
        if (i >= t.n_nodes) return false;
        if (t.count[i] > 0) {
            if (t.count[i] > YCGE_LEAF_MAX_COUNT || t.start[i] < 0 || t.start[i] >= YCGE_LEAF_MAX_START || t.start[i] + t.count[i] > t.n_leaf) return false;
            ref = ~(((t.count[i] - 1) << 26) | t.start[i]);
        } else ref = pair_of[i];
        return true;
    };
    out.resize(n_pairs);
    for (int i = 0; i < t.n_nodes; i++) {
        if (t.count[i] > 0) continue;
        int l = t.left[i], r = t.right[i], lref, rref;
        if (!ref_of(l, lref) || !ref_of(r, rref)) return fail(ctx, YCGE_ERR_LIMIT, "BVH node out of range or leaf larger than 32 primitives");
        float lb[6] = {0, 0, 0, 0, 0, 0}, rb[6] = {0, 0, 0, 0, 0, 0};
        if (l >= 0) { lb[0] = t.min_x[l]; lb[1] = t.min_y[l]; lb[2] = t.min_z[l]; lb[3] = t.max_x[l]; lb[4] = t.max_y[l]; lb[5] = t.max_z[l]; }
        if (r >= 0) { rb[0] = t.min_x[r]; rb[1] = t.min_y[r]; rb[2] = t.min_z[r]; rb[3] = t.max_x[r]; rb[4] = t.max_y[r]; rb[5] = t.max_z[r]; }
        PairNode pn;
        pn.q0 = make_float4(lb[0], lb[1], lb[2], lb[3]);
        pn.q1 = make_float4(lb[4], lb[5], rb[0], rb[1]);
        pn.q2 = make_float4(rb[2], rb[3], rb[4], rb[5]);
        int zero = 0;
        float lf, rf, zf;
        memcpy(&lf, &lref, 4); memcpy(&rf, &rref, 4); memcpy(&zf, &zero, 4);
        pn.q3 = make_float4(lf, rf, zf, zf);
        out[pair_of[i]] = pn;
    }
    int rr;
    if (!ref_of(t.root, rr)) return fail(ctx, YCGE_ERR_LIMIT, "BVH root out of range");
    root.ref = rr;
    root.lo[0] = t.min_x[t.root]; root.lo[1] = t.min_y[t.root]; root.lo[2] = t.min_z[t.root];
    root.hi[0] = t.max_x[t.root]; root.hi[1] = t.max_y[t.root]; root.hi[2] = t.max_z[t.root];
    return 0;
}

// ---- ANSI thresholds: smallest binary32 c in [0,1] with LinearToSrgb8((double)c) >= bound --------------------
int linear_to_srgb8(double c) { // ANSITerminalRenderer.cs:298-307
    if (c < 0.0) c = 0.0;
    if (c > 1.0) c = 1.0;
    double s = c <= 0.0031308 ? 12.92 * c : 1.055 * std::pow(c, 1.0 / 2.4) - 0.055;
    int v = (int)std::nearbyint(s * 255.0);
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}
float ansi_threshold(int bound) {
    uint32_t lo = 0, hi = 0x3F800000u; // f(lo) < bound <= f(hi)
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        float c;
        memcpy(&c, &mid, 4);
        if (linear_to_srgb8((double)c) >= bound) hi = mid; else lo = mid;
    }
    float c;
    memcpy(&c, &hi, 4);
    return c;
}

// Priority of the BACK streams.  Measured at 1080p with 3 slots: equal priority 302 frames/s, highest priority 287 (the next
// frame's trace kernel then starves behind two resident wavefront kernels and the serial FRONT becomes the bottleneck).
int back_priority() {
    if (const char *e = getenv("YCGE_BACK_PRIORITY")) return atoi(e); // development aid
    return 0;
}

// (re)allocates the extra slots for the current geometry; the caller has synchronised every stream
int alloc_slots(ycge_ctx *c, int n) {
    c->slots.clear();
    c->n_slots = 1;
    const size_t px = (size_t)c->W * c->H;
    for (int k = 1; k < n; k++) {
        std::unique_ptr<ycge_ctx::Slot> sl(new ycge_ctx::Slot());
        CK(c, sl->sa.alloc(px)); CK(c, sl->sb.alloc(px));
        CK(c, cudaMemsetAsync(sl->sa.p, 0, px * sizeof(float4), c->stream));
        CK(c, cudaMemsetAsync(sl->sb.p, 0, px * sizeof(float4), c->stream));
        if (k >= 2) {
            CK(c, sl->gnd.alloc(px)); CK(c, sl->gas.alloc(px));
            CK(c, cudaMemsetAsync(sl->gnd.p, 0, px * sizeof(float4), c->stream));
            CK(c, cudaMemsetAsync(sl->gas.p, 0, px * sizeof(float4), c->stream));
        }
        CK(c, sl->logs.alloc((size_t)c->sw * c->sh));
        CK(c, cudaMemsetAsync(sl->logs.p, 0, sl->logs.n * sizeof(float), c->stream));
        CK(c, cudaStreamCreateWithPriority(&sl->st, cudaStreamNonBlocking, back_priority()));
        CK(c, cudaEventCreateWithFlags(&sl->front_done, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&sl->fin_done, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&sl->host_done, cudaEventDisableTiming));
        c->slots.push_back(std::move(sl));
    }
    c->n_slots = std::max(1, n);
    c->last_fin = nullptr;
    CK(c, sync_ctx_streams(c));
    return 0;
}

// back slots of the frame-parallel path (ycge_back_config); re-created by ycge_resize
int alloc_back_slots(ycge_ctx *c, int n_slots) { // the caller has synchronised the device
    c->back_slots.clear();
    const size_t px = (size_t)c->W * c->H;
    for (int k = 0; k < n_slots; k++) {
        std::unique_ptr<ycge_ctx::BackSlot> b(new ycge_ctx::BackSlot());
        CK(c, b->hist.alloc(px)); CK(c, b->gnd.alloc(px)); CK(c, b->gas.alloc(px)); CK(c, b->sa.alloc(px)); CK(c, b->sb.alloc(px));
        CK(c, b->pre.alloc(px * 25));
        CK(c, b->logs.alloc((size_t)c->sw * c->sh));
        CK(c, b->cells.alloc((size_t)c->fbW * c->fbH));
        CK(c, cudaMemset(b->hist.p, 0, px * sizeof(float4))); CK(c, cudaMemset(b->gnd.p, 0, px * sizeof(float4))); CK(c, cudaMemset(b->gas.p, 0, px * sizeof(float4)));
        CK(c, cudaMemset(b->sa.p, 0, px * sizeof(float4))); CK(c, cudaMemset(b->sb.p, 0, px * sizeof(float4)));
        CK(c, cudaMemset(b->logs.p, 0, b->logs.n * sizeof(float)));
        CK(c, cudaMemset(b->cells.p, 0, b->cells.n * sizeof(ycge_cell)));
        c->back_slots.push_back(std::move(b));
    }
    return 0;
}

int alloc_planes(ycge_ctx *c) {
    size_t n = (size_t)c->W * c->H;
    CK(c, c->cur.alloc(n)); CK(c, c->gnd0.alloc(n)); CK(c, c->gnd1.alloc(n)); CK(c, c->gas0.alloc(n)); CK(c, c->gas1.alloc(n));
    CK(c, c->hist.alloc(n)); CK(c, c->sa.alloc(n)); CK(c, c->sb.alloc(n)); CK(c, c->prim.alloc(n));
    int step = std::max(2, c->ss * 2);
    c->sw = (c->W + step - 1) / step;
    c->sh = (c->H + step - 1) / step;
    CK(c, c->logs.alloc((size_t)c->sw * c->sh));
    CK(c, c->cells.alloc((size_t)c->fbW * c->tile_rows));
    CK(c, cudaMemsetAsync(c->cur.p, 0, n * sizeof(float4), c->stream));
    CK(c, cudaMemsetAsync(c->hist.p, 0, n * sizeof(float4), c->stream));
    CK(c, cudaMemsetAsync(c->sa.p, 0, n * sizeof(float4), c->stream));
    CK(c, cudaMemsetAsync(c->sb.p, 0, n * sizeof(float4), c->stream));
    CK(c, cudaMemsetAsync(c->gnd0.p, 0, n * sizeof(float4), c->stream)); CK(c, cudaMemsetAsync(c->gnd1.p, 0, n * sizeof(float4), c->stream));
    CK(c, cudaMemsetAsync(c->gas0.p, 0, n * sizeof(float4), c->stream)); CK(c, cudaMemsetAsync(c->gas1.p, 0, n * sizeof(float4), c->stream));
    CK(c, cudaMemsetAsync(c->logs.p, 0, c->logs.n * sizeof(float), c->stream));
    CK(c, cudaMemsetAsync(c->cells.p, 0, c->cells.n * sizeof(ycge_cell), c->stream));
    c->taa_valid = false;
    c->denoised = nullptr;
    c->last_gset = 0;
    return alloc_slots(c, c->n_slots);
}

int set_geometry(ycge_ctx *c, int fb_w, int fb_h, int ss, int tile_row0, int tile_rows) {
    if (fb_w <= 0 || fb_h <= 0) return fail(c, YCGE_ERR_INVALID, "framebuffer size must be positive");
    c->fbW = fb_w; c->fbH = fb_h; c->ss = ss < 1 ? 1 : ss;
    c->W = c->fbW * c->ss; c->H = c->fbH * 2 * c->ss;
    if (tile_row0 < 0 || tile_row0 >= fb_h) return fail(c, YCGE_ERR_INVALID, "tile_row0 out of range");
    if (tile_rows <= 0) tile_rows = fb_h - tile_row0;
    if (tile_row0 + tile_rows > fb_h) return fail(c, YCGE_ERR_INVALID, "tile exceeds framebuffer");
    c->tile_row0 = tile_row0; c->tile_rows = tile_rows;
    c->sharded = !(tile_row0 == 0 && tile_rows == fb_h);
    int rc = alloc_planes(c);
    if (rc == 0 && c->front_ahead_ready) { // the extra planes of FRONT-only frames follow the geometry (the caller has synchronised the device)
        const size_t px = (size_t)c->W * c->H;
        CK(c, c->cur_b.alloc(px)); CK(c, c->gnd2.alloc(px)); CK(c, c->gas2.alloc(px));
        CK(c, cudaMemset(c->gnd2.p, 0, px * sizeof(float4))); CK(c, cudaMemset(c->gas2.p, 0, px * sizeof(float4)));
    }
    return rc;
}

void normalize3(float v[3]) { // Vec3.Normalized
    float l2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (l2 <= 0.0f) return;
    float inv = 1.0f / std::sqrt(l2);
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
}
void cross3h(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
float fracf_h(float v) { return v - std::floor(v); }

bool should_reset_history(const ycge_ctx *c) { // TemporalAA.cs:58-67
    float dx = c->snap_cam[0] - c->last_cam[0], dy = c->snap_cam[1] - c->last_cam[1], dz = c->snap_cam[2] - c->last_cam[2];
    float trans = (dx != dx) ? 0.0f : std::sqrt(dx * dx + dy * dy + dz * dz);
    float dyaw = (c->last_yaw != c->last_yaw) ? 0.0f : std::fabs(c->snap_yaw - c->last_yaw);
    float dpitch = (c->last_pitch != c->last_pitch) ? 0.0f : std::fabs(c->snap_pitch - c->last_pitch);
    float tr = c->P.motion_trans_reset > 0.0f ? c->P.motion_trans_reset : 0.0f;
    float rr = c->P.motion_rot_reset > 0.0f ? c->P.motion_rot_reset : 0.0f;
    return trans > tr || dyaw > rr || dpitch > rr;
}

inline int div_up(int a, int b) { return (a + b - 1) / b; }

// ---- the per-frame launch sequence ----------------------------------------------------------------------------
int denoise_run(ycge_ctx *c);
void halo_rows(const ycge_ctx *c, int &lo, int &a, int &slo, int &sa);
int frame_begin_impl(ycge_ctx *c, bool front_only = false) {
    if (!c->have_scene) return fail(c, YCGE_ERR_NO_SCENE, "Scene BVH not built; call ycge_scene_upload() after populating the scene");
    if (c->frame_open) return fail(c, YCGE_ERR_INVALID, "ycge_frame_begin called twice without ycge_frame_finish");
    { int rc = wave_check(c); if (rc) return rc; } // a wavefront kernel of an EARLIER frame gave up waiting (the phase API has no wait of its own: the flag is mapped host memory, reading it is free)
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int W = c->W, H = c->H, ss = c->ss;
    // camera snapshot (:161-169) and history decision (:171)
    memcpy(c->snap_cam, c->cam, sizeof c->cam); c->snap_yaw = c->yaw; c->snap_pitch = c->pitch;
    bool reset = should_reset_history(c) || c->force_reset || !c->taa_valid;
    c->force_reset = false;
    long long frame = ++c->frame_counter;                                   // :175
    int frameIdx = (int)(frame & 0x7fffffff);                              // :177
    FrameConsts fc;
    fc.jitter_rot_x = fracf_h((float)(frameIdx + 1) * 0.61803398875f);     // :178
    fc.jitter_rot_y = fracf_h((float)(frameIdx + 1) * 0.38196601125f);     // :179
    fc.rot0 = fracf_h((float)(frameIdx + 1) * 0.7548776662466927f);        // RaytraceSampler.cs:32
    fc.rot1 = fracf_h((float)(frameIdx + 1) * 0.5698402909980532f);
    float aspect = (float)W / (float)H;                                     // :159
    float fovRad = c->fov * (3.14159274f / 180.0f);                          // :428
    fc.half_h = ycge_tanf(0.5f * fovRad);
    fc.half_w = fc.half_h * aspect;
    float cp = ycge_cosf(c->snap_pitch);                                     // :413-417
    float fwd[3] = {ycge_sinf(c->snap_yaw) * cp, ycge_sinf(c->snap_pitch), -ycge_cosf(c->snap_yaw) * cp};
    normalize3(fwd);
    float worldUp[3] = {0.0f, 1.0f, 0.0f}, right[3], up[3];
    cross3h(fwd, worldUp, right); normalize3(right);
    cross3h(right, fwd, up); normalize3(up);
    for (int k = 0; k < 3; k++) { fc.cam[k] = c->snap_cam[k]; fc.fwd[k] = fwd[k]; fc.right[k] = right[k]; fc.up[k] = up[k]; }
    fc.frame = frame; fc.W = W; fc.H = H;

    // row ranges: the tile plus the halo each later pass needs (DESIGN.md "Multi-GPU")
    const int K = std::max(1, c->P.atrous_iterations);
    const int ty0 = c->tile_row0 * 2 * ss, ty1 = (c->tile_row0 + c->tile_rows) * 2 * ss;
    std::vector<int> halo_after(K + 1, 0); // halo_after[k] = rows still needed around the tile after pass k-1 (k=0: TAA output)
    if (!front_only) for (int k = K - 1; k >= 0; k--) halo_after[k] = halo_after[k + 1] + 2 * (1 << k); // FRONT only: the passes run elsewhere
    auto range = [&](int halo, int &a, int &b) { a = std::max(0, ty0 - halo); b = std::min(H, ty1 + halo); };

    // FRONT-only frames: the trace on its own stream (see the members)
    // Worth it where the tile's trace leaves the GPU idle (a third of the frame or less: 4 and 8 GPUs, +22 % / +50 % at 8); on
    // half-frame tiles the traces fill the GPU and overlapping them only adds contention (2 GPUs: 645 -> 571 frames/s)
    const bool ahead = front_only && c->front_ahead && !c->want_stats && !c->debug_rays && c->n_slots <= 1 &&
                       (c->front_ahead_force || c->front_ahead_ready || c->tile_rows * 3 <= c->fbH);
    if (ahead && !c->front_ahead_ready) {
        CK(c, cudaDeviceSynchronize());
        int lo = 0, hi = 0;
        CK(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int k = 0; k < 2; k++) {
            CK(c, cudaStreamCreateWithPriority(&c->trace_stream[k], cudaStreamNonBlocking, hi));
            CK(c, cudaEventCreateWithFlags(&c->ev_trace_done[k], cudaEventDisableTiming));
            CK(c, cudaEventCreateWithFlags(&c->ev_taa_done[k], cudaEventDisableTiming));
        }
        const size_t px = (size_t)W * H;
        CK(c, c->cur_b.alloc(px)); CK(c, c->gnd2.alloc(px)); CK(c, c->gas2.alloc(px)); CK(c, c->counters_b.alloc(1));
        CK(c, cudaMemset(c->gnd2.p, 0, px * sizeof(float4))); CK(c, cudaMemset(c->gas2.p, 0, px * sizeof(float4)));
        c->front_ahead_ready = true;
    }
    const int fp = (int)(frame & 1);
    cudaStream_t ts = ahead ? c->trace_stream[fp] : s;
    float4 *cur_plane = (ahead && fp) ? c->cur_b.p : c->cur.p;
    TraceCounters *cnt_dev = (ahead && fp) ? c->counters_b.p : c->counters.p;
    c->counters_last = cnt_dev;
    // slot and guide set of this frame; `parity` selects the guide set the trace kernel writes (img.gnd/gas[parity])
    const int slot = (int)(frame % c->n_slots), gset = (int)(frame % n_gsets(c)), gprev = c->last_gset == gset ? (gset + 1) % n_gsets(c) : c->last_gset;
    c->cur_slot = slot; c->cur_gset = gset;
    const SlotView sv = slot_view(c, slot);
    c->back = c->pipelined ? sv.st : s;
    if (c->pipelined) CK(c, cudaStreamWaitEvent(s, sv.fin_done, 0)); // the slot's previous frame has left its buffers (a never-recorded event does not wait)
    else if (c->last_fin) { CK(c, cudaStreamWaitEvent(s, c->last_fin, 0)); c->last_fin = nullptr; } // a synchronous frame after pipelined ones
    const int parity = 0;
    ImagePlanes img;
    img.cur = cur_plane; img.gnd[0] = gnd_of(c, gset); img.gnd[1] = gnd_of(c, gprev); img.gas[0] = gas_of(c, gset); img.gas[1] = gas_of(c, gprev);
    img.hist = c->hist.p; img.sa = sv.sa; img.sb = sv.sb; img.prim = c->prim.p; img.rays = c->debug_rays ? c->rays_dbg.p : nullptr;

    int launches = 0;
    if (ahead) CK(c, cudaStreamWaitEvent(ts, c->ev_taa_done[fp], 0)); // TAA of frame f-2 (and, by its stream's order, the copies of frame f-3) has left this radiance plane / guide set
    CK(c, cudaEventRecord(c->ev[0], ts));
    c->dn.early_reset = false;
    if (c->sharded && c->peers && K >= 2) {
        // Peer hand-off: the output buffer of the first in-place iteration (it = 1: OLD = scratchA, NEW = scratchB) is free
        // as soon as the previous frame's following pass has read it, i.e. now.  Resetting it and telling the rank above
        // right away (instead of after this frame's trace / TAA / pass 0 / pre-pass) lets that rank, which runs ahead in
        // the frame pipeline, finish its boundary rows without waiting for this rank's whole front end.
        int a1, b1; range(halo_after[2], a1, b1);
        const int m0 = c->has_above ? std::max(0, a1 - 4) : a1;
        CK(c, cudaMemsetAsync(sv.sb + (size_t)m0 * W, 0xFF, (size_t)(b1 - m0) * W * sizeof(float4), s));
        if (c->has_above) { peer_signal_kernel<<<1, 1, 0, s>>>(c->above_flags, (int)frame); launches++; }
        c->dn.early_reset = true;
    }
    CK(c, cudaMemsetAsync(cnt_dev, 0, sizeof(TraceCounters), ts));
    { // K1
        int a, b; range(halo_after[0] + 1, a, b);
        fc.y0 = a; fc.y1 = b;
        TraceParams tp;
        tp.diffuse_bounces = c->P.diffuse_bounces; tp.max_mirror_bounces = c->P.max_mirror_bounces; tp.max_refractions = c->P.max_refractions;
        tp.mirror_threshold = c->P.mirror_threshold; tp.eps = c->P.eps; tp.sigma_rad = c->P.diffuse_sigma_deg * (3.14159274f / 180.0f); // :460
        tp.seed_salt = c->P.seed_salt;
        tp.host_err = c->wave_err_dev;
        // Frames in flight on this GPU: the thread-per-path form, whose CTAs come and go, shares the SMs better with the resident
        // wavefront kernels of the previous frames than the persistent ray-stream form does (measured, 3 slots: 300 against 284
        // frames/s); a frame that has the GPU to itself takes the ray-stream form (1.09 against 1.18 ms).  Same results either way.
        const int variant = c->pipelined ? 0 : c->trace_variant;
        if (variant == 1) { // ray stream: persistent warps, as many CTAs as fit the GPU at once
            if (c->stream_ctas <= 0) {
                int occ = 0, occ_s = 0;
                cudaDeviceProp prop;
                CK(c, cudaGetDeviceProperties(&prop, c->device));
                CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trace_stream_kernel<0>, 128, 0));
                CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, trace_stream_kernel<1>, 128, 0));
                c->stream_ctas = std::max(1, occ * prop.multiProcessorCount);
                c->stream_ctas_stats = std::max(1, occ_s * prop.multiProcessorCount);
            }
            const int n_tiles = div_up(W, 8) * div_up(b - a, 4);
            const int refill_min = getenv("YCGE_STREAM_REFILL") ? atoi(getenv("YCGE_STREAM_REFILL")) : 32; // development aid; see trace_stream.cuh
            const int gs = std::min(c->want_stats ? c->stream_ctas_stats : c->stream_ctas, div_up(n_tiles, 4));
            switch ((c->want_stats ? 1 : 0) | (c->trace_lean ? 2 : 0)) { // MODE (trace.cuh): bit 0 event counters, bit 1 lean scene
                case 0: trace_stream_kernel<0><<<gs, 128, 0, ts>>>(c->ds, fc, tp, img, parity, cnt_dev, c->totals.p, refill_min); break;
                case 1: trace_stream_kernel<1><<<gs, 128, 0, ts>>>(c->ds, fc, tp, img, parity, cnt_dev, c->totals.p, refill_min); break;
                case 2: trace_stream_kernel<2><<<gs, 128, 0, ts>>>(c->ds, fc, tp, img, parity, cnt_dev, c->totals.p, refill_min); break;
                default: trace_stream_kernel<3><<<gs, 128, 0, ts>>>(c->ds, fc, tp, img, parity, cnt_dev, c->totals.p, refill_min); break;
            }
        } else {
            dim3 grid(div_up(W, 16), div_up(b - a, 8));
            switch ((c->want_stats ? 1 : 0) | (c->trace_lean ? 2 : 0)) {
                case 0: trace_kernel<0><<<grid, 128, 0, ts>>>(c->ds, fc, tp, img, parity, cnt_dev, c->totals.p); break;
                case 1: trace_kernel<1><<<grid, 128, 0, ts>>>(c->ds, fc, tp, img, parity, cnt_dev, c->totals.p); break;
                case 2: trace_kernel<2><<<grid, 128, 0, ts>>>(c->ds, fc, tp, img, parity, cnt_dev, c->totals.p); break;
                default: trace_kernel<3><<<grid, 128, 0, ts>>>(c->ds, fc, tp, img, parity, cnt_dev, c->totals.p); break;
            }
        }
        launches++;
    }
    CK(c, cudaEventRecord(c->ev[1], ts));
    if (ahead) { CK(c, cudaEventRecord(c->ev_trace_done[fp], ts)); CK(c, cudaStreamWaitEvent(s, c->ev_trace_done[fp], 0)); }
    { // K2
        int a, b; range(halo_after[0], a, b);
        TaaArgs t;
        t.cur = cur_plane; t.gnd_now = img.gnd[parity]; t.gnd_prev = img.gnd[parity ^ 1]; t.gas_now = img.gas[parity]; t.gas_prev = img.gas[parity ^ 1];
        t.hist = c->hist.p; t.W = W; t.H = H; t.y0 = a; t.y1 = b; t.reset = reset ? 1 : 0;
        float al = c->P.taa_alpha; al = al < 0.0f ? 0.0f : (al > 1.0f ? 1.0f : al); // :305
        t.alpha = al; t.pad = c->P.luminance_pad;
        taa_kernel<<<dim3(div_up(W, 32), div_up(b - a, 8)), dim3(32, 8), 0, s>>>(t);
        launches++;
        c->taa_valid = true;
    }
    CK(c, cudaEventRecord(c->ev[2], s));
    if (ahead) CK(c, cudaEventRecord(c->ev_taa_done[fp], s));
    if (front_only) { // the frame's à-trous passes, exposure and cells run on another rank (ycge_back_*); taa.CommitCamera :266
        c->launches_last = launches;
        c->last_gset = gset;
        memcpy(c->last_cam, c->snap_cam, sizeof c->last_cam); c->last_yaw = c->snap_yaw; c->last_pitch = c->snap_pitch;
        CK(c, cudaGetLastError());
        return 0;
    }
    { // K3: the reference's ping-pong including its in-place iteration (:648-719), resumable (see denoise_run)
        ycge_ctx::Denoise &d = c->dn;
        d.phys[0] = c->hist.p; d.phys[1] = sv.sa; d.phys[2] = sv.sb;
        d.cur_id = 0; d.dst_id = 1; d.it = 0; d.K = K; d.parity = gset; c->last_gset = gset; d.ty0 = ty0; d.ty1 = ty1; d.pending = false;
        for (int k = 0; k <= K; k++) d.halo_after[k] = halo_after[k];
        c->launches_last = launches;
        c->frame_open = true;
        int rc = denoise_run(c);
        if (rc) { c->frame_open = false; return rc; }
        return 0;
    }
}

// Runs à-trous passes from the saved state.  Unsharded: all of them, then the exposure samples.  Sharded: stops in
// front of an in-place pass after preparing it (pre-pass + sentinel fill, neither needs the neighbour's rows).
int denoise_run(ycge_ctx *c) {
    ycge_ctx::Denoise &d = c->dn;
    cudaStream_t s = c->stream;
    const int W = c->W, H = c->H, ss = c->ss;
    const EdgeDiv ed = c->edge_div;
    const bool fast = c->fast_div;
    const float4 *gnd = c->io ? c->io->gnd : gnd_of(c, d.parity), *gas = c->io ? c->io->gas : gas_of(c, d.parity); // d.parity = guide set of the frame
    SlotView sv = slot_view(c, c->io ? 0 : c->cur_slot);
    if (c->io) { sv.pre = c->io->pre; sv.logs = c->io->logs; sv.n_logs = c->io->n_logs; s = c->io->stream; }
    const int ticket_slot = c->io ? c->io->ticket_slot : c->cur_slot;
    DevBuf<float4> &pre = *sv.pre;
    int launches = 0;
    auto range = [&](int halo, int &a, int &b) { a = std::max(0, d.ty0 - halo); b = std::min(H, d.ty1 + halo); };
    auto to_back = [&]() -> int { // FRONT ends here: everything later only reads this frame's own slot and guide set
        if (!c->pipelined || s == c->back) return 0;
        CK(c, cudaEventRecord(sv.front_done, s));
        CK(c, cudaStreamWaitEvent(c->back, sv.front_done, 0));
        s = c->back;
        return 0;
    };
    while (d.it < d.K) {
        const int it = d.it;
        if (it >= 1) { int rc = to_back(); if (rc) return rc; }
        int a, b; range(d.halo_after[it + 1], a, b);
        const int step = 1 << it;
        if (d.cur_id == d.dst_id) {
            // in-place pass on scratch X: OLD = X, NEW = the other scratch (dead at this point), which then becomes X
            const int X = d.cur_id, Y = (X == 1) ? 2 : 1;
            const bool wave = step == 2 && c->use_wave; // the systolic form covers the stride of the reference's one in-place pass
            if (!d.pending) {
                // (1) everything that does not depend on new values, fully parallel
                if (wave) {
                    const WfGeom wg = wf_geom(W, H, a, b);
                    const size_t need = std::max((size_t)W * H * 25, wf_record_count(wg));
                    if (pre.n < need) { CK(c, cudaDeviceSynchronize()); CK(c, pre.alloc(need)); }
                    WavePreArgs pa;
                    pa.old_ = d.phys[X]; pa.gnd = gnd; pa.gas = gas; pa.rec = pre.p; pa.g = wg; pa.e = ed;
                    if (fast) atrous_wave_pre_kernel<true><<<dim3(div_up(W, 32), div_up(b - a, 8)), dim3(32, 8), 0, s>>>(pa);
                    else atrous_wave_pre_kernel<false><<<dim3(div_up(W, 32), div_up(b - a, 8)), dim3(32, 8), 0, s>>>(pa);
                } else {
                    if (pre.n < (size_t)W * H * 25) { CK(c, cudaDeviceSynchronize()); CK(c, pre.alloc((size_t)W * H * 25)); }
                    AtrousPreArgs pa;
                    pa.old_ = d.phys[X]; pa.gnd = gnd; pa.gas = gas; pa.pre = pre.p; pa.plane = (size_t)W * H;
                    pa.W = W; pa.H = H; pa.y0 = a; pa.y1 = b; pa.step = step; pa.e = ed;
                    if (fast) atrous_pre_kernel<true><<<dim3(div_up(W, 32), div_up(b - a, 8)), dim3(32, 8), 0, s>>>(pa);
                    else atrous_pre_kernel<false><<<dim3(div_up(W, 32), div_up(b - a, 8)), dim3(32, 8), 0, s>>>(pa);
                }
                launches++;
                // NEW starts as the sentinel on the rows this pass produces; the rows just above `a` (a sharded tile's
                // upper boundary, produced by the previous rank) are delivered by the caller before ycge_frame_inplace
                d.pa = a; d.pb = b;
                if (!(d.early_reset && it == 1)) {
                    int m0 = a;
                    if (c->peers && c->has_above) { int lo, aa, slo, sa_; halo_rows(c, lo, aa, slo, sa_); m0 = lo; } // the boundary rows arrive through the sentinel too
                    CK(c, cudaMemsetAsync(d.phys[Y] + (size_t)m0 * W, 0xFF, (size_t)(b - m0) * W * sizeof(float4), s));
                    if (c->peers && c->has_above) { peer_signal_kernel<<<1, 1, 0, s>>>(c->above_flags, (int)c->frame_counter); launches++; }
                }
                if (c->sharded) { d.pending = true; c->launches_last += launches; return 0; }
            }
            d.pending = false;
            // (2) the wavefront
            if (wave) {
                WaveArgs wa;
                wa.rec = pre.p; wa.new_ = d.phys[Y]; wa.g = wf_geom(W, H, d.pa, d.pb); wa.dc = ed.dc; wa.rc = ed.rc;
                wa.err = c->wave_err_dev; wa.trace = nullptr;
                wa.peer_new = nullptr; wa.peer_y0 = wa.peer_y1 = 0; wa.ready = c->flags.p; wa.frame = (int)c->frame_counter;
                if (c->peers && c->has_below) {
                    int lo, aa, slo, sa_; halo_rows(c, lo, aa, slo, sa_);
                    if (sa_ > slo) { wa.peer_new = (Y == 2) ? c->below_sb : c->below_sa; wa.peer_y0 = slo; wa.peer_y1 = sa_; }
                }
                if (getenv("YCGE_CHAIN_TRACE")) { // development aid: first / last step of every band, dumped by ycge_get_stats
                    if (c->chain_trace.n < (size_t)32 * wa.g.n_warps) CK(c, c->chain_trace.alloc((size_t)32 * wa.g.n_warps));
                    CK(c, cudaMemsetAsync(c->chain_trace.p, 0, c->chain_trace.n * 8, s));
                    wa.trace = c->chain_trace.p;
                }
                // one CTA per band (the band's warp and its halo warp); clusters of YCGE_WF_CLUSTER consecutive bands of one row
                // parity take one ticket and hand rows over through distributed shared memory
                const int CLS = c->wave_cluster;
                const int n_tickets = CLS > 1 ? 2 * div_up(wa.g.n_warps / 2, CLS) : wa.g.n_warps;
                wa.ticket = c->tickets.p + 16 * (2 * ticket_slot + 2); wa.ticket_base = c->ticket_base_of(0, ticket_slot);
                c->ticket_advance(0, ticket_slot, (unsigned int)n_tickets);
                CK(c, cudaEventRecord(c->ev[7], s));
                if (wa.g.n_warps > 0) {
                    cudaLaunchConfig_t lc = {};
                    lc.gridDim = dim3(CLS > 1 ? n_tickets * CLS : wa.g.n_warps); lc.blockDim = dim3(64); lc.stream = s;
                    // a cluster is placed inside one GPC and the block scheduler packs it onto few SMs: without a limit three bands
                    // can share an SM (and two of them a sub-partition) while other SMs idle.  Dynamic shared memory that nobody
                    // uses caps the residency at two CTAs per SM.
                    lc.dynamicSmemBytes = CLS > 1 ? (size_t)c->wave_pad_smem : 0;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = CLS > 1 ? CLS : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    lc.attrs = at; lc.numAttrs = 1;
                    const bool peer = wa.peer_new != nullptr;
                    if (CLS > 1) {
                        if (fast && peer) CK(c, cudaLaunchKernelEx(&lc, atrous_wave_kernel<true, true, YCGE_WF_CLUSTER>, wa));
                        else if (fast) CK(c, cudaLaunchKernelEx(&lc, atrous_wave_kernel<true, false, YCGE_WF_CLUSTER>, wa));
                        else if (peer) CK(c, cudaLaunchKernelEx(&lc, atrous_wave_kernel<false, true, YCGE_WF_CLUSTER>, wa));
                        else CK(c, cudaLaunchKernelEx(&lc, atrous_wave_kernel<false, false, YCGE_WF_CLUSTER>, wa));
                    } else {
                        if (fast && peer) CK(c, cudaLaunchKernelEx(&lc, atrous_wave_kernel<true, true, 1>, wa));
                        else if (fast) CK(c, cudaLaunchKernelEx(&lc, atrous_wave_kernel<true, false, 1>, wa));
                        else if (peer) CK(c, cudaLaunchKernelEx(&lc, atrous_wave_kernel<false, true, 1>, wa));
                        else CK(c, cudaLaunchKernelEx(&lc, atrous_wave_kernel<false, false, 1>, wa));
                    }
                    launches++;
                }
                CK(c, cudaEventRecord(c->ev[8], s));
                c->chain_timed = true;
                std::swap(d.phys[X], d.phys[Y]);
                const int tmp = d.cur_id;
                d.cur_id = d.dst_id;
                d.dst_id = (tmp == 1) ? 2 : 1;
                d.it++;
                continue;
            }
            AtrousChainArgs ia;
            ia.old_ = d.phys[X]; ia.new_ = d.phys[Y]; ia.pre = pre.p; ia.plane = (size_t)W * H;
            ia.W = W; ia.H = H; ia.step = step; ia.shift = it; ia.dc = ed.dc; ia.rc = ed.rc; ia.trace = nullptr; ia.err = c->wave_err_dev;
            ia.peer_new = nullptr; ia.peer_y0 = ia.peer_y1 = 0; ia.ready = c->flags.p; ia.frame = (int)c->frame_counter;
            if (c->peers && c->has_below) {
                int lo, aa, slo, sa_; halo_rows(c, lo, aa, slo, sa_);
                if (sa_ > slo) { ia.peer_new = (Y == 2) ? c->below_sb : c->below_sa; ia.peer_y0 = slo; ia.peer_y1 = sa_; }
                // NOTE: the ping-pong is the same on every rank, so the rank below also has its NEW in physical buffer Y
            }
            if (getenv("YCGE_CHAIN_TRACE")) { // development aid: per-chain timestamps, dumped by ycge_get_stats
                if (c->chain_trace.n < (size_t)H * step * 32) CK(c, c->chain_trace.alloc((size_t)H * step * 32));
                CK(c, cudaMemsetAsync(c->chain_trace.p, 0, c->chain_trace.n * 8, s));
                ia.trace = c->chain_trace.p;
            }
            if (c->inplace_ctas_per_launch <= 0) {
                cudaDeviceProp prop;
                CK(c, cudaGetDeviceProperties(&prop, c->device));
                int per_sm = YCGE_AIC_CTAS_PER_SM; // persistent launch: the same number of CTAs on every SM
                if (const char *e = getenv("YCGE_CHAIN_CTAS_PER_SM")) per_sm = std::max(1, atoi(e)); // development aid
                c->inplace_ctas_per_launch = std::max(1, per_sm * prop.multiProcessorCount);
                int occ = 0; // static launch: every CTA of a launch co-resident
                if (fast) CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, atrous_chain_static_kernel<true, true>, YCGE_AIC_WARPS * 32, 0));
                else CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, atrous_chain_static_kernel<false, true>, YCGE_AIC_WARPS * 32, 0));
                c->inplace_ctas_static = std::max(1, occ * prop.multiProcessorCount);
            }
            // A frame that has the GPU to itself uses the static form (rows in blockIdx order, launches sized so that all CTAs
            // of one are co-resident); frames that overlap (pipelined) use the persistent ticket form, which waits for nothing
            // that is not already running.
            const bool use_static = !c->pipelined && !c->io && !getenv("YCGE_CHAIN_PERSISTENT");
            const int rows_per_launch = std::max(1, c->inplace_ctas_static * YCGE_AIC_WARPS / step);
            CK(c, cudaEventRecord(c->ev[7], s));
            auto launch_chain = [&](cudaStream_t st, int r0, int r1, bool peer) {
                AtrousChainArgs q = ia;
                q.y0 = r0; q.y1 = r1;
                const int warps = (r1 - r0) * step; // one warp per chain, `step` chains per row
                const dim3 t(YCGE_AIC_WARPS * 32);
                if (use_static) {
                    const dim3 g(div_up(warps, YCGE_AIC_WARPS));
                    q.ticket = nullptr; q.ticket_base = 0; q.n_chains = (unsigned int)warps;
                    if (fast && peer) atrous_chain_static_kernel<true, true><<<g, t, 0, st>>>(q);
                    else if (fast) atrous_chain_static_kernel<true, false><<<g, t, 0, st>>>(q);
                    else if (peer) atrous_chain_static_kernel<false, true><<<g, t, 0, st>>>(q);
                    else atrous_chain_static_kernel<false, false><<<g, t, 0, st>>>(q);
                    launches++;
                    return;
                }
                const dim3 g(std::min(div_up(warps, YCGE_AIC_WARPS), c->inplace_ctas_per_launch));
                const int tk = (st == c->aux) ? 1 : 0; // launches that may overlap need separate ticket counters
                q.ticket = c->tickets.p + 16 * (tk ? 1 : 2 * ticket_slot + 2); q.ticket_base = c->ticket_base_of(tk, ticket_slot);
                q.n_chains = (unsigned int)warps;
                c->ticket_advance(tk, ticket_slot, (unsigned int)warps + g.x * YCGE_AIC_WARPS); // every warp's last ticket is a miss
                if (fast && peer) atrous_chain_kernel<true, true><<<g, t, 0, st>>>(q);
                else if (fast) atrous_chain_kernel<true, false><<<g, t, 0, st>>>(q);
                else if (peer) atrous_chain_kernel<false, true><<<g, t, 0, st>>>(q);
                else atrous_chain_kernel<false, false><<<g, t, 0, st>>>(q);
                launches++;
            };
            if (ia.peer_new && (!use_static || rows_per_launch >= b - a)) {
                // the rows that are also stored into the rank below run the (slower) peer variant as a second, concurrent
                // kernel on a side stream; everything above them runs the plain variant
                const int split = std::max(a, ia.peer_y0 - ((ia.peer_y0 - a) % std::max(1, YCGE_AIC_WARPS / step)));
            CK(c, cudaEventRecord(c->e_fork, s));
                if (split > a) launch_chain(s, a, split, false);
                CK(c, cudaStreamWaitEvent(c->aux, c->e_fork, 0));
                launch_chain(c->aux, split, b, true);
                CK(c, cudaEventRecord(c->e_join, c->aux));
                CK(c, cudaStreamWaitEvent(s, c->e_join, 0));
            } else {
                const int per = use_static ? rows_per_launch : b - a;
                for (int r0 = a; r0 < b; r0 += per) launch_chain(s, r0, std::min(b, r0 + per), ia.peer_new != nullptr);
            }
            CK(c, cudaEventRecord(c->ev[8], s));
            c->chain_timed = true;
            std::swap(d.phys[X], d.phys[Y]);
        } else {
            AtrousArgs aa;
            aa.src = d.phys[d.cur_id]; aa.gnd = gnd; aa.gas = gas; aa.dst = d.phys[d.dst_id];
            aa.W = W; aa.H = H; aa.y0 = a; aa.y1 = b; aa.step = step; aa.e = ed;
            if (fast) atrous_kernel<true><<<dim3(div_up(W, 32), div_up(b - a, 8)), dim3(32, 8), 0, s>>>(aa);
            else atrous_kernel<false><<<dim3(div_up(W, 32), div_up(b - a, 8)), dim3(32, 8), 0, s>>>(aa);
            launches++;
        }
        const int tmp = d.cur_id; // var tmp = cur; cur = dst; dst = (tmp == scratchA) ? scratchB : scratchA;   :718
        d.cur_id = d.dst_id;
        d.dst_id = (tmp == 1) ? 2 : 1;
        d.it++;
    }
    { int rc = to_back(); if (rc) return rc; }
    c->denoised = d.phys[d.cur_id];
    CK(c, cudaEventRecord(c->ev[3], s));
    { // K4a
        int step = std::max(2, ss * 2); // :226; sample row k is pixel row k*step = top row of cell row k
        int srow0 = div_up(d.ty0, step), srow1 = std::min(c->sh, div_up(d.ty1, step));
        if (c->sharded) CK(c, cudaMemsetAsync(sv.logs, 0, sv.n_logs * sizeof(float), s));
        if (srow1 > srow0) {
            exposure_log_kernel<<<dim3(div_up(c->sw, 128), srow1 - srow0), 128, 0, s>>>(c->denoised, gas, sv.logs, W, c->sw, step, srow0, srow1);
            launches++;
        }
    }
    CK(c, cudaEventRecord(c->ev[4], s));
    CK(c, cudaGetLastError());
    c->launches_last += launches;
    d.it = d.K + 1; // done
    return 0;
}

// the boundary rows of a prepared in-place pass: [lo, a) are needed from the rank above, [slo, sa) are owed to the rank below
void halo_rows(const ycge_ctx *c, int &lo, int &a, int &slo, int &sa) {
    const ycge_ctx::Denoise &d = c->dn;
    a = d.pa; lo = std::max(0, a - (2 << d.it)); // the pass reaches 2 taps of stride 2^it upwards
    sa = 0; slo = 0;
    if (d.ty1 < c->H) { // the next tile starts at d.ty1 with the same halo
        sa = std::max(0, d.ty1 - d.halo_after[d.it + 1]);
        slo = std::max(0, sa - (2 << d.it));
    }
}

int finish_launch(ycge_ctx *c, const float *logs, const float4 *den, cudaStream_t s, bool timed, ycge_cell *cells = nullptr) {
    ExposureParams ep;
    ep.tone_exposure = c->P.tone_exposure; ep.ae_key = c->P.ae_key; ep.ae_speed = c->P.ae_speed; ep.ae_min = c->P.ae_min; ep.ae_max = c->P.ae_max;
    ep.auto_exposure = c->P.auto_exposure;
    exposure_finish_kernel<<<1, 1024, 0, s>>>(logs, c->sw * c->sh, ep, c->expo.p);
    if (timed) CK(c, cudaEventRecord(c->ev[5], s));
    CellArgs ca;
    ca.den = den; ca.expo = c->expo.p; ca.cells = cells ? cells : c->cells.p; ca.W = c->W; ca.fbW = c->fbW; ca.ss = c->ss;
    ca.cy0 = c->tile_row0; ca.cy1 = c->tile_row0 + c->tile_rows;
    ca.gamma = c->P.tone_gamma; ca.saturation = c->P.saturation; ca.vibrance = c->P.vibrance;
    for (int k = 0; k < 5; k++) ca.th[k] = c->ansi_th[k];
    cells_kernel<<<dim3(div_up(c->fbW, 128), c->tile_rows), 128, 0, s>>>(ca);
    if (timed) CK(c, cudaEventRecord(c->ev[6], s));
    CK(c, cudaGetLastError());
    return 0;
}

int frame_finish_impl(ycge_ctx *c) {
    if (!c->frame_open) return fail(c, YCGE_ERR_INVALID, "ycge_frame_finish without ycge_frame_begin");
    if (c->dn.it <= c->dn.K) return fail(c, YCGE_ERR_INVALID, "ycge_frame_finish while an in-place pass is pending (ycge_frame_halo / ycge_frame_inplace)");
    const SlotView sv = slot_view(c, c->cur_slot);
    cudaStream_t fs = c->pipelined ? c->back : c->stream;
    if (c->pipelined && c->last_fin) CK(c, cudaStreamWaitEvent(fs, c->last_fin, 0)); // exposure state and cells are handed on frame to frame
    int rc = finish_launch(c, sv.logs, c->denoised, fs, true);
    if (rc) return rc;
    if (c->pipelined) {
        if (c->host_out) { // streaming path: this frame's cells to the caller's (pinned) buffer, still inside the FINISH chain
            CK(c, cudaMemcpy2DAsync(c->host_out, (size_t)c->host_stride * sizeof(ycge_cell), c->cells.p, (size_t)c->fbW * sizeof(ycge_cell),
                                    (size_t)c->fbW * sizeof(ycge_cell), (size_t)c->tile_rows, cudaMemcpyDeviceToHost, fs));
            CK(c, cudaEventRecord(sv.host_done, fs));
        }
        CK(c, cudaEventRecord(sv.fin_done, fs));
        c->last_fin = sv.fin_done;
    }
    c->launches_last += 2;
    // taa.CommitCamera (:266, TemporalAA.cs:69-76)
    memcpy(c->last_cam, c->snap_cam, sizeof c->last_cam); c->last_yaw = c->snap_yaw; c->last_pitch = c->snap_pitch;
    c->frame_open = false;
    return 0;
}

// orders the ctx's stream after the FINISH of the last pipelined frame, so that "synchronise the ctx's stream" keeps meaning
// "everything submitted is done"
int join_pipeline(ycge_ctx *c) {
    if (c->last_fin) { CK(c, cudaStreamWaitEvent(c->stream, c->last_fin, 0)); c->last_fin = nullptr; }
    return 0;
}

int read_cells_impl(ycge_ctx *c, ycge_cell *out, int stride) {
    if (!out) return fail(c, YCGE_ERR_INVALID, "out is NULL");
    { int rc = join_pipeline(c); if (rc) return rc; }
    if (stride <= 0) stride = c->fbW;
    if (stride < c->fbW) return fail(c, YCGE_ERR_INVALID, "stride smaller than fb_w");
    CK(c, cudaMemcpy2DAsync(out, (size_t)stride * sizeof(ycge_cell), c->cells.p, (size_t)c->fbW * sizeof(ycge_cell), (size_t)c->fbW * sizeof(ycge_cell),
                            (size_t)c->tile_rows, cudaMemcpyDeviceToHost, c->stream));
    CK(c, sync_ctx_streams(c));
    return wave_check(c);
}

} // namespace

#include "ycge_multi.inl"
ycge_ctx::~ycge_ctx() {}

// =============================================================================================== exported C ABI
extern "C" {

YCGE_API void ycge_default_params(ycge_params *p) {
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->diffuse_bounces = 1; p->max_mirror_bounces = 2; p->max_refractions = 2; p->atrous_iterations = 3;
    p->mirror_threshold = 0.9f; p->eps = 1e-4f; p->taa_alpha = 0.01f; p->motion_trans_reset = 0.0025f; p->motion_rot_reset = 0.0025f;
    p->diffuse_sigma_deg = 25.0f; p->luminance_pad = 0.10f; p->c_phi = 3.0f; p->n_phi = 0.35f; p->z_phi = 2.0f; p->a_phi = 0.20f;
    p->tone_exposure = 1.0f; p->tone_gamma = 2.2f; p->ae_key = 0.18f; p->ae_speed = 0.2f; p->ae_min = 0.10f; p->ae_max = 1.50f;
    p->saturation = 2.0f; p->vibrance = 0.0f; p->auto_exposure = 1; p->seed_salt = 0x9E3779B97F4A7C15ULL;
}

YCGE_API const char *ycge_last_error(ycge_ctx *ctx) { return ctx ? ctx->err.c_str() : tl_error.c_str(); }

YCGE_API int ycge_create(const ycge_config *cfg, ycge_ctx **out) try {
    if (!cfg || !out) return fail(nullptr, YCGE_ERR_INVALID, "cfg/out is NULL");
    *out = nullptr;
    if (cfg->n_devices < 0) return fail(nullptr, YCGE_ERR_INVALID, "negative n_devices");
    if (cfg->n_devices >= 2) return group_create(cfg, out);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, YCGE_ERR_CUDA, std::string("no usable CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, YCGE_ERR_INVALID, "device ordinal out of range");
    std::unique_ptr<ycge_ctx> c(new ycge_ctx());
    c->device = cfg->device;
    c->P = cfg->params;
    if (const char *e = getenv("YCGE_TRACE_VARIANT")) c->trace_variant = atoi(e) ? 1 : 0; // development aid
    if (c->P.atrous_iterations > 8) return fail(nullptr, YCGE_ERR_INVALID, "atrous_iterations > 8 not supported");
    // the deferred reflection / refraction branches of one pixel (RaytraceRenderer.cs:463-468) live in a stack of YCGE_PATH_STACK
    // items: a transparent hit pushes two and continues with one, so mirror depth d needs d + 1 slots
    if (c->P.max_mirror_bounces < 0 || c->P.max_mirror_bounces + 1 > YCGE_PATH_STACK) return fail(nullptr, YCGE_ERR_LIMIT, "max_mirror_bounces must be in [0, 15] (depth of the device's per-pixel branch stack)");
    if (c->P.max_refractions < 0 || c->P.diffuse_bounces < 0) return fail(nullptr, YCGE_ERR_INVALID, "negative bounce count");
    CK(nullptr, cudaSetDevice(c->device));
    CK(nullptr, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    for (auto &ev : c->ev) CK(nullptr, cudaEventCreate(&ev));
    CK(nullptr, cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
    CK(nullptr, cudaEventCreateWithFlags(&c->e_fork, cudaEventDisableTiming));
    CK(nullptr, cudaEventCreateWithFlags(&c->e_join, cudaEventDisableTiming));
    CK(nullptr, c->expo.alloc(1));
    ExposureState es;
    es.ae_exposure = 1.0f; es.effective = 1.0f; es.log_sum = 0.0f; es.cnt = 0; // ToneMapper.cs:13,17
    CK(nullptr, cudaMemcpyAsync(c->expo.p, &es, sizeof es, cudaMemcpyHostToDevice, c->stream));
    CK(nullptr, c->counters.alloc(1));
    CK(nullptr, cudaMemsetAsync(c->counters.p, 0, sizeof(TraceCounters), c->stream));
    CK(nullptr, c->tickets.alloc(132 * 16));
    CK(nullptr, cudaMemsetAsync(c->tickets.p, 0, c->tickets.n * sizeof(unsigned int), c->stream));
    CK(nullptr, cudaStreamCreateWithPriority(&c->st0, cudaStreamNonBlocking, back_priority()));
    CK(nullptr, cudaEventCreateWithFlags(&c->front_done0, cudaEventDisableTiming));
    CK(nullptr, cudaEventCreateWithFlags(&c->fin_done0, cudaEventDisableTiming));
    CK(nullptr, cudaEventCreateWithFlags(&c->host_done0, cudaEventDisableTiming));
    CK(nullptr, c->flags.alloc(16));
    CK(nullptr, cudaMemsetAsync(c->flags.p, 0, 16 * sizeof(int), c->stream));
    { // the wavefront kernels' "gave up waiting" flag lives in mapped host memory: the host reads it after any wait, for free
        int *h = nullptr;
        CK(nullptr, cudaHostAlloc((void **)&h, sizeof(int), cudaHostAllocMapped));
        *h = 0;
        c->wave_err_host = h;
        CK(nullptr, cudaHostGetDevicePointer((void **)&c->wave_err_dev, h, 0));
    }
    if (const char *e = getenv("YCGE_FRONT_AHEAD")) { c->front_ahead = atoi(e) != 0; c->front_ahead_force = atoi(e) > 1; } // 0: the trace of a FRONT-only frame stays on the context's stream; 2: own streams whatever the tile size
    if (const char *e = getenv("YCGE_WAVE")) c->use_wave = atoi(e) != 0; // 1 = the systolic wavefront kernels (wavefront.cuh)
    if (const char *e = getenv("YCGE_WAVE_CLUSTER")) c->wave_cluster = atoi(e) > 1 ? YCGE_WF_CLUSTER : 1; // opt-in: thread-block clusters of 8 bands
    if (const char *e = getenv("YCGE_WAVE_PAD_SMEM")) c->wave_pad_smem = atoi(e);                          // development aid
    if (c->wave_cluster > 1) {
        const int pad = c->wave_pad_smem;
        CK(nullptr, cudaFuncSetAttribute(atrous_wave_kernel<true, true, YCGE_WF_CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad));
        CK(nullptr, cudaFuncSetAttribute(atrous_wave_kernel<true, false, YCGE_WF_CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad));
        CK(nullptr, cudaFuncSetAttribute(atrous_wave_kernel<false, true, YCGE_WF_CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad));
        CK(nullptr, cudaFuncSetAttribute(atrous_wave_kernel<false, false, YCGE_WF_CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad));
    }
    CK(nullptr, c->totals.alloc(1));
    CK(nullptr, cudaMemsetAsync(c->totals.p, 0, sizeof(TraceTotals), c->stream));
    { // edge-stopping divisors; verify the fast division over every non-negative binary32 numerator (a few ms, once)
        EdgeDiv &e = c->edge_div;
        e.dc = std::max(1e-6f, c->P.c_phi); e.dn = std::max(1e-6f, c->P.n_phi); e.dz = std::max(1e-6f, c->P.z_phi); e.da = std::max(1e-6f, c->P.a_phi);
        e.rc = (float)(1.0 / (double)e.dc); e.rn = (float)(1.0 / (double)e.dn); e.rz = (float)(1.0 / (double)e.dz); e.ra = (float)(1.0 / (double)e.da);
        DevBuf<unsigned int> mm;
        CK(nullptr, mm.alloc(8));
        CK(nullptr, cudaMemsetAsync(mm.p, 0, 32, c->stream));
        div_selftest_kernel<<<(0x7F800000u >> 8) + 1, 256, 0, c->stream>>>(e, mm.p);
        unsigned int h[8] = {1, 1, 1, 1, 1, 1, 1, 1};
        CK(nullptr, cudaMemcpyAsync(h, mm.p, 32, cudaMemcpyDeviceToHost, c->stream));
        CK(nullptr, sync_ctx_streams(c.get()));
        c->fast_div = (h[0] | h[1] | h[2] | h[3] | h[4]) == 0;
        if (!(e.dc < 1e30f && e.dn < 1e30f && e.dz < 1e30f && e.da < 1e30f)) c->fast_div = false;
    }
    static const int bounds[5] = {48, 114, 154, 194, 234}; // ANSITerminalRenderer.cs:288-296
    for (int k = 0; k < 5; k++) c->ansi_th[k] = ansi_threshold(bounds[k]);
    int rc = set_geometry(c.get(), cfg->fb_w, cfg->fb_h, cfg->ss, cfg->tile_row0, cfg->tile_rows);
    if (rc != 0) { tl_error = c->err; return rc; }
    CK(nullptr, sync_ctx_streams(c.get()));
    memset(&c->ds, 0, sizeof c->ds);
    *out = c.release();
    return 0;
} YCGE_CATCH

YCGE_API void ycge_destroy(ycge_ctx *ctx) {
    if (!ctx) return;
    if (ctx->group) { delete ctx; return; } // the group's destructor releases its per-device contexts, streams and events
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaDeviceSynchronize();
    ctx->slots.clear();
    if (ctx->st0) cudaStreamDestroy(ctx->st0);
    if (ctx->front_done0) cudaEventDestroy(ctx->front_done0);
    if (ctx->fin_done0) cudaEventDestroy(ctx->fin_done0);
    if (ctx->host_done0) cudaEventDestroy(ctx->host_done0);
    for (auto &p : ctx->ipc_opened) if (p) cudaIpcCloseMemHandle(p);
    if (ctx->aux) { cudaStreamSynchronize(ctx->aux); cudaStreamDestroy(ctx->aux); }
    for (int k = 0; k < 2; k++) {
        if (ctx->trace_stream[k]) { cudaStreamSynchronize(ctx->trace_stream[k]); cudaStreamDestroy(ctx->trace_stream[k]); }
        if (ctx->ev_trace_done[k]) cudaEventDestroy(ctx->ev_trace_done[k]);
        if (ctx->ev_taa_done[k]) cudaEventDestroy(ctx->ev_taa_done[k]);
    }
    if (ctx->e_fork) cudaEventDestroy(ctx->e_fork);
    if (ctx->e_join) cudaEventDestroy(ctx->e_join);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->wave_err_host) cudaFreeHost((void *)ctx->wave_err_host);
    delete ctx;
}

YCGE_API int ycge_resize(ycge_ctx *c, int32_t fb_w, int32_t fb_h, int32_t ss) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_resize is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaDeviceSynchronize()); // pipelined frames may still run on the slots' streams
    c->last_fin = nullptr; c->in_flight.clear(); c->pre.release();
    int row0 = c->sharded ? c->tile_row0 : 0, rows = c->sharded ? c->tile_rows : 0;
    if (c->sharded && row0 + rows > fb_h) return fail(c, YCGE_ERR_INVALID, "tile exceeds resized framebuffer");
    int rc = set_geometry(c, fb_w, fb_h, ss, row0, rows);
    if (rc) return rc;
    if (!c->back_slots.empty()) { rc = alloc_back_slots(c, (int)c->back_slots.size()); if (rc) return rc; } // same number, new geometry
    if (c->debug_rays) CK(c, c->rays_dbg.alloc((size_t)c->W * c->H * 6));
    // TemporalAA.Resize (TemporalAA.cs:33-45): the camera memory is cleared; frame counter and exposure survive (:110-138)
    c->last_cam[0] = c->last_cam[1] = c->last_cam[2] = NAN; c->last_yaw = NAN; c->last_pitch = NAN;
    return 0;
} YCGE_CATCH

YCGE_API int ycge_set_stream(ycge_ctx *c, void *cuda_stream) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_set_stream is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    { int rc = join_pipeline(c); if (rc) return rc; }
    CK(c, sync_ctx_streams(c));
    if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
    c->stream = (cudaStream_t)cuda_stream;
    return 0;
} YCGE_CATCH

// ---- meshes ------------------------------------------------------------------------------------------------
static int store_mesh(ycge_ctx *c, int id, int n, const float *soa12 /* n x (A,e1,e2,n) */, const TreeView &tv, const ycge_material &mat) {
    std::vector<PairNode> pairs;
    TreeRoot root;
    int rc = flatten(c, tv, pairs, root);
    if (rc) return rc;
    std::vector<DevTri> tris(tv.n_leaf);
    std::vector<int> ids(tv.n_leaf);
    for (int s = 0; s < tv.n_leaf; s++) {
        int t = tv.leaf[s];
        if (t < 0 || t >= n) return fail(c, YCGE_ERR_INVALID, "leafTriIndex out of range");
        const float *q = soa12 + 12 * (size_t)t;
        tris[s].t0 = make_float4(q[0], q[1], q[2], q[3]);
        tris[s].t1 = make_float4(q[4], q[5], q[6], q[7]);
        tris[s].t2 = make_float4(q[8], q[9], q[10], q[11]);
        ids[s] = t;
    }
    std::unique_ptr<MeshStore> m(new MeshStore());
    CK(c, m->nodes.upload(pairs, c->stream));
    CK(c, m->tris.upload(tris, c->stream));
    CK(c, m->tri_id.upload(ids, c->stream));
    m->root = root; m->material = mat; m->n_tris = n; m->n_pairs = (int)pairs.size();
    c->meshes[id] = std::move(m);
    c->have_scene = false; // object table must be rebuilt
    return 0;
}

YCGE_API int ycge_mesh_upload_soa(ycge_ctx *c, int32_t id, const ycge_mesh_soa *mesh) try {
    if (c && c->group) return group_each_front(c, [&](ycge_ctx *f) { return ycge_mesh_upload_soa(f, id, mesh); });
    if (!c || !mesh || !mesh->bvh) return fail(c, YCGE_ERR_INVALID, "ctx/mesh/mesh->bvh is NULL");
    if (mesh->n_tris < 0) return fail(c, YCGE_ERR_INVALID, "negative triangle count");
    if (mesh->n_tris > 0 && (!mesh->ax || !mesh->ay || !mesh->az || !mesh->e1x || !mesh->e1y || !mesh->e1z || !mesh->e2x || !mesh->e2y || !mesh->e2z || !mesh->nx || !mesh->ny || !mesh->nz))
        return fail(c, YCGE_ERR_INVALID, "a triangle array of the mesh is NULL");
    CK(c, cudaSetDevice(c->device));
    int n = mesh->n_tris;
    std::vector<float> soa((size_t)n * 12);
    for (int i = 0; i < n; i++) {
        float *d = &soa[(size_t)i * 12];
        d[0] = mesh->ax[i]; d[1] = mesh->ay[i]; d[2] = mesh->az[i]; d[3] = mesh->e1x[i]; d[4] = mesh->e1y[i]; d[5] = mesh->e1z[i];
        d[6] = mesh->e2x[i]; d[7] = mesh->e2y[i]; d[8] = mesh->e2z[i]; d[9] = mesh->nx[i]; d[10] = mesh->ny[i]; d[11] = mesh->nz[i];
    }
    return store_mesh(c, id, n, soa.data(), view_of(*mesh->bvh), mesh->material);
} YCGE_CATCH

YCGE_API int ycge_mesh_upload_triangles(ycge_ctx *c, int32_t id, int32_t n, const float *abc, const ycge_material *material) try {
    if (c && c->group) return group_each_front(c, [&](ycge_ctx *f) { return ycge_mesh_upload_triangles(f, id, n, abc, material); });
    if (!c || !abc || !material || n < 0) return fail(c, YCGE_ERR_INVALID, "bad argument");
    CK(c, cudaSetDevice(c->device));
    std::vector<float> soa((size_t)n * 12);
    std::vector<BuildItem> items((size_t)n);
    for (int i = 0; i < n; i++) { // MeshBVH.cs:83-97
        const float *t = abc + 9 * (size_t)i;
        items[i] = triangle_item(i, t);
        float *d = &soa[(size_t)i * 12];
        d[0] = t[0]; d[1] = t[1]; d[2] = t[2];
        float lx = t[3] - t[0], ly = t[4] - t[1], lz = t[5] - t[2];
        float mx = t[6] - t[0], my = t[7] - t[1], mz = t[8] - t[2];
        d[3] = lx; d[4] = ly; d[5] = lz; d[6] = mx; d[7] = my; d[8] = mz;
        float nnx = ly * mz - lz * my, nny = lz * mx - lx * mz, nnz = lx * my - ly * mx;
        float invLen = 1.0f / detail::net_max(1e-20f, std::sqrt(nnx * nnx + nny * nny + nnz * nnz));
        d[9] = nnx * invLen; d[10] = nny * invLen; d[11] = nnz * invLen;
    }
    FlatTree tree;
    build_reference_tree(items, 8, true, tree);
    return store_mesh(c, id, n, soa.data(), view_of(tree), *material);
} YCGE_CATCH

// completes a device build in flight: waits for its root record, checks that every triangle was placed
static int resolve_mesh(ycge_ctx *c, MeshStore &m) {
    if (!m.pending) return 0;
    CK(c, cudaEventSynchronize(m.pending));
    const DbCounters cnt = m.pending_host->cnt;
    m.root = m.pending_host->root;
    cudaEventDestroy(m.pending); m.pending = nullptr;
    cudaFreeHost(m.pending_host); m.pending_host = nullptr;
    if (cnt.done_items != m.n_tris) return fail(c, YCGE_ERR_CUDA, "device BVH build did not place every triangle");
    m.n_pairs = cnt.n_nodes > 1 ? (int)(cnt.n_nodes - 1) / 2 : 0; // a full binary tree: inner = (nodes - 1) / 2
    m.sort_fallbacks = cnt.fallbacks;
    return 0;
}
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// the same tree, built on the device (bvh_device.cuh).  Everything is enqueued on the context's stream; the
// host only waits for the 32-byte root record (the object table of ycge_scene_upload carries the root box by value).
YCGE_API int ycge_mesh_build_device(ycge_ctx *c, int32_t id, int32_t n, const float *abc, const ycge_material *material) try {
    if (c && c->group) return group_each_front(c, [&](ycge_ctx *f) { return ycge_mesh_build_device(f, id, n, abc, material); });
    if (!c || !abc || !material || n < 0) return fail(c, YCGE_ERR_INVALID, "bad argument");
    if (n == 0) return ycge_mesh_upload_triangles(c, id, n, abc, material);
    if (n >= YCGE_LEAF_MAX_START) return fail(c, YCGE_ERR_LIMIT, "mesh larger than 2^26 triangles");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    int n_sm = 0;
    CK(c, cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device));
    const int grid = 2 * n_sm;
    const size_t N = (size_t)n;
    // scratch: one grow-only arena per context (cudaMalloc / cudaFree of a dozen buffers cost more than the build itself)
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t o_abc = carve(N * 9 * 4), o_box = carve(N * 9 * 4), o_soa = carve(N * 12 * 4), o_nbox = carve(N * 2 * 6 * 4);
    const size_t o_int = carve((N * 6 + N + grid + 16) * 4), o_nint = carve(N * 2 * 3 * 4), o_nodes = carve(N * 2 * sizeof(DbNode));
    const size_t o_cnt = carve(sizeof(DbCounters)), o_root = carve(sizeof(TreeRoot));
    if (c->db_scratch.n < off) { CK(c, cudaStreamSynchronize(s)); CK(c, c->db_scratch.alloc(off)); }
    unsigned char *base = c->db_scratch.p;
    std::unique_ptr<MeshStore> m(new MeshStore());
    CK(c, m->nodes.alloc(N)); CK(c, m->tris.alloc(N)); CK(c, m->tri_id.alloc(N));
    float *d_abc = reinterpret_cast<float *>(base + o_abc), *d_box = reinterpret_cast<float *>(base + o_box);
    int *d_int = reinterpret_cast<int *>(base + o_int), *d_nint = reinterpret_cast<int *>(base + o_nint);
    TreeRoot *d_root = reinterpret_cast<TreeRoot *>(base + o_root);
    DbCounters *d_cnt = reinterpret_cast<DbCounters *>(base + o_cnt);
    CK(c, cudaMemcpyAsync(d_abc, abc, N * 9 * sizeof(float), cudaMemcpyHostToDevice, s));
    DbArgs a;
    a.n = n; a.abc = d_abc;
    a.blo = d_box; a.bhi = d_box + 3 * N; a.cen = d_box + 6 * N; a.soa12 = reinterpret_cast<float *>(base + o_soa);
    a.idx = d_int; a.tmp = a.idx + N; a.pre = a.tmp + N; a.rpos = a.pre + N; a.lpos = a.rpos + N; a.queue = a.lpos + N; a.ready = a.queue + N;
    a.nodes = reinterpret_cast<DbNode *>(base + o_nodes); a.cnt = d_cnt;
    a.nbox = reinterpret_cast<float *>(base + o_nbox); a.nsize = d_nint; a.ninner = a.nsize + 2 * N; a.nflag = a.ninner + 2 * N;
    a.pairs = m->nodes.p; a.tris = m->tris.p; a.tri_id = m->tri_id.p; a.root = d_root;
    CK(c, cudaMemsetAsync(a.ready, 0, (N + grid + 16) * sizeof(int), s));
    CK(c, cudaMemsetAsync(a.nflag, 0, 2 * N * sizeof(int), s));
    const bool timing = getenv("YCGE_DB_TIME") != nullptr; // development aid
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    const double h0 = now_ms();
    if (timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventRecord(e0, s); }
    db_items_kernel<<<div_up(n, 256), 256, 0, s>>>(a);
    db_init_kernel<<<1, 1, 0, s>>>(a);
    if (timing) cudaEventRecord(e1, s);
    db_build_kernel<<<grid, YCGE_DB_THREADS, 0, s>>>(a);
    if (timing) cudaEventRecord(e2, s);
    db_boxes_kernel<<<div_up(2 * n, 256), 256, 0, s>>>(a);
    db_emit_kernel<<<div_up(2 * n, 256), 256, 0, s>>>(a);
    db_tris_kernel<<<div_up(n, 256), 256, 0, s>>>(a);
    CK(c, cudaGetLastError());
    // root record and counters: to pinned memory behind an event; whoever needs them first waits (ycge_scene_upload,
    // ycge_mesh_debug_read): the call returns with the build still running
    CK(c, cudaHostAlloc((void **)&m->pending_host, sizeof(MeshStore::Pending), cudaHostAllocDefault));
    CK(c, cudaEventCreateWithFlags(&m->pending, cudaEventDisableTiming));
    CK(c, cudaMemcpyAsync(&m->pending_host->root, d_root, sizeof(TreeRoot), cudaMemcpyDeviceToHost, s));
    CK(c, cudaMemcpyAsync(&m->pending_host->cnt, d_cnt, sizeof(DbCounters), cudaMemcpyDeviceToHost, s));
    CK(c, cudaEventRecord(m->pending, s));
    m->material = *material; m->n_tris = n;
    if (timing) {
        CK(c, cudaStreamSynchronize(s));
        const DbCounters &cnt = m->pending_host->cnt;
        float t01 = 0, t12 = 0;
        cudaEventElapsedTime(&t01, e0, e1); cudaEventElapsedTime(&t12, e1, e2);
        fprintf(stderr, "ycge_mesh_build_device: %d triangles: items %.3f ms, build %.3f ms, enqueue->done %.3f ms host, nodes %u, fallbacks %u\n", n, t01, t12, now_ms() - h0, cnt.n_nodes, cnt.fallbacks);
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    }
    c->meshes[id] = std::move(m);
    c->have_scene = false; // object table must be rebuilt
    return 0;
} YCGE_CATCH

// Development / test aid: the stored device arrays of a mesh.  what: 0 = pair nodes (64 B each), 1 = triangles in leaf order
// (48 B), 2 = leaf slot -> triangle index (4 B), 3 = the root record (32 B).  dst == NULL: *bytes = size of that array.
YCGE_API int ycge_mesh_debug_read(ycge_ctx *c, int32_t id, int32_t what, void *dst, size_t *bytes) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_mesh_debug_read is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || !bytes) return fail(c, YCGE_ERR_INVALID, "bad argument");
    auto it = c->meshes.find(id);
    if (it == c->meshes.end()) return fail(c, YCGE_ERR_INVALID, "no such mesh");
    MeshStore &m = *it->second;
    { int rc = resolve_mesh(c, m); if (rc) return rc; }
    const void *src = nullptr;
    size_t need = 0;
    switch (what) {
        case 0: src = m.nodes.p; need = (size_t)m.n_pairs * sizeof(PairNode); break;
        case 1: src = m.tris.p; need = (size_t)m.n_tris * sizeof(DevTri); break;
        case 2: src = m.tri_id.p; need = (size_t)m.n_tris * sizeof(int); break;
        case 3: need = sizeof(TreeRoot); break;
        default: return fail(c, YCGE_ERR_INVALID, "unknown array");
    }
    if (!dst) { *bytes = need; return 0; }
    if (*bytes < need) return fail(c, YCGE_ERR_INVALID, "destination too small");
    CK(c, cudaSetDevice(c->device));
    CK(c, sync_ctx_streams(c));
    if (what == 3) memcpy(dst, &m.root, need);
    else if (need) CK(c, cudaMemcpy(dst, src, need, cudaMemcpyDeviceToHost));
    *bytes = need;
    return 0;
} YCGE_CATCH

// ---- volumes -----------------------------------------------------------------------------------------------
YCGE_API int ycge_volume_upload(ycge_ctx *c, int32_t id, const ycge_volume *v) try {
    if (c && c->group) return group_each_front(c, [&](ycge_ctx *f) { return ycge_volume_upload(f, id, v); });
    if (!c || !v || !v->mat || !v->meta || !v->palette) return fail(c, YCGE_ERR_INVALID, "bad argument");
    if (v->palette_n_ids < 0 || v->palette_meta_levels < 0) return fail(c, YCGE_ERR_INVALID, "negative voxel palette size");
    if ((size_t)v->palette_n_ids * (size_t)(v->palette_meta_levels < 1 ? 1 : v->palette_meta_levels) > (size_t)1 << 24) return fail(c, YCGE_ERR_LIMIT, "voxel palette larger than 2^24 entries");
    if (v->nx <= 0 || v->ny <= 0 || v->nz <= 0) return fail(c, YCGE_ERR_UNBOUNDED, "empty VolumeGrid has no bounds");
    CK(c, cudaSetDevice(c->device));
    std::unique_ptr<VolumeStore> vs(new VolumeStore());
    DevVolume &d = vs->dv;
    memset(&d, 0, sizeof d);
    d.nx = v->nx; d.ny = v->ny; d.nz = v->nz;
    d.nbx = (v->nx + 7) >> 3; d.nby = (v->ny + 7) >> 3; d.nbz = (v->nz + 7) >> 3;
    size_t cap = (size_t)d.nbx * d.nby * d.nbz * 512;
    if (cap > (size_t)0x7fffffff) return fail(c, YCGE_ERR_LIMIT, "VolumeGrid larger than 2^31 voxels (the reference indexes with int)");
    for (int k = 0; k < 3; k++) { d.min_corner[k] = v->min_corner[k]; d.voxel_size[k] = v->voxel_size[k]; }
    d.wireframe = v->wireframe; d.wire_width_frac = v->wire_width_frac; d.wire_max_distance = v->wire_max_distance;
    CK(c, vs->vox.alloc(cap));
    CK(c, vs->raw_mat.alloc(cap));
    CK(c, vs->raw_meta.alloc(cap));
    CK(c, cudaMemcpyAsync(vs->raw_mat.p, v->mat, cap * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(vs->raw_meta.p, v->meta, cap * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    vs->n_ids = v->palette_n_ids; vs->levels = v->palette_meta_levels < 1 ? 1 : v->palette_meta_levels; vs->def = v->palette_default;
    vs->palette.assign(v->palette, v->palette + (size_t)vs->n_ids * vs->levels);
    for (int mi : vs->palette) if (mi < 0 || mi >= 255) return fail(c, YCGE_ERR_LIMIT, "voxel palette material index must be in [0,254]");
    if (vs->def < 0 || vs->def >= 255) return fail(c, YCGE_ERR_LIMIT, "voxel palette default material index must be in [0,254]");
    // materialLookup(id, meta) resolved now: one byte per voxel
    DevBuf<int> pal;
    CK(c, pal.upload(vs->palette, c->stream));
    voxel_pack_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(vs->raw_mat.p, vs->raw_meta.p, vs->vox.p, cap, pal.p, vs->n_ids, vs->levels, vs->def);
    CK(c, cudaGetLastError());
    CK(c, vs->occ.alloc(cap / 512));
    voxel_occupancy_kernel<<<(unsigned)((cap / 64 + 255) / 256), 256, 0, c->stream>>>(vs->vox.p, vs->occ.p, cap / 64);
    CK(c, cudaGetLastError());
    CK(c, sync_ctx_streams(c));
    vs->raw_mat.release(); vs->raw_meta.release();
    vs->packed = true;
    d.vox = vs->vox.p;
    d.occ = vs->occ.p;
    c->volumes[id] = std::move(vs);
    c->have_scene = false;
    return 0;
} YCGE_CATCH

// ---- textures ----------------------------------------------------------------------------------------------
YCGE_API int ycge_texture_upload(ycge_ctx *c, int32_t id, int32_t w, int32_t h, const uint32_t *rgba) try {
    if (c && c->group) return group_each_front(c, [&](ycge_ctx *f) { return ycge_texture_upload(f, id, w, h, rgba); });
    if (!c || id < 0 || w < 0 || h < 0 || ((size_t)w * h > 0 && !rgba)) return fail(c, YCGE_ERR_INVALID, "bad argument");
    if ((size_t)w * h > (size_t)0x7fffffff) return fail(c, YCGE_ERR_LIMIT, "texture larger than 2^31 texels (the reference indexes with int)");
    CK(c, cudaSetDevice(c->device));
    CK(c, sync_ctx_streams(c));
    std::unique_ptr<TextureStore> t(new TextureStore());
    t->w = w; t->h = h;
    if ((size_t)w * h > 0) {
        CK(c, t->px.alloc((size_t)w * h));
        CK(c, cudaMemcpyAsync(t->px.p, rgba, (size_t)w * h * 4, cudaMemcpyHostToDevice, c->stream));
        CK(c, sync_ctx_streams(c));
    }
    c->textures[id] = std::move(t);
    c->have_scene = false; // the texture table is rebuilt by ycge_scene_upload
    return 0;
} YCGE_CATCH

// ---- scene -------------------------------------------------------------------------------------------------
static bool object_bounds(ycge_ctx *c, const ycge_object &o, Aabb &box, float cen[3]) {
    const float *p = o.p;
    const float E = 1e-4f;
    switch (o.kind) {
        case YCGE_SPHERE: case YCGE_DISK: { // BoundedObjects.cs:20-29, Surfaces.cs:96-105
            float R = o.kind == YCGE_SPHERE ? p[3] : p[6];
            for (int k = 0; k < 3; k++) { box.lo[k] = p[k] - R; box.hi[k] = p[k] + R; }
            break; }
        case YCGE_PLANE: // Surfaces.cs:30-36: +-1e6 box, centroid exactly 0
            for (int k = 0; k < 3; k++) { box.lo[k] = -1e6f; box.hi[k] = 1e6f; cen[k] = 0.0f; }
            return true;
        case YCGE_XYRECT: box.lo[0] = p[0]; box.lo[1] = p[2]; box.lo[2] = p[4] - E; box.hi[0] = p[1]; box.hi[1] = p[3]; box.hi[2] = p[4] + E; break;
        case YCGE_XZRECT: box.lo[0] = p[0]; box.lo[1] = p[4] - E; box.lo[2] = p[2]; box.hi[0] = p[1]; box.hi[1] = p[4] + E; box.hi[2] = p[3]; break;
        case YCGE_YZRECT: box.lo[0] = p[4] - E; box.lo[1] = p[0]; box.lo[2] = p[2]; box.hi[0] = p[4] + E; box.hi[1] = p[1]; box.hi[2] = p[3]; break;
        case YCGE_BOX: for (int k = 0; k < 3; k++) { box.lo[k] = p[k]; box.hi[k] = p[3 + k]; } break;
        case YCGE_CYLINDER_Y: box.lo[0] = p[0] - p[3]; box.lo[1] = p[4]; box.lo[2] = p[2] - p[3]; box.hi[0] = p[0] + p[3]; box.hi[1] = p[5]; box.hi[2] = p[2] + p[3]; break;
        case YCGE_TRIANGLE: { BuildItem it = triangle_item(0, p); box = it.box; break; } // Triangle.cs:54-66 (same rule as MeshBVH)
        case YCGE_MESH: {
            auto it = c->meshes.find(o.ref_id);
            if (it == c->meshes.end() || it->second->root.ref == YCGE_REF_NONE) return false;
            for (int k = 0; k < 3; k++) { box.lo[k] = it->second->root.lo[k]; box.hi[k] = it->second->root.hi[k]; }
            break; }
        case YCGE_VOLUME: { // VolumeGrid.cs:386-403
            auto it = c->volumes.find(o.ref_id);
            if (it == c->volumes.end()) return false;
            const DevVolume &d = it->second->dv;
            box.lo[0] = d.min_corner[0]; box.lo[1] = d.min_corner[1]; box.lo[2] = d.min_corner[2];
            box.hi[0] = d.min_corner[0] + d.nx * d.voxel_size[0]; box.hi[1] = d.min_corner[1] + d.ny * d.voxel_size[1]; box.hi[2] = d.min_corner[2] + d.nz * d.voxel_size[2];
            break; }
        default: return false;
    }
    for (int k = 0; k < 3; k++) cen[k] = 0.5f * (box.lo[k] + box.hi[k]);
    return true;
}

YCGE_API int ycge_scene_upload(ycge_ctx *c, const ycge_scene *s) try {
    if (c && c->group) { int rc = group_each_front(c, [&](ycge_ctx *f) { return ycge_scene_upload(f, s); }); c->have_scene = rc == 0; return rc; }
    if (!c || !s) return fail(c, YCGE_ERR_INVALID, "ctx/scene is NULL");
    if (s->n_objects < 0 || s->n_lights < 0 || s->n_materials < 0) return fail(c, YCGE_ERR_INVALID, "negative count");
    if ((s->n_objects > 0 && !s->objects) || (s->n_lights > 0 && !s->lights) || (s->n_materials > 0 && !s->materials)) return fail(c, YCGE_ERR_INVALID, "a count is positive but its array is NULL");
    CK(c, cudaSetDevice(c->device));
    CK(c, sync_ctx_streams(c));
    c->have_scene = false;
    for (auto &kv : c->meshes) { int rc = resolve_mesh(c, *kv.second); if (rc) return rc; } // device builds in flight: their root boxes are needed now
    // materials: the scene's table, then one entry per uploaded mesh
    std::vector<float4> mats;
    std::vector<DevTexture> dtex;
    std::map<int, int> tex_slot;
    bool tex_missing = false;
    auto push_mat = [&](const ycge_material &m) {
        mats.push_back(make_float4(m.albedo[0], m.albedo[1], m.albedo[2], m.reflectivity));
        mats.push_back(make_float4(m.emission[0], m.emission[1], m.emission[2], m.transparency));
        mats.push_back(make_float4(m.transmission[0], m.transmission[1], m.transmission[2], m.ior));
        int tex = -1; // slot in the device texture table (DiffuseTexture == null: -1)
        if (m.tex_id >= 0) {
            auto it = c->textures.find(m.tex_id);
            if (it == c->textures.end()) tex_missing = true;
            else {
                if (!tex_slot.count(m.tex_id)) {
                    tex_slot[m.tex_id] = (int)dtex.size();
                    dtex.push_back(DevTexture{it->second->px.p, it->second->w, it->second->h, 0});
                }
                tex = tex_slot[m.tex_id];
            }
        }
        float texf; memcpy(&texf, &tex, 4);
        mats.push_back(make_float4(m.specular, texf, m.tex_weight, m.uv_scale));
    };
    for (int i = 0; i < s->n_materials; i++) push_mat(s->materials[i]);
    std::map<int, int> mesh_slot, vol_slot;
    std::vector<DevMesh> dmeshes;
    std::vector<DevVolume> dvols;
    std::vector<DevObject> dobjs((size_t)s->n_objects);
    std::vector<BuildItem> items;
    for (int i = 0; i < s->n_objects; i++) {
        const ycge_object &o = s->objects[i];
        DevObject d;
        memset(&d, 0, sizeof d);
        d.kind = o.kind; d.mat_a = o.mat_a; d.mat_b = o.mat_b; d.override_sr = o.override_sr;
        d.checker_scale = o.checker_scale; d.specular = o.specular; d.reflectivity = o.reflectivity; d.ref = -1;
        memcpy(d.p, o.p, sizeof d.p);
        const float *p = o.p;
        bool needs_mat = true;
        switch (o.kind) {
            case YCGE_PLANE: d.d[0] = p[3] * p[0] + p[4] * p[1] + p[5] * p[2]; break;                     // Surfaces.cs:26
            case YCGE_DISK: d.d[0] = p[3] * p[0] + p[4] * p[1] + p[5] * p[2]; d.d[1] = p[6] * p[6]; break;  // Surfaces.cs:92-93 (Normal.Dot(Center))
            case YCGE_CYLINDER_Y: d.d[0] = p[3] * p[3]; break;                                           // BoundedObjects.cs:136
            case YCGE_TRIANGLE: { // Triangle.cs:36-44
                float e1x = p[3] - p[0], e1y = p[4] - p[1], e1z = p[5] - p[2], e2x = p[6] - p[0], e2y = p[7] - p[1], e2z = p[8] - p[2];
                float nnx = e1y * e2z - e1z * e2y, nny = e1z * e2x - e1x * e2z, nnz = e1x * e2y - e1y * e2x;
                float invLen = 1.0f / detail::net_max(1e-20f, std::sqrt(nnx * nnx + nny * nny + nnz * nnz));
                d.d[0] = e1x; d.d[1] = e1y; d.d[2] = e1z; d.d[3] = e2x; d.d[4] = e2y; d.d[5] = e2z;
                d.d[6] = nnx * invLen; d.d[7] = nny * invLen; d.d[8] = nnz * invLen;
                break; }
            case YCGE_MESH: {
                auto it = c->meshes.find(o.ref_id);
                if (it == c->meshes.end()) return fail(c, YCGE_ERR_INVALID, "object references a mesh id that was not uploaded");
                if (!mesh_slot.count(o.ref_id)) {
                    MeshStore &m = *it->second;
                    DevMesh dm;
                    dm.nodes = m.nodes.p; dm.tris = m.tris.p; dm.tri_id = m.tri_id.p; dm.root = m.root; dm.n_tris = m.n_tris;
                    dm.material = (int)(mats.size() / 4);
                    push_mat(m.material);
                    mesh_slot[o.ref_id] = (int)dmeshes.size();
                    dmeshes.push_back(dm);
                }
                d.ref = mesh_slot[o.ref_id];
                needs_mat = false;
                break; }
            case YCGE_VOLUME: {
                auto it = c->volumes.find(o.ref_id);
                if (it == c->volumes.end()) return fail(c, YCGE_ERR_INVALID, "object references a volume id that was not uploaded");
                if (!vol_slot.count(o.ref_id)) { vol_slot[o.ref_id] = (int)dvols.size(); dvols.push_back(it->second->dv); }
                d.ref = vol_slot[o.ref_id];
                for (int mi : it->second->palette) if (mi >= s->n_materials) return fail(c, YCGE_ERR_INVALID, "voxel palette references a material outside the scene table");
                if (it->second->def >= s->n_materials) return fail(c, YCGE_ERR_INVALID, "voxel palette default references a material outside the scene table");
                needs_mat = false;
                break; }
            default: break;
        }
        if (o.kind < 0 || o.kind > YCGE_VOLUME) return fail(c, YCGE_ERR_INVALID, "unknown object kind");
        if (needs_mat && (o.mat_a < 0 || o.mat_a >= s->n_materials || o.mat_b < 0 || o.mat_b >= s->n_materials))
            return fail(c, YCGE_ERR_INVALID, "object material index out of range");
        dobjs[i] = d;
        if (!s->bvh) {
            BuildItem it;
            it.index = i;
            if (!object_bounds(c, o, it.box, it.c)) return fail(c, YCGE_ERR_UNBOUNDED, "Unbounded Hittable");
            items.push_back(it);
        }
    }
    std::vector<PairNode> pairs;
    std::vector<int> leaf;
    TreeRoot root;
    if (s->bvh) {
        TreeView tv = view_of(*s->bvh);
        int rc = flatten(c, tv, pairs, root);
        if (rc) return rc;
        leaf.assign(tv.leaf, tv.leaf + tv.n_leaf);
    } else { // new BVH(Objects)  BVH.cs:29-97
        FlatTree tree;
        build_reference_tree(items, 4, false, tree);
        int rc = flatten(c, view_of(tree), pairs, root);
        if (rc) return rc;
        leaf = tree.leaf_index;
    }
    for (int v : leaf) if (v < 0 || v >= s->n_objects) return fail(c, YCGE_ERR_INVALID, "leafObjIndex out of range");
    if (tex_missing) return fail(c, YCGE_ERR_INVALID, "material references a texture id that was not uploaded (ycge_texture_upload)");
    std::vector<DevLight> lights((size_t)s->n_lights);
    for (int i = 0; i < s->n_lights; i++) {
        for (int k = 0; k < 3; k++) { lights[i].pos[k] = s->lights[i].pos[k]; lights[i].color[k] = s->lights[i].color[k]; }
        lights[i].intensity = s->lights[i].intensity; lights[i].pad = 0.0f;
    }
    CK(c, c->s_nodes.upload(pairs, c->stream));
    CK(c, c->s_leaf.upload(leaf, c->stream));
    CK(c, c->s_objects.upload(dobjs, c->stream));
    CK(c, c->s_materials.upload(mats, c->stream));
    CK(c, c->s_meshes.upload(dmeshes, c->stream));
    CK(c, c->s_volumes.upload(dvols, c->stream));
    CK(c, c->s_lights.upload(lights, c->stream));
    CK(c, c->s_textures.upload(dtex, c->stream));
    DevScene &ds = c->ds;
    memset(&ds, 0, sizeof ds);
    ds.textures = c->s_textures.p; ds.n_textures = (int)dtex.size();
    ds.nodes = c->s_nodes.p; ds.leaf_obj = c->s_leaf.p; ds.objects = c->s_objects.p; ds.materials = c->s_materials.p;
    ds.meshes = c->s_meshes.p; ds.volumes = c->s_volumes.p; ds.lights = c->s_lights.p; ds.root = root;
    ds.n_lights = s->n_lights; ds.is_volume_scene = s->is_volume_scene;
    for (int k = 0; k < 3; k++) { ds.bg_top[k] = s->bg_top[k]; ds.bg_bottom[k] = s->bg_bottom[k]; ds.ambient[k] = s->ambient_color[k]; }
    ds.ambient_intensity = s->ambient_intensity;
    { // lean scene: nothing that needs the DDA, the texture sampler or the deferred-branch stack (mats[4 m + 1].w = transparency)
        bool lean = dvols.empty() && dtex.empty() && !s->is_volume_scene && !getenv("YCGE_NO_LEAN");
        for (size_t m = 0; lean && m + 3 < mats.size(); m += 4) if (mats[m + 1].w > 0.0f) lean = false;
        c->trace_lean = lean;
    }
    c->have_scene = true;
    return 0;
} YCGE_CATCH

YCGE_API int ycge_lights_update(ycge_ctx *c, int32_t n, const ycge_light *l) try {
    if (c && c->group) return group_each_front(c, [&](ycge_ctx *f) { return ycge_lights_update(f, n, l); });
    if (!c || n < 0 || (n > 0 && !l)) return fail(c, YCGE_ERR_INVALID, "bad argument");
    CK(c, cudaSetDevice(c->device));
    std::vector<DevLight> lights((size_t)n);
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) { lights[i].pos[k] = l[i].pos[k]; lights[i].color[k] = l[i].color[k]; }
        lights[i].intensity = l[i].intensity; lights[i].pad = 0.0f;
    }
    CK(c, sync_ctx_streams(c));
    if ((size_t)n > c->s_lights.n) CK(c, c->s_lights.alloc((size_t)n));
    if (n) CK(c, cudaMemcpyAsync(c->s_lights.p, lights.data(), (size_t)n * sizeof(DevLight), cudaMemcpyHostToDevice, c->stream));
    CK(c, sync_ctx_streams(c));
    c->ds.lights = c->s_lights.p; c->ds.n_lights = n;
    return 0;
} YCGE_CATCH

YCGE_API int ycge_globals_update(ycge_ctx *c, const float bg_top[3], const float bg_bottom[3], const float ambient_color[3], float ambient_intensity) try {
    if (c && c->group) return group_each_front(c, [&](ycge_ctx *f) { return ycge_globals_update(f, bg_top, bg_bottom, ambient_color, ambient_intensity); });
    if (!c || !bg_top || !bg_bottom || !ambient_color) return fail(c, YCGE_ERR_INVALID, "bad argument");
    for (int k = 0; k < 3; k++) { c->ds.bg_top[k] = bg_top[k]; c->ds.bg_bottom[k] = bg_bottom[k]; c->ds.ambient[k] = ambient_color[k]; }
    c->ds.ambient_intensity = ambient_intensity;
    return 0;
} YCGE_CATCH

// ---- per frame ---------------------------------------------------------------------------------------------
YCGE_API int ycge_set_camera(ycge_ctx *c, const float pos[3], float yaw, float pitch) try {
    if (c && c->group) { if (!pos) return fail(c, YCGE_ERR_INVALID, "bad argument"); return group_each_front(c, [&](ycge_ctx *f) { return ycge_set_camera(f, pos, yaw, pitch); }); }
    if (!c || !pos) return fail(c, YCGE_ERR_INVALID, "bad argument");
    c->cam[0] = pos[0]; c->cam[1] = pos[1]; c->cam[2] = pos[2]; c->yaw = yaw; c->pitch = pitch;
    return 0;
} YCGE_CATCH
YCGE_API int ycge_set_trace_variant(ycge_ctx *c, int32_t variant) try {
    if (c && c->group) return group_each_front(c, [&](ycge_ctx *f) { return ycge_set_trace_variant(f, variant); });
    if (!c || variant < 0 || variant > 1) return fail(c, YCGE_ERR_INVALID, "trace variant must be 0 (thread per pixel path) or 1 (ray stream)");
    c->trace_variant = variant;
    return 0;
} YCGE_CATCH
YCGE_API int ycge_set_inplace_variant(ycge_ctx *c, int32_t variant) try {
    if (c && c->group) { for (ycge_ctx *b : c->group->back) { int rc = ycge_set_inplace_variant(b, variant); if (rc) return rc; } return 0; }
    if (!c || variant < 0 || variant > 1) return fail(c, YCGE_ERR_INVALID, "in-place variant must be 0 (one warp per chain) or 1 (systolic bands)");
    if (c->frame_open) return fail(c, YCGE_ERR_INVALID, "ycge_set_inplace_variant inside a frame");
    c->use_wave = variant == 1;
    return 0;
} YCGE_CATCH
YCGE_API int ycge_set_fov(ycge_ctx *c, float fov_deg) {
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    if (c->group) return group_each_front(c, [&](ycge_ctx *f) { return ycge_set_fov(f, fov_deg); });
    c->fov = fov_deg; return 0;
}
YCGE_API int ycge_reset_history(ycge_ctx *c) {
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    if (c->group) return group_each_front(c, [&](ycge_ctx *f) { return ycge_reset_history(f); });
    c->force_reset = true; return 0;
}

YCGE_API int ycge_frame_begin(ycge_ctx *c) {
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    if (c->group) return fail(c, YCGE_ERR_INVALID, "ycge_frame_begin is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    return frame_begin_impl(c);
}
YCGE_API int ycge_frame_finish(ycge_ctx *c) {
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    if (c->group) return fail(c, YCGE_ERR_INVALID, "ycge_frame_finish is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    return frame_finish_impl(c);
}
YCGE_API int ycge_frame_halo(ycge_ctx *c, ycge_halo *h) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_frame_halo is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || !h) return fail(c, YCGE_ERR_INVALID, "bad argument");
    memset(h, 0, sizeof *h);
    if (!c->frame_open || !c->dn.pending) return 0;
    int lo, a, slo, sa;
    halo_rows(c, lo, a, slo, sa);
    const int X = c->dn.cur_id, Y = (X == 1) ? 2 : 1;
    float4 *nw = c->dn.phys[Y];
    const size_t row = (size_t)c->W * sizeof(float4);
    if (c->peers) return 1; // the wavefront kernels exchange the rows themselves
    if (a > lo) { h->recv_ptr = nw + (size_t)lo * c->W; h->recv_bytes = (size_t)(a - lo) * row; h->recv_row0 = lo; h->recv_rows = a - lo; }
    if (sa > slo) { h->send_ptr = nw + (size_t)slo * c->W; h->send_bytes = (size_t)(sa - slo) * row; h->send_row0 = slo; h->send_rows = sa - slo; }
    return 1;
} YCGE_CATCH
YCGE_API int ycge_stash_config(ycge_ctx *c, int32_t n_slots) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_stash_config is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || n_slots < 0 || n_slots > 64) return fail(c, YCGE_ERR_INVALID, "bad argument");
    CK(c, cudaSetDevice(c->device));
    CK(c, sync_ctx_streams(c));
    c->stash.clear();
    const size_t rows = (size_t)c->tile_rows * 2 * c->ss;
    for (int k = 0; k < n_slots; k++) {
        std::unique_ptr<ycge_ctx::Stash> st(new ycge_ctx::Stash());
        CK(c, st->den.alloc(rows * c->W));
        CK(c, st->logs.alloc((size_t)c->sw * c->sh));
        c->stash.push_back(std::move(st));
    }
    return 0;
} YCGE_CATCH
YCGE_API int ycge_frame_stash(ycge_ctx *c, int32_t slot) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_frame_stash is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || slot < 0 || slot >= (int)c->stash.size()) return fail(c, YCGE_ERR_INVALID, "bad stash slot");
    if (!c->frame_open || c->dn.it <= c->dn.K) return fail(c, YCGE_ERR_INVALID, "no finished frame to stash");
    ycge_ctx::Stash &st = *c->stash[slot];
    const size_t row0 = (size_t)c->tile_row0 * 2 * c->ss, rows = (size_t)c->tile_rows * 2 * c->ss;
    CK(c, cudaMemcpyAsync(st.den.p, c->denoised + row0 * c->W, rows * c->W * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    CK(c, cudaMemcpyAsync(st.logs.p, c->logs.p, st.logs.n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    memcpy(c->last_cam, c->snap_cam, sizeof c->last_cam); c->last_yaw = c->snap_yaw; c->last_pitch = c->snap_pitch; // taa.CommitCamera
    c->frame_open = false;
    return 0;
} YCGE_CATCH
YCGE_API int ycge_frame_finish_stashed(ycge_ctx *c, int32_t slot, void *cuda_stream) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_frame_finish_stashed is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || slot < 0 || slot >= (int)c->stash.size()) return fail(c, YCGE_ERR_INVALID, "bad stash slot");
    ycge_ctx::Stash &st = *c->stash[slot];
    const size_t row0 = (size_t)c->tile_row0 * 2 * c->ss;
    return finish_launch(c, st.logs.p, st.den.p - row0 * c->W, (cudaStream_t)cuda_stream, false); // cells_kernel indexes rows absolutely
} YCGE_CATCH
YCGE_API int ycge_stash_logs_ptr(ycge_ctx *c, int32_t slot, void **ptr, size_t *bytes) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_stash_logs_ptr is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || !ptr || !bytes || slot < 0 || slot >= (int)c->stash.size()) return fail(c, YCGE_ERR_INVALID, "bad argument");
    *ptr = c->stash[slot]->logs.p; *bytes = c->stash[slot]->logs.n * sizeof(float);
    return 0;
} YCGE_CATCH
YCGE_API int ycge_peer_export(ycge_ctx *c, ycge_peer *out) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_peer_export is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || !out) return fail(c, YCGE_ERR_INVALID, "bad argument");
    CK(c, cudaSetDevice(c->device));
    memset(out, 0, sizeof *out);
    out->sa = c->sa.p; out->sb = c->sb.p; out->flags = c->flags.p;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ycge_peer assumes 64-byte IPC handles");
    CK(c, cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->sa_ipc, c->sa.p));
    CK(c, cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->sb_ipc, c->sb.p));
    CK(c, cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->flags_ipc, c->flags.p));
    return 0;
} YCGE_CATCH
YCGE_API int ycge_peer_attach(ycge_ctx *c, const ycge_peer *above, const ycge_peer *below, int32_t via_ipc) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_peer_attach is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    if (!c->sharded) return fail(c, YCGE_ERR_INVALID, "peers are for row-tile contexts");
    { // a tile forwards nothing: the rows the rank below needs must be rows this rank computes itself
        int need = 0;
        for (int it = 1; it < std::max(1, c->P.atrous_iterations); it += 2) need = 2 << it; // in-place passes are the odd iterations
        if (below && c->tile_rows * 2 * c->ss < need)
            return fail(c, YCGE_ERR_INVALID, "tile too small for the peer hand-off (fewer pixel rows than the in-place pass reaches); use the send/recv hand-off of ycge_frame_halo");
    }
    CK(c, cudaSetDevice(c->device));
    CK(c, sync_ctx_streams(c));
    for (auto &p : c->ipc_opened) if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
    c->has_above = above != nullptr; c->has_below = below != nullptr;
    c->above_flags = nullptr; c->below_sa = c->below_sb = nullptr;
    auto open = [&](const unsigned char *h, void *raw, int slot, void **out) -> int {
        if (!via_ipc) { *out = raw; return 0; }
        cudaIpcMemHandle_t mh;
        memcpy(&mh, h, sizeof mh);
        CK(c, cudaIpcOpenMemHandle(out, mh, cudaIpcMemLazyEnablePeerAccess));
        c->ipc_opened[slot] = *out;
        return 0;
    };
    int rc;
    if (above) { void *p; if ((rc = open(above->flags_ipc, above->flags, 0, &p))) return rc; c->above_flags = (int *)p; }
    if (below) {
        void *p;
        if ((rc = open(below->sa_ipc, below->sa, 1, &p))) return rc; c->below_sa = (float4 *)p;
        if ((rc = open(below->sb_ipc, below->sb, 2, &p))) return rc; c->below_sb = (float4 *)p;
    }
    c->peers = true;
    return 0;
} YCGE_CATCH
YCGE_API int ycge_frame_inplace(ycge_ctx *c) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_frame_inplace is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    if (!c->frame_open || !c->dn.pending) return fail(c, YCGE_ERR_INVALID, "no in-place pass is pending");
    return denoise_run(c);
} YCGE_CATCH

// ---- frame-parallel sharding: FRONT on row tiles, BACK of whole frames round-robin over the ranks ------------------
YCGE_API int ycge_frame_front(ycge_ctx *c) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_frame_front is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    if (c->peers) return fail(c, YCGE_ERR_INVALID, "a ctx with attached peers runs whole frames (ycge_frame_begin)");
    return frame_begin_impl(c, true);
} YCGE_CATCH
YCGE_API int ycge_back_config(ycge_ctx *c, int32_t n_slots) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_back_config is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || n_slots < 0 || n_slots > 16) return fail(c, YCGE_ERR_INVALID, "n_slots must be in [0,16]");
    if (c->sharded) return fail(c, YCGE_ERR_INVALID, "the BACK runs on a whole-frame ctx");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaDeviceSynchronize());
    return alloc_back_slots(c, n_slots);
} YCGE_CATCH
YCGE_API int ycge_back_ptr(ycge_ctx *c, int32_t slot, int32_t kind, void **ptr, size_t *bytes) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_back_ptr is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || !ptr || !bytes || slot < 0 || slot >= (int)c->back_slots.size()) return fail(c, YCGE_ERR_INVALID, "bad argument");
    ycge_ctx::BackSlot &b = *c->back_slots[slot];
    switch (kind) {
        case YCGE_PTR_CELLS: *ptr = b.cells.p; *bytes = b.cells.n * sizeof(ycge_cell); return 0;
        case YCGE_PTR_LOG_SAMPLES: *ptr = b.logs.p; *bytes = b.logs.n * sizeof(float); return 0;
        case YCGE_PTR_HIST: *ptr = b.hist.p; *bytes = b.hist.n * sizeof(float4); return 0;
        case YCGE_PTR_GND: *ptr = b.gnd.p; *bytes = b.gnd.n * sizeof(float4); return 0;
        case YCGE_PTR_GAS: *ptr = b.gas.p; *bytes = b.gas.n * sizeof(float4); return 0;
        default: return fail(c, YCGE_ERR_INVALID, "unknown pointer kind");
    }
} YCGE_CATCH
YCGE_API int ycge_back_denoise(ycge_ctx *c, int32_t slot, void *cuda_stream) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_back_denoise is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || slot < 0 || slot >= (int)c->back_slots.size()) return fail(c, YCGE_ERR_INVALID, "bad back slot");
    if (c->frame_open) return fail(c, YCGE_ERR_INVALID, "a frame is open on this ctx");
    CK(c, cudaSetDevice(c->device));
    ycge_ctx::BackSlot &b = *c->back_slots[slot];
    ycge_ctx::IoOverride io = {b.gnd.p, b.gas.p, &b.pre, b.logs.p, b.logs.n, (cudaStream_t)cuda_stream, 32 + slot};
    ycge_ctx::Denoise &d = c->dn;
    const int K = std::max(1, c->P.atrous_iterations);
    d.phys[0] = b.hist.p; d.phys[1] = b.sa.p; d.phys[2] = b.sb.p;
    d.cur_id = 0; d.dst_id = 1; d.it = 0; d.K = K; d.parity = 0; d.ty0 = 0; d.ty1 = c->H; d.pending = false; d.early_reset = false;
    for (int k = 0; k <= K; k++) d.halo_after[k] = 0;
    c->io = &io;
    c->launches_last = 0;
    int rc = denoise_run(c);
    c->io = nullptr;
    b.denoised = c->denoised;
    return rc;
} YCGE_CATCH
YCGE_API int ycge_back_finish(ycge_ctx *c, int32_t slot, void *cuda_stream) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_back_finish is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || slot < 0 || slot >= (int)c->back_slots.size()) return fail(c, YCGE_ERR_INVALID, "bad back slot");
    ycge_ctx::BackSlot &b = *c->back_slots[slot];
    if (!b.denoised) return fail(c, YCGE_ERR_INVALID, "ycge_back_finish before ycge_back_denoise");
    CK(c, cudaSetDevice(c->device));
    return finish_launch(c, b.logs.p, b.denoised, (cudaStream_t)cuda_stream, false, b.cells.p);
} YCGE_CATCH

YCGE_API int ycge_render_frame(ycge_ctx *c, ycge_cell *out, int32_t stride_cells) try {
    if (c && c->group) { // the synchronous drop-in call: one frame through the same path, then wait for its cells
        if (!out) return fail(c, YCGE_ERR_INVALID, "out is NULL");
        int64_t id = 0;
        int rc = group_submit(c, out, stride_cells, &id);
        return rc ? rc : group_frame_wait(c, id);
    }
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    if (c->sharded) return fail(c, YCGE_ERR_INVALID, "a row-tile ctx is driven with ycge_frame_begin / _halo / _inplace / _finish");
    int rc = frame_begin_impl(c);
    if (rc) return rc;
    rc = frame_finish_impl(c);
    if (rc) return rc;
    return read_cells_impl(c, out, stride_cells);
} YCGE_CATCH
YCGE_API int ycge_render_frame_stats(ycge_ctx *c, ycge_cell *out, int32_t stride_cells) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_render_frame_stats is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    c->want_stats = true;
    int rc = ycge_render_frame(c, out, stride_cells);
    c->want_stats = false;
    return rc;
} YCGE_CATCH
YCGE_API int ycge_render_frames_async(ycge_ctx *c, int32_t n) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_render_frames_async is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || n < 0) return fail(c, YCGE_ERR_INVALID, "bad argument");
    if (c->sharded) return fail(c, YCGE_ERR_INVALID, "a row-tile ctx is driven with ycge_frame_begin / _halo / _inplace / _finish");
    c->pipelined = c->n_slots >= 2;
    c->host_out = nullptr;
    int rc = 0;
    for (int i = 0; i < n && !rc; i++) {
        rc = frame_begin_impl(c);
        if (!rc) rc = frame_finish_impl(c);
    }
    c->pipelined = false;
    if (rc) return rc;
    // everything enqueued later on the ctx's stream (reads, synchronous frames, ycge_wait) comes after the last FINISH
    if (c->last_fin) { CK(c, cudaStreamWaitEvent(c->stream, c->last_fin, 0)); c->last_fin = nullptr; }
    return 0;
} YCGE_CATCH
YCGE_API int ycge_pipeline_config(ycge_ctx *c, int32_t n_slots) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_pipeline_config is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || n_slots < 1 || n_slots > 64) return fail(c, YCGE_ERR_INVALID, "n_slots must be in [1,64]");
    if (c->sharded && n_slots > 1) return fail(c, YCGE_ERR_INVALID, "a row-tile ctx pipelines over ranks (ycge_frame_stash), not over slots");
    if (c->frame_open) return fail(c, YCGE_ERR_INVALID, "a frame is open");
    if (n_slots == c->n_slots) return 0;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaDeviceSynchronize());
    c->in_flight.clear();
    // the guide set of the last frame (next frame's TAA reads it) moves to where the new numbering expects it
    const size_t px = (size_t)c->W * c->H;
    const int keep = (int)(c->frame_counter % std::max(2, (int)n_slots)); // the next frame writes set (frame_counter + 1) % G
    DevBuf<float4> tg, ta;
    CK(c, tg.alloc(px)); CK(c, ta.alloc(px));
    CK(c, cudaMemcpy(tg.p, gnd_of(c, c->last_gset), px * sizeof(float4), cudaMemcpyDeviceToDevice));
    CK(c, cudaMemcpy(ta.p, gas_of(c, c->last_gset), px * sizeof(float4), cudaMemcpyDeviceToDevice));
    int rc = alloc_slots(c, n_slots);
    if (rc) return rc;
    CK(c, cudaMemcpy(gnd_of(c, keep), tg.p, px * sizeof(float4), cudaMemcpyDeviceToDevice));
    CK(c, cudaMemcpy(gas_of(c, keep), ta.p, px * sizeof(float4), cudaMemcpyDeviceToDevice));
    c->last_gset = keep;
    return 0;
} YCGE_CATCH
/* Streaming path: SetCamera + TryFlipAndBlit without the wait.  The frame is enqueued (pipelined over the slots) and its
 * cells are copied into `out` (caller-owned, should be pinned) as part of the frame's FINISH; ycge_frame_wait(id) blocks
 * until that copy has landed.  At most n_slots frames may be un-waited. */
YCGE_API int ycge_submit_frame(ycge_ctx *c, ycge_cell *out, int32_t stride_cells, int64_t *frame_id) try {
    if (c && c->group) { if (!out) return fail(c, YCGE_ERR_INVALID, "bad argument"); return group_submit(c, out, stride_cells, frame_id); }
    if (!c || !out) return fail(c, YCGE_ERR_INVALID, "bad argument");
    if (c->sharded) return fail(c, YCGE_ERR_INVALID, "a row-tile ctx is driven with ycge_frame_begin / _halo / _inplace / _finish");
    if (stride_cells <= 0) stride_cells = c->fbW;
    if (stride_cells < c->fbW) return fail(c, YCGE_ERR_INVALID, "stride smaller than fb_w");
    if ((int)c->in_flight.size() >= c->n_slots) return fail(c, YCGE_ERR_LIMIT, "more un-waited frames than pipeline slots; call ycge_frame_wait");
    c->pipelined = c->n_slots >= 2;
    c->host_out = out; c->host_stride = stride_cells;
    int rc = frame_begin_impl(c);
    if (!rc) rc = frame_finish_impl(c);
    const bool piped = c->pipelined;
    c->pipelined = false; c->host_out = nullptr;
    if (rc) return rc;
    const SlotView sv = slot_view(c, c->cur_slot);
    if (!piped) { // one slot: same stream, same order, still asynchronous to the host
        CK(c, cudaMemcpy2DAsync(out, (size_t)stride_cells * sizeof(ycge_cell), c->cells.p, (size_t)c->fbW * sizeof(ycge_cell),
                                (size_t)c->fbW * sizeof(ycge_cell), (size_t)c->tile_rows, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaEventRecord(sv.host_done, c->stream));
    }
    c->in_flight.push_back(std::make_pair(c->frame_counter, sv.host_done));
    if (frame_id) *frame_id = c->frame_counter;
    return 0;
} YCGE_CATCH
YCGE_API int ycge_frame_wait(ycge_ctx *c, int64_t frame_id) try {
    if (c && c->group) return group_frame_wait(c, frame_id);
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    for (size_t i = 0; i < c->in_flight.size(); i++) {
        if (c->in_flight[i].first != frame_id) continue;
        CK(c, cudaEventSynchronize(c->in_flight[i].second));
        c->in_flight.erase(c->in_flight.begin(), c->in_flight.begin() + i + 1); // frames finish in order
        return wave_check(c);
    }
    return fail(c, YCGE_ERR_INVALID, "frame id is not in flight");
} YCGE_CATCH
YCGE_API int ycge_wait(ycge_ctx *c) try {
    if (c && c->group) return group_wait(c);
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    { int rc = join_pipeline(c); if (rc) return rc; }
    CK(c, sync_ctx_streams(c));
    return wave_check(c);
} YCGE_CATCH
YCGE_API int ycge_read_cells(ycge_ctx *c, ycge_cell *out, int32_t stride_cells) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_read_cells is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c) return fail(nullptr, YCGE_ERR_INVALID, "ctx is NULL");
    return read_cells_impl(c, out, stride_cells);
} YCGE_CATCH
YCGE_API int ycge_ansi_emit(ycge_ctx *c, uint8_t *out, size_t cap, size_t *n_bytes) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_ansi_emit is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || !out || !n_bytes) return fail(c, YCGE_ERR_INVALID, "bad argument");
    if (c->frame_counter == 0) return fail(c, YCGE_ERR_INVALID, "no frame rendered yet");
    CK(c, cudaSetDevice(c->device));
    { int rc = join_pipeline(c); if (rc) return rc; }
    cudaStream_t s = c->stream;
    const int rows = c->tile_rows, fbW = c->fbW;
    const size_t worst = 64 + (size_t)rows * 16 + (size_t)rows * fbW * 24; // 21 bytes of escape + 3 of glyph per cell at most
    if (c->ansi.n < worst) CK(c, c->ansi.alloc(worst));
    if (c->ansi_rows.n < (size_t)2 * rows + 1) CK(c, c->ansi_rows.alloc((size_t)2 * rows + 1));
    unsigned int *len = c->ansi_rows.p, *off = len + rows, *total = off + rows;
    const int wpb = 8;
    ansi_rows_kernel<false><<<div_up(rows, wpb), wpb * 32, 0, s>>>(c->cells.p, fbW, rows, c->tile_row0, len, nullptr, nullptr);
    ansi_scan_kernel<<<1, 1024, 0, s>>>(len, off, rows, c->ansi.p, total);
    ansi_rows_kernel<true><<<div_up(rows, wpb), wpb * 32, 0, s>>>(c->cells.p, fbW, rows, c->tile_row0, nullptr, off, c->ansi.p);
    CK(c, cudaGetLastError());
    unsigned int n = 0;
    CK(c, cudaMemcpyAsync(&n, total, sizeof n, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    *n_bytes = n;
    if (n > cap) return fail(c, YCGE_ERR_LIMIT, "ANSI stream larger than the caller's buffer");
    CK(c, cudaMemcpyAsync(out, c->ansi.p, n, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return 0;
} YCGE_CATCH
YCGE_API int ycge_device_ptr(ycge_ctx *c, int32_t kind, void **ptr, size_t *bytes) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_device_ptr is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || !ptr || !bytes) return fail(c, YCGE_ERR_INVALID, "bad argument");
    if (kind == YCGE_PTR_CELLS) { *ptr = c->cells.p; *bytes = c->cells.n * sizeof(ycge_cell); return 0; }
    if (kind == YCGE_PTR_LOG_SAMPLES) { *ptr = c->logs.p; *bytes = c->logs.n * sizeof(float); return 0; }
    if (kind == YCGE_PTR_HIST) { *ptr = c->hist.p; *bytes = c->hist.n * sizeof(float4); return 0; }
    if (kind == YCGE_PTR_GND) { *ptr = gnd_of(c, c->last_gset); *bytes = (size_t)c->W * c->H * sizeof(float4); return 0; }
    if (kind == YCGE_PTR_GAS) { *ptr = gas_of(c, c->last_gset); *bytes = (size_t)c->W * c->H * sizeof(float4); return 0; }
    if (kind == YCGE_PTR_EXPOSURE) { *ptr = c->expo.p; *bytes = sizeof(ExposureState); return 0; }
    return fail(c, YCGE_ERR_INVALID, "unknown pointer kind");
} YCGE_CATCH

// ---- introspection -----------------------------------------------------------------------------------------
YCGE_API int ycge_debug_read(ycge_ctx *c, int32_t kind, void *dst, size_t bytes) try {
    if (c && c->group) return fail(c, YCGE_ERR_INVALID, "ycge_debug_read is not available on a multi-GPU context (ycge_config.n_devices >= 2)");
    if (!c || !dst) return fail(c, YCGE_ERR_INVALID, "bad argument");
    CK(c, cudaSetDevice(c->device));
    { int rc = join_pipeline(c); if (rc) return rc; }
    CK(c, sync_ctx_streams(c));
    size_t n = (size_t)c->W * c->H;
    const void *src = nullptr;
    size_t need = 0;
    const int g = c->last_gset;
    switch (kind) {
        case YCGE_DBG_RAYS:
            if (!c->debug_rays) { // enable the tap; the next frame fills it
                CK(c, c->rays_dbg.alloc(n * 6));
                c->debug_rays = true;
                return fail(c, YCGE_ERR_INVALID, "ray tap enabled now; render a frame and read again");
            }
            src = c->rays_dbg.p; need = n * 24; break;
        case YCGE_DBG_HDR: src = c->cur.p; need = n * 16; break;
        case YCGE_DBG_ALBEDO_SKY: src = gas_of(c, g); need = n * 16; break;
        case YCGE_DBG_NORMAL_DEPTH: src = gnd_of(c, g); need = n * 16; break;
        case YCGE_DBG_TAA: src = c->hist.p; need = n * 16; break;
        case YCGE_DBG_DENOISED: src = c->denoised; need = n * 16; break;
        case YCGE_DBG_PRIM_ID: src = c->prim.p; need = n * 8; break;
        case YCGE_DBG_LOG_SAMPLES: src = slot_view(c, c->cur_slot).logs; need = c->logs.n * 4; break;
        default: return fail(c, YCGE_ERR_INVALID, "unknown debug kind");
    }
    if (!src) return fail(c, YCGE_ERR_INVALID, "no frame rendered yet");
    if (bytes < need) return fail(c, YCGE_ERR_INVALID, "destination too small");
    CK(c, cudaMemcpy(dst, src, need, cudaMemcpyDeviceToHost));
    return 0;
} YCGE_CATCH

YCGE_API int ycge_get_stats(ycge_ctx *c, ycge_stats *out) try {
    if (c && c->group) { // what a multi-GPU context can say: frames, rays traced by all fronts (halo rows included), the exposure state
        if (!out) return fail(c, YCGE_ERR_INVALID, "bad argument");
        memset(out, 0, sizeof *out);
        { int rc = group_wait(c); if (rc) return rc; }
        YcgeGroup &G = *c->group;
        out->frames = (uint64_t)G.frame;
        for (int g = 0; g < G.n; g++) {
            ycge_stats st;
            int rc = ycge_get_stats(G.front[g], &st);
            if (rc) return rc;
            out->rays += st.rays; out->rays_total += st.rays_total; out->kernel_launches += st.kernel_launches;
            if (st.ms_trace > out->ms_trace) out->ms_trace = st.ms_trace;
            if (st.ms_taa > out->ms_taa) out->ms_taa = st.ms_taa;
        }
        if (G.frame > 0) {
            ycge_stats st;
            int rc = ycge_get_stats(G.back[(int)((G.frame - 1) % G.n)], &st);
            if (rc) return rc;
            out->ae_exposure = st.ae_exposure; out->log_sum = st.log_sum; out->log_cnt = st.log_cnt; out->fast_div = st.fast_div;
        }
        return 0;
    }
    if (!c || !out) return fail(c, YCGE_ERR_INVALID, "bad argument");
    CK(c, cudaSetDevice(c->device));
    { int rc = join_pipeline(c); if (rc) return rc; }
    CK(c, sync_ctx_streams(c));
    memset(out, 0, sizeof *out);
    TraceCounters tc;
    CK(c, cudaMemcpy(&tc, c->counters_last ? c->counters_last : c->counters.p, sizeof tc, cudaMemcpyDeviceToHost));
    ExposureState es;
    CK(c, cudaMemcpy(&es, c->expo.p, sizeof es, cudaMemcpyDeviceToHost));
    TraceTotals tt;
    CK(c, cudaMemcpy(&tt, c->totals.p, sizeof tt, cudaMemcpyDeviceToHost));
    out->frames = (uint64_t)c->frame_counter; out->rays = tc.rays; out->rays_total = tt.rays_total;
    out->top_nodes_popped = tc.top_nodes; out->mesh_nodes_popped = tc.mesh_nodes; out->leaf_refs = tc.leaf_refs;
    out->tris_tested = tc.tris; out->prims_tested = tc.prims; out->dda_cells = tc.dda;
    if (tc.stack_overflow) return fail(c, YCGE_ERR_LIMIT, "traversal stack overflow (tree deeper than the device stack)");
    if (c->frame_counter > 0) {
        float ms[6];
        for (int k = 0; k < 6; k++) if (cudaEventElapsedTime(&ms[k], c->ev[k], c->ev[k + 1]) != cudaSuccess) ms[k] = 0.0f;
        out->ms_trace = ms[0]; out->ms_taa = ms[1]; out->ms_atrous = ms[2]; out->ms_exposure = ms[3] + ms[4]; out->ms_cells = ms[5];
        float tot = 0.0f;
        if (cudaEventElapsedTime(&tot, c->ev[0], c->ev[6]) == cudaSuccess) out->ms_total = tot;
        if (c->chain_timed && cudaEventElapsedTime(&tot, c->ev[7], c->ev[8]) == cudaSuccess) out->ms_atrous_chain = tot;
        (void)cudaGetLastError(); // an event that was never recorded (e.g. the finish events on the pipelined path) must not poison later checks
    }
    out->ae_exposure = es.ae_exposure; out->log_sum = es.log_sum; out->log_cnt = es.cnt;
    if (const char *path = getenv("YCGE_CHAIN_TRACE")) {
        if (c->chain_trace.n) {
            std::vector<unsigned long long> h(c->chain_trace.n);
            cudaMemcpy(h.data(), c->chain_trace.p, h.size() * 8, cudaMemcpyDeviceToHost);
            if (FILE *f = fopen(path, "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
        }
    }
    out->kernel_launches = c->launches_last;
    out->fast_div = c->fast_div ? 1 : 0;
    return 0;
} YCGE_CATCH

YCGE_API int ycge_get_frame_counter(ycge_ctx *c, int64_t *frame) try {
    if (c && c->group && frame) { *frame = c->group->frame; return 0; }
    if (!c || !frame) return fail(c, YCGE_ERR_INVALID, "bad argument");
    *frame = c->frame_counter;
    return 0;
} YCGE_CATCH

YCGE_API int ycge_rng_kat(ycge_ctx *c, int32_t which, int32_t n, const int32_t *x, const int32_t *y, const int64_t *frame, int32_t n_draws,
                          uint32_t *out_bits, uint64_t *out_seed) {
    if (!c || n <= 0 || n_draws <= 0 || !x || !y || !frame || !out_bits || !out_seed) return fail(c, YCGE_ERR_INVALID, "bad argument");
    CK(c, cudaSetDevice(c->device));
    DevBuf<int> dx, dy;
    DevBuf<long long> df;
    DevBuf<unsigned int> db;
    DevBuf<unsigned long long> dsd;
    CK(c, dx.alloc(n)); CK(c, dy.alloc(n)); CK(c, df.alloc(n)); CK(c, db.alloc((size_t)n * n_draws)); CK(c, dsd.alloc(n));
    CK(c, cudaMemcpy(dx.p, x, n * sizeof(int), cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(dy.p, y, n * sizeof(int), cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(df.p, frame, n * sizeof(long long), cudaMemcpyHostToDevice));
    rng_kat_kernel<<<div_up(n, 128), 128, 0, c->stream>>>(which, n, dx.p, dy.p, df.p, n_draws, db.p, dsd.p, c->P.seed_salt);
    CK(c, cudaGetLastError());
    CK(c, sync_ctx_streams(c));
    CK(c, cudaMemcpy(out_bits, db.p, (size_t)n * n_draws * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    CK(c, cudaMemcpy(out_seed, dsd.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}

} // extern "C"
