// ycge_multi.inl — several GPUs of one box behind ONE ycge_ctx (ycge_config.n_devices >= 2); included by ycge_lib.cu.
//
// Frames in parallel (DESIGN.md "Multi-GPU"): nothing of a frame's à-trous passes, exposure samples or cells feeds the next
// frame -- only the TAA history + guides (per pixel), the 16-byte exposure state and the camera memory do.  So per frame f
//   FRONT   every GPU g traces + TAA-blends its row tile (+1 halo row for the 3x3 luma clamp) on a row-tile context;
//   COPY    the tile's rows of (history, normal+depth, albedo+sky) go by cudaMemcpyPeerAsync over NVLink into a back slot
//           of GPU root = f mod N, on the FRONT's own stream (the next frame's TAA may then overwrite the history);
//   BACK    root runs the à-trous passes (incl. the wavefront) and the exposure samples of the whole frame on the slot's
//           stream, while all GPUs go on with the next fronts: N x slots frames are in flight;
//   FINISH  in frame order: root waits for the exposure state of frame f-1 (a 16-byte peer copy from root-1), runs the
//           ordered exposure sum + cells, copies the cells straight to the caller's host buffer and hands the state on.
// One host thread enqueues everything; cross-device ordering is by events only.  Bit-identical to a one-GPU context
// (tests/test_multigpu.py).  The children are ordinary contexts of this library (a row-tile ctx and a whole-frame ctx with
// back slots per GPU): the group adds no kernel.
struct YcgeGroup {
    static const int RING = 64; // per-frame events; more than any number of frames in flight
    int n = 0, S = 2;
    std::vector<int> dev;
    std::vector<ycge_ctx *> front, back;
    std::vector<int> row0, rows;                        // front tiles, cell rows
    std::vector<std::vector<cudaStream_t>> s_back;      // [gpu][slot]
    std::vector<cudaStream_t> s_fin;                    // [gpu]
    std::vector<std::vector<cudaEvent_t>> ev_copied, ev_expo, ev_host; // [gpu][ring]
    std::vector<std::vector<cudaEvent_t>> ev_back, ev_fin;             // [gpu][slot]
    std::vector<std::vector<char>> fin_used;                            // [gpu][slot]
    long long frame = 0;                                // frames submitted
    bool balanced = false;
    std::vector<std::pair<long long, std::pair<int, int>>> in_flight; // frame id -> (root, ring index)
    std::vector<ycge_cell> staging;                     // synchronous ycge_render_frame into pageable memory goes through here? no: direct
    ycge_config cfg;
    ~YcgeGroup() {
        for (size_t g = 0; g < dev.size(); g++) {
            cudaSetDevice(dev[g]);
            cudaDeviceSynchronize();
            if (g < front.size() && front[g]) ycge_destroy(front[g]);
            if (g < back.size() && back[g]) ycge_destroy(back[g]);
            if (g < s_back.size()) for (auto s : s_back[g]) if (s) cudaStreamDestroy(s);
            if (g < s_fin.size() && s_fin[g]) cudaStreamDestroy(s_fin[g]);
            for (auto *vv : {&ev_copied, &ev_expo, &ev_host, &ev_back, &ev_fin}) if (g < vv->size()) for (auto e : (*vv)[g]) if (e) cudaEventDestroy(e);
        }
    }
};

namespace {

#define GCK(c, call)                                                                                                  \
    do {                                                                                                              \
        cudaError_t e__ = (call);                                                                                     \
        if (e__ != cudaSuccess) return fail(c, YCGE_ERR_CUDA, std::string("multi-GPU: " #call ": ") + cudaGetErrorString(e__)); \
    } while (0)
#define GRC(c, call)                                                                                                  \
    do {                                                                                                              \
        int rc__ = (call);                                                                                            \
        if (rc__ != 0) return fail(c, rc__, std::string("multi-GPU: ") + ycge_last_error(nullptr));                   \
    } while (0)

void group_equal_tiles(YcgeGroup &G, int fb_h) {
    G.row0.assign(G.n, 0); G.rows.assign(G.n, 0);
    for (int g = 0; g < G.n; g++) { G.row0[g] = g * fb_h / G.n; G.rows[g] = (g + 1) * fb_h / G.n - G.row0[g]; }
}
int group_apply_tiles(ycge_ctx *c) { // the fronts allocate whole-frame planes: moving a tile is bookkeeping
    YcgeGroup &G = *c->group;
    for (int g = 0; g < G.n; g++) {
        ycge_ctx *f = G.front[g];
        f->tile_row0 = G.row0[g]; f->tile_rows = G.rows[g];
        f->sharded = true;
    }
    return 0;
}
void ctx_reset_state(ycge_ctx *f) { // back to "no frame rendered yet" (after the calibration frames of group_balance)
    f->frame_counter = 0; f->taa_valid = false; f->force_reset = false; f->frame_open = false; f->last_gset = 0;
    f->last_cam[0] = f->last_cam[1] = f->last_cam[2] = NAN; f->last_yaw = NAN; f->last_pitch = NAN;
    cudaSetDevice(f->device);
    cudaMemsetAsync(f->totals.p, 0, sizeof(TraceTotals), f->stream);
}

int group_create(const ycge_config *cfg, ycge_ctx **out) {
    if (cfg->tile_row0 != 0 || cfg->tile_rows != 0) return fail(nullptr, YCGE_ERR_INVALID, "a multi-GPU context owns the whole frame: tile_row0 / tile_rows must be 0");
    if (cfg->n_devices > 8) return fail(nullptr, YCGE_ERR_INVALID, "at most 8 devices");
    if (cfg->fb_h < cfg->n_devices) return fail(nullptr, YCGE_ERR_INVALID, "fewer cell rows than devices");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(nullptr, YCGE_ERR_CUDA, "no usable CUDA device (this library has no CPU path)");
    for (int a = 0; a < cfg->n_devices; a++) {
        if (cfg->devices[a] < 0 || cfg->devices[a] >= n_dev) return fail(nullptr, YCGE_ERR_INVALID, "device ordinal out of range");
    }
    std::unique_ptr<ycge_ctx> c(new ycge_ctx());
    c->group.reset(new YcgeGroup());
    YcgeGroup &G = *c->group;
    G.cfg = *cfg; G.n = cfg->n_devices;
    G.dev.assign(cfg->devices, cfg->devices + G.n);
    c->device = G.dev[0];
    c->P = cfg->params;
    c->fbW = cfg->fb_w; c->fbH = cfg->fb_h; c->ss = std::max(1, cfg->ss); c->W = c->fbW * c->ss; c->H = c->fbH * 2 * c->ss;
    c->tile_row0 = 0; c->tile_rows = c->fbH;
    for (int a = 0; a < G.n; a++) { // peer access, every ordered pair
        GCK(nullptr, cudaSetDevice(G.dev[a]));
        for (int b = 0; b < G.n; b++) {
            if (G.dev[a] == G.dev[b]) continue; // a device may be listed more than once (several tiles on one GPU: how the one-GPU test box exercises this path)
            int can = 0;
            GCK(nullptr, cudaDeviceCanAccessPeer(&can, G.dev[a], G.dev[b]));
            if (!can) return fail(nullptr, YCGE_ERR_CUDA, "multi-GPU: the listed devices cannot access each other's memory (peer access)");
            cudaError_t e = cudaDeviceEnablePeerAccess(G.dev[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(nullptr, YCGE_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            (void)cudaGetLastError();
        }
    }
    group_equal_tiles(G, c->fbH);
    G.front.assign(G.n, nullptr); G.back.assign(G.n, nullptr);
    G.s_back.resize(G.n); G.s_fin.assign(G.n, nullptr);
    G.ev_copied.resize(G.n); G.ev_expo.resize(G.n); G.ev_host.resize(G.n); G.ev_back.resize(G.n); G.ev_fin.resize(G.n); G.fin_used.resize(G.n);
    int lo = 0, hi = 0;
    for (int g = 0; g < G.n; g++) {
        ycge_config fc = *cfg;
        fc.n_devices = 0; fc.device = G.dev[g]; fc.tile_row0 = G.row0[g]; fc.tile_rows = G.rows[g];
        GRC(nullptr, ycge_create(&fc, &G.front[g]));
        G.front[g]->sharded = true; // also when one device holds every row: it is driven by ycge_frame_front
        ycge_config bc = *cfg;
        bc.n_devices = 0; bc.device = G.dev[g]; bc.tile_row0 = 0; bc.tile_rows = 0;
        GRC(nullptr, ycge_create(&bc, &G.back[g]));
        GRC(nullptr, ycge_back_config(G.back[g], G.S));
        GCK(nullptr, cudaSetDevice(G.dev[g]));
        GCK(nullptr, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        { // FRONT stream at high priority: short kernels every GPU's next front waits for (the BACKs of other frames are large grids)
            cudaStream_t s = nullptr;
            GCK(nullptr, cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
            if (G.front[g]->own_stream && G.front[g]->stream) cudaStreamDestroy(G.front[g]->stream);
            G.front[g]->stream = s; G.front[g]->own_stream = true;
        }
        G.s_back[g].assign(G.S, nullptr);
        for (int k = 0; k < G.S; k++) GCK(nullptr, cudaStreamCreateWithFlags(&G.s_back[g][k], cudaStreamNonBlocking));
        GCK(nullptr, cudaStreamCreateWithFlags(&G.s_fin[g], cudaStreamNonBlocking));
        auto mk = [&](std::vector<cudaEvent_t> &v, int count) -> int {
            v.assign(count, nullptr);
            for (int k = 0; k < count; k++) GCK(nullptr, cudaEventCreateWithFlags(&v[k], cudaEventDisableTiming));
            return 0;
        };
        if (mk(G.ev_copied[g], YcgeGroup::RING) || mk(G.ev_expo[g], YcgeGroup::RING) || mk(G.ev_host[g], YcgeGroup::RING) || mk(G.ev_back[g], G.S) || mk(G.ev_fin[g], G.S)) return YCGE_ERR_CUDA;
        G.fin_used[g].assign(G.S, 0);
    }
    *out = c.release();
    return 0;
}

// Tiles re-cut from measured trace times (sky rows end after one ray, mesh rows trace six): three calibration fronts with equal
// tiles, then every context goes back to "no frame rendered yet", so that the frames that follow are exactly frames 1, 2, ...
int group_balance(ycge_ctx *c) {
    YcgeGroup &G = *c->group;
    G.balanced = true;
    if (G.n < 2) return 0;
    std::vector<double> ms(G.n, 0.0);
    for (int k = 0; k < 3; k++) for (int g = 0; g < G.n; g++) GRC(c, ycge_frame_front(G.front[g]));
    for (int g = 0; g < G.n; g++) {
        GCK(c, cudaSetDevice(G.dev[g]));
        GCK(c, cudaStreamSynchronize(G.front[g]->stream));
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, G.front[g]->ev[0], G.front[g]->ev[1]) == cudaSuccess) ms[g] = t;
        (void)cudaGetLastError();
    }
    const int fb_h = c->fbH;
    std::vector<double> cost(fb_h, 0.0), cum(fb_h + 1, 0.0);
    for (int g = 0; g < G.n; g++) for (int r = G.row0[g]; r < G.row0[g] + G.rows[g]; r++) cost[r] = std::max(ms[g], 0.0) / std::max(G.rows[g], 1) + 0.001;
    for (int r = 0; r < fb_h; r++) cum[r + 1] = cum[r] + cost[r];
    std::vector<int> edges(1, 0);
    for (int g = 1; g < G.n; g++) {
        const double target = cum[fb_h] * g / G.n;
        int e = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        e = std::min(std::max(e, edges.back() + 1), fb_h - (G.n - g));
        edges.push_back(e);
    }
    edges.push_back(fb_h);
    for (int g = 0; g < G.n; g++) { G.row0[g] = edges[g]; G.rows[g] = edges[g + 1] - edges[g]; }
    group_apply_tiles(c);
    for (int g = 0; g < G.n; g++) ctx_reset_state(G.front[g]);
    return 0;
}

int group_submit(ycge_ctx *c, ycge_cell *out, int stride_cells, int64_t *frame_id) {
    YcgeGroup &G = *c->group;
    if (!c->have_scene) return fail(c, YCGE_ERR_NO_SCENE, "Scene BVH not built; call ycge_scene_upload() after populating the scene");
    if (stride_cells <= 0) stride_cells = c->fbW;
    if (stride_cells < c->fbW) return fail(c, YCGE_ERR_INVALID, "stride smaller than fb_w");
    if ((int)G.in_flight.size() >= G.n * G.S) return fail(c, YCGE_ERR_LIMIT, "more un-waited frames than back slots (2 per device); call ycge_frame_wait");
    if (!G.balanced) { int rc = group_balance(c); if (rc) return rc; }
    const long long f = G.frame;
    const int n = G.n, root = (int)(f % n), slot = (int)((f / n) % G.S), idx = (int)(f % YcgeGroup::RING);
    const int W = c->W, ss = c->ss;
    ycge_ctx *B = G.back[root];
    ycge_ctx::BackSlot &bs = *B->back_slots[slot];
    for (int g = 0; g < n; g++) {
        ycge_ctx *F = G.front[g];
        GRC(c, ycge_frame_front(F));
        GCK(c, cudaSetDevice(G.dev[g]));
        cudaStream_t s = F->stream;
        if (G.fin_used[root][slot]) GCK(c, cudaStreamWaitEvent(s, G.ev_fin[root][slot], 0)); // the slot's previous frame has left it
        const size_t off = (size_t)G.row0[g] * 2 * ss * W, cnt = (size_t)G.rows[g] * 2 * ss * W * sizeof(float4);
        GCK(c, cudaMemcpyPeerAsync(bs.hist.p + off, G.dev[root], F->hist.p + off, G.dev[g], cnt, s));
        GCK(c, cudaMemcpyPeerAsync(bs.gnd.p + off, G.dev[root], gnd_of(F, F->last_gset) + off, G.dev[g], cnt, s));
        GCK(c, cudaMemcpyPeerAsync(bs.gas.p + off, G.dev[root], gas_of(F, F->last_gset) + off, G.dev[g], cnt, s));
        GCK(c, cudaEventRecord(G.ev_copied[g][idx], s));
    }
    GCK(c, cudaSetDevice(G.dev[root]));
    cudaStream_t sb = G.s_back[root][slot], sf = G.s_fin[root];
    for (int g = 0; g < n; g++) GCK(c, cudaStreamWaitEvent(sb, G.ev_copied[g][idx], 0));
    GRC(c, ycge_back_denoise(B, slot, (void *)sb));
    GCK(c, cudaEventRecord(G.ev_back[root][slot], sb));
    GCK(c, cudaStreamWaitEvent(sf, G.ev_back[root][slot], 0));
    if (f > 0 && n > 1) { const int pr = (int)((f - 1) % n); GCK(c, cudaStreamWaitEvent(sf, G.ev_expo[pr][(int)((f - 1) % YcgeGroup::RING)], 0)); }
    GRC(c, ycge_back_finish(B, slot, (void *)sf));
    if (out) GCK(c, cudaMemcpy2DAsync(out, (size_t)stride_cells * sizeof(ycge_cell), bs.cells.p, (size_t)c->fbW * sizeof(ycge_cell), (size_t)c->fbW * sizeof(ycge_cell),
                                      (size_t)c->fbH, cudaMemcpyDeviceToHost, sf));
    GCK(c, cudaEventRecord(G.ev_host[root][idx], sf));
    if (n > 1) {
        const int nx = (root + 1) % n;
        GCK(c, cudaMemcpyPeerAsync(G.back[nx]->expo.p, G.dev[nx], B->expo.p, G.dev[root], sizeof(ExposureState), sf));
        GCK(c, cudaEventRecord(G.ev_expo[root][idx], sf));
    }
    GCK(c, cudaEventRecord(G.ev_fin[root][slot], sf));
    G.fin_used[root][slot] = 1;
    G.frame = f + 1;
    c->frame_counter = G.frame;
    G.in_flight.push_back(std::make_pair((long long)G.frame, std::make_pair(root, idx)));
    if (frame_id) *frame_id = G.frame;
    return 0;
}
int group_check(ycge_ctx *c) {
    YcgeGroup &G = *c->group;
    for (int g = 0; g < G.n; g++) { int rc = wave_check(G.front[g]); if (!rc) rc = wave_check(G.back[g]); if (rc) return fail(c, rc, G.front[g]->err.empty() ? G.back[g]->err : G.front[g]->err); }
    return 0;
}
int group_frame_wait(ycge_ctx *c, int64_t frame_id) {
    YcgeGroup &G = *c->group;
    for (size_t i = 0; i < G.in_flight.size(); i++) {
        if (G.in_flight[i].first != frame_id) continue;
        GCK(c, cudaSetDevice(G.dev[G.in_flight[i].second.first]));
        GCK(c, cudaEventSynchronize(G.ev_host[G.in_flight[i].second.first][G.in_flight[i].second.second]));
        G.in_flight.erase(G.in_flight.begin(), G.in_flight.begin() + i + 1); // frames finish in order
        return group_check(c);
    }
    return fail(c, YCGE_ERR_INVALID, "frame id is not in flight");
}
int group_wait(ycge_ctx *c) {
    YcgeGroup &G = *c->group;
    for (int g = 0; g < G.n; g++) { GCK(c, cudaSetDevice(G.dev[g])); GCK(c, cudaDeviceSynchronize()); }
    G.in_flight.clear();
    return group_check(c);
}
template <class Fn> int group_each_front(ycge_ctx *c, Fn fn) {
    YcgeGroup &G = *c->group;
    for (int g = 0; g < G.n; g++) { int rc = fn(G.front[g]); if (rc) return fail(c, rc, std::string("multi-GPU, device ") + std::to_string(G.dev[g]) + ": " + ycge_last_error(nullptr)); }
    return 0;
}

} // namespace
