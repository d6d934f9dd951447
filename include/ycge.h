/*
 * ycge.h — C ABI of the B200-native frame producer for YetAnotherConsoleGameEngine's
 * per-frame ray tracing path.
 *
 * The reference has no FFI for rendering; the seam this library sits behind is the
 * private interface RaytraceEntity.IConsoleRenderer (ConsoleGame/RaytraceEntity.cs:12-18):
 *     void SetCamera(Vec3 pos, float yaw, float pitch);   void SetFov(float fovDeg);
 *     void TryFlipAndBlit(Framebuffer fb);                void Resize(Framebuffer fb, int superSample);
 * A new C# class `CudaRaytraceWrapper : IConsoleRenderer` (host_cs/CudaRaytraceRenderer.cs,
 * INTEGRATION.md) P/Invokes the entry points below.  Plain pointers and sizes only; no
 * callbacks into managed code; every pointer argument is caller-owned and only read
 * during the call unless stated otherwise.
 *
 * All entry points return 0 on success or a negative ycge_status; the message is
 * available from ycge_last_error().  One ctx = one host thread at a time (the reference
 * calls its renderer from the single game-loop thread, Renderer/Terminal.cs:136-176).
 *
 * Paths are relative to /root/reference/ConsoleGame/ unless stated otherwise.
 */
#ifndef YCGE_H
#define YCGE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define YCGE_API __declspec(dllexport)
#else
#define YCGE_API __attribute__((visibility("default")))
#endif

typedef struct ycge_ctx ycge_ctx;

typedef enum ycge_status {
    YCGE_OK = 0,
    YCGE_ERR_INVALID = -1,   /* bad argument (ArgumentException in the reference) */
    YCGE_ERR_CUDA = -2,      /* CUDA runtime failure; ycge_last_error() carries the cudaError string */
    YCGE_ERR_NO_SCENE = -3,  /* render before ycge_scene_upload ("Scene BVH not built", Scenes/Scene.cs:73) */
    YCGE_ERR_UNBOUNDED = -4, /* "Unbounded Hittable" (Objects/BVH.cs:39) */
    YCGE_ERR_LIMIT = -5      /* a structural limit of the device layout was exceeded */
} ycge_status;

/* ---- Material (RayTracing/Material.cs:5-61).  The C# scalars are binary64 but are only
 * ever read through (float) casts or `> 0.0` tests on the path (RaytraceRenderer.cs:500-559,776),
 * so the float-rounded value is exact for the path. 64 bytes. */
typedef struct ycge_material {
    float albedo[3];
    float reflectivity;
    float emission[3];
    float transparency;
    float transmission[3];
    float ior;
    float specular;      /* never read by shading (kept for round trips) */
    int32_t tex_id;      /* -1: no DiffuseTexture */
    float tex_weight;
    float uv_scale;
} ycge_material;

/* ---- Top-level objects (RayTracing/Objects/{BoundedObjects,Surfaces,Triangle}.cs, Mesh.cs, VolumeGrid.cs).
 * `p` holds the public fields of the C# object after its constructor ran:
 *   SPHERE      center.xyz, radius                                   BoundedObjects.cs:8-18
 *   PLANE       point.xyz, normal.xyz (normalised by the ctor)        Surfaces.cs:9-28
 *   DISK        center.xyz, normal.xyz (normalised), radius           Surfaces.cs:73-94
 *   XYRECT      x0,x1,y0,y1,z                                         Surfaces.cs:144-171
 *   XZRECT      x0,x1,z0,z1,y                                         Surfaces.cs:216-243
 *   YZRECT      y0,y1,z0,z1,x                                         Surfaces.cs:288-315
 *   BOX         min.xyz, max.xyz                                      BoundedObjects.cs:72-89
 *   CYLINDER_Y  center.xyz, radius, yMin, yMax, capped(0/1)           BoundedObjects.cs:118-137
 *   TRIANGLE    A.xyz, B.xyz, C.xyz                                   Triangle.cs:10-34
 *   MESH        ref_id = id given to ycge_mesh_upload_*                Mesh.cs:9-24
 *   VOLUME      ref_id = id given to ycge_volume_upload                VolumeGrid.cs:55-93
 *
 * Material functions (Func<Vec3,Vec3,float,Material>) are data, not delegates: every lambda
 * in the reference is either a constant material or Checker(a,b,scale) (Scenes/Scenes.cs:408-428).
 *   checker_scale == 0  → constant: materials[mat_a]
 *   checker_scale != 0  → ((int)floor(P.x/scale) + (int)floor(P.z/scale)) & 1 ? materials[mat_b] : materials[mat_a]
 * override_sr = 1 for the flat primitives and Box, which overwrite Specular/Reflectivity of the
 * function's result with their own fields (Surfaces.cs:64-66,135-137,207-209,279-281,351-353). */
typedef enum ycge_object_kind {
    YCGE_SPHERE = 0, YCGE_PLANE = 1, YCGE_DISK = 2, YCGE_XYRECT = 3, YCGE_XZRECT = 4, YCGE_YZRECT = 5,
    YCGE_BOX = 6, YCGE_CYLINDER_Y = 7, YCGE_TRIANGLE = 8, YCGE_MESH = 9, YCGE_VOLUME = 10
} ycge_object_kind;

typedef struct ycge_object {
    int32_t kind;
    int32_t mat_a;
    int32_t mat_b;
    float checker_scale;
    int32_t override_sr;
    float specular;
    float reflectivity;
    int32_t ref_id;
    float p[12];
} ycge_object; /* 80 bytes */

typedef struct ycge_light { /* RayTracing/Objects/PointLight.cs */
    float pos[3];
    float color[3];
    float intensity;
} ycge_light;

/* ---- Flat binary BVH exactly as the reference keeps it (Objects/BVH.cs:11-25, MeshBVH.cs:18-39):
 * SoA node arrays, count>0 => leaf referencing leaf_index[start..start+count). */
typedef struct ycge_bvh {
    int32_t n_nodes;
    int32_t root;
    int32_t n_leaf_refs;
    const float *min_x, *min_y, *min_z, *max_x, *max_y, *max_z;
    const int32_t *left, *right, *start, *count;
    const int32_t *leaf_index;
} ycge_bvh;

typedef struct ycge_scene { /* Scenes/Scene.cs:12-26 */
    float bg_top[3];
    float bg_bottom[3];
    float ambient_color[3];
    float ambient_intensity;
    int32_t is_volume_scene; /* `scene is VolumeScene` (RaytraceRenderer.cs:761): binary shadow occlusion */
    int32_t n_lights;
    const ycge_light *lights;
    int32_t n_materials;
    const ycge_material *materials;
    int32_t n_objects;
    const ycge_object *objects; /* order = Scene.Objects enumeration order = primary-hit objId */
    const ycge_bvh *bvh;        /* optional: the host's own `new BVH(Objects)`; NULL => library builds it */
} ycge_scene;

/* ---- Triangle mesh.  Two forms:
 *  - SoA, mirroring MeshBVH's private arrays (MeshBVH.cs:32-39) with the host's tree (required);
 *  - plain triangles A,B,C (9 floats each, MeshLoader face order); the library derives
 *    e1,e2,n (MeshBVH.cs:83-97) and builds the SAH tree (MeshBVH.cs:371-576). */
typedef struct ycge_mesh_soa {
    int32_t n_tris;
    const float *ax, *ay, *az, *e1x, *e1y, *e1z, *e2x, *e2y, *e2z, *nx, *ny, *nz;
    ycge_material material;
    const ycge_bvh *bvh;
} ycge_mesh_soa;

/* ---- Voxel grid (Objects/VolumeGrid.cs:17-53).  mat/meta are the reference's pinned arrays
 * in bricked-Morton order (VolumeGrid.cs:235-252), capacity ceil(nx/8)*ceil(ny/8)*ceil(nz/8)*512.
 * materialLookup(id, meta) is a closed table: row = id (ids outside [0,n_ids) use default),
 * column = clamp(meta, 0, meta_levels-1) (Scenes/VoxelMaterialPalette.cs:48-98, Scenes/Scenes.cs:107-118). */
typedef struct ycge_volume {
    int32_t nx, ny, nz;
    float min_corner[3];
    float voxel_size[3];
    const int32_t *mat;
    const int32_t *meta;
    int32_t wireframe;
    float wire_width_frac;
    float wire_max_distance;
    int32_t palette_n_ids;
    int32_t palette_meta_levels;
    const int32_t *palette; /* [n_ids * meta_levels] material indices into ycge_scene.materials */
    int32_t palette_default;
} ycge_volume;

/* ---- Compile-time constants of the reference, as a POD (defaults = the reference's values;
 * parity is only claimed at defaults).  RaytraceRenderer.cs:31-43,65,218-226; ToneMapper.cs:8-21. */
typedef struct ycge_params {
    int32_t diffuse_bounces;     /* 1 */
    int32_t max_mirror_bounces;  /* 2 */
    int32_t max_refractions;     /* 2 */
    int32_t atrous_iterations;   /* 3 */
    float mirror_threshold;      /* 0.9 */
    float eps;                   /* 1e-4 */
    float taa_alpha;             /* 0.01 */
    float motion_trans_reset;    /* 0.0025 */
    float motion_rot_reset;      /* 0.0025 */
    float diffuse_sigma_deg;     /* 25 */
    float luminance_pad;         /* 0.10 */
    float c_phi, n_phi, z_phi, a_phi; /* 3, 0.35, 2, 0.2 */
    float tone_exposure;         /* 1 */
    float tone_gamma;            /* 2.2 */
    float ae_key, ae_speed, ae_min, ae_max; /* 0.18, 0.2, 0.10, 1.50 */
    float saturation, vibrance;  /* 2.0, 0.0 */
    int32_t auto_exposure;       /* 1 */
    uint64_t seed_salt;          /* 0x9E3779B97F4A7C15 */
} ycge_params;

typedef struct ycge_config {
    int32_t fb_w, fb_h, ss;   /* cells and supersample: hiW = fb_w*ss, hiH = fb_h*2*ss (RaytraceRenderer.cs:86-87) */
    int32_t device;           /* CUDA device ordinal */
    int32_t tile_row0;        /* first cell row owned by this ctx (row-tile sharding); 0 for a whole frame */
    int32_t tile_rows;        /* number of cell rows owned; 0 => fb_h - tile_row0 */
    ycge_params params;
    /* Several GPUs of one box behind ONE context ("device list 1/2/4/8").  n_devices >= 2: the library owns
     * devices[0 .. n_devices), enables peer access between them and renders frames in parallel: every GPU traces and
     * TAA-blends a row tile of each frame (scene replicated, tiles balanced by measured trace time), the tiles' history +
     * guide rows go by peer copies over NVLink to the GPU that runs this frame's a-trous passes, exposure and cell
     * conversion (round robin), the 16-byte exposure state travels GPU to GPU in frame order, and the cells land in the
     * caller's buffer -- bit-identical to a one-GPU context.  Such a context takes the scene / camera / light calls and
     * ycge_render_frame, ycge_submit_frame / ycge_frame_wait (up to 2 * n_devices frames in flight), ycge_wait,
     * ycge_get_stats, ycge_get_frame_counter.  n_devices 0 or 1: one GPU, `device`.  tile_row0 / tile_rows must be 0.  An ordinal may
     * be listed more than once (several tiles on one GPU; the test suite does that on a one-GPU box). */
    int32_t n_devices;
    int32_t devices[8];
} ycge_config;

/* ---- One console cell = one Chexel (Renderer/Chexel.cs:99-125) plus the two quantisations
 * its consumers derive from it: 32 bytes, row-major, y*fb_w + x. */
typedef struct ycge_cell {
    uint16_t glyph;   /* UTF-16 code unit; always U+2580 (RaytraceRenderer.cs:260) */
    uint8_t fg16;     /* ChexelColor.color_16 of fg = top half   (Chexel.cs:70-88) */
    uint8_t bg16;     /* ... of bg = bottom half */
    uint8_t fg_ansi;  /* ANSITerminalRenderer.ChexelToAnsi256 (ANSITerminalRenderer.cs:246-286) */
    uint8_t bg_ansi;
    uint16_t attr;    /* Win32TerminalRenderer.MapAttributes (Win32TerminalRenderer.cs:109-112) */
    float fg[3];      /* ChexelColor.color_f32 (gamma-encoded SDR, clamped) */
    float bg[3];
} ycge_cell;

typedef struct ycge_stats {
    uint64_t frames;          /* frames rendered since create */
    uint64_t rays;            /* Scene.Hit / Scene.Occluded invocations in the last frame */
    uint64_t rays_total;      /* the same, accumulated since ycge_create (never reset) */
    /* reference-defined traversal events of the last frame; filled only by ycge_render_frame_stats */
    uint64_t top_nodes_popped, mesh_nodes_popped, leaf_refs, tris_tested, prims_tested, dda_cells;
    float ms_trace, ms_taa, ms_atrous, ms_exposure, ms_cells, ms_total; /* CUDA-event times of the last timed frame */
    float ae_exposure;        /* ToneMapper.aeExposure after the last frame */
    float log_sum;            /* UpdateExposure's logSum of the last frame */
    int32_t log_cnt;
    int32_t kernel_launches;  /* kernels launched for the last frame */
    float ms_atrous_chain;    /* the wavefront kernel of the in-place à-trous pass alone (part of ms_atrous) */
    int32_t fast_div;         /* 1: the FMA division sequence passed its exhaustive check for this ctx's divisors */
} ycge_stats;

typedef enum ycge_debug_kind {
    YCGE_DBG_RAYS = 0,        /* hiW*hiH * 6 float: origin, dir (MakeJitteredRay) */
    YCGE_DBG_HDR = 1,         /* hiW*hiH * 4 float: currentHdr rgb + luma */
    YCGE_DBG_ALBEDO_SKY = 2,  /* hiW*hiH * 4 float: gAlbedo rgb + sky flag */
    YCGE_DBG_NORMAL_DEPTH = 3,/* hiW*hiH * 4 float: normalised gNormal + gDepth */
    YCGE_DBG_TAA = 4,         /* hiW*hiH * 4 float: taaHistory rgb + luma */
    YCGE_DBG_DENOISED = 5,    /* hiW*hiH * 4 float: à-trous output rgb + luma */
    YCGE_DBG_PRIM_ID = 6,     /* hiW*hiH * 2 int32: primary objId, subId (sky = -1,-1) */
    YCGE_DBG_LOG_SAMPLES = 7  /* ceil(hiW/step)*ceil(hiH/step) float: log(1e-6+lum) or NaN when skipped */
} ycge_debug_kind;

/* ---- lifecycle -------------------------------------------------------------------------- */
YCGE_API void ycge_default_params(ycge_params *out);
YCGE_API int ycge_create(const ycge_config *cfg, ycge_ctx **out);                 /* new RaytraceRenderer(...)  RaytraceRenderer.cs:74-108 */
YCGE_API void ycge_destroy(ycge_ctx *ctx);
YCGE_API const char *ycge_last_error(ycge_ctx *ctx);                               /* ctx may be NULL: thread-local message */
YCGE_API int ycge_resize(ycge_ctx *ctx, int32_t fb_w, int32_t fb_h, int32_t ss);  /* Resize  RaytraceRenderer.cs:110-138 (keeps frame counter + exposure) */

/* ---- scene ------------------------------------------------------------------------------ */
YCGE_API int ycge_mesh_upload_soa(ycge_ctx *ctx, int32_t id, const ycge_mesh_soa *mesh);          /* MeshBVH.cs:18-39 */
YCGE_API int ycge_mesh_upload_triangles(ycge_ctx *ctx, int32_t id, int32_t n_tris, const float *abc,
                                        const ycge_material *material);                       /* new MeshBVH(tris)  MeshBVH.cs:41-130 */
/* the same call, the tree built ON THE DEVICE (csrc/bvh_device.cuh): MeshBVH.BuildRecursive
 * (MeshBVH.cs:371-576) node for node -- binned SAH with the reference's bin mapping, its in-place two-pointer partition in
 * closed form, its Array.Sort fallback -- from a device-wide work queue, written straight in the device layout.  A scene
 * switch (RaytraceEntity.cs:234-246) then costs milliseconds of GPU time instead of the host build. */
YCGE_API int ycge_mesh_build_device(ycge_ctx *ctx, int32_t id, int32_t n_tris, const float *abc, const ycge_material *material);
/* Development / test aid: the stored device arrays of a mesh.  what: 0 = pair nodes (64 B each), 1 = triangles in leaf
 * order (48 B), 2 = leaf slot -> triangle index (4 B), 3 = root record (32 B).  dst == NULL: *bytes = size of the array. */
YCGE_API int ycge_mesh_debug_read(ycge_ctx *ctx, int32_t id, int32_t what, void *dst, size_t *bytes);
YCGE_API int ycge_volume_upload(ycge_ctx *ctx, int32_t id, const ycge_volume *vol);              /* new VolumeGrid(...)  VolumeGrid.cs:55-93 */
/* new Texture(path) (Renderer/Texture.cs:25-49, :81-90): `rgba` = the reference's int[] pixels (byte 0 = R, row-major, row 0
 * first), read during the call.  Materials refer to it through ycge_material.tex_id; upload before ycge_scene_upload.
 * Sampling is Texture.SampleBilinear (:143-162) inside SampleAlbedo (RaytraceRenderer.cs:724-735).  Static images only:
 * camera/video-backed textures (isDynamic) are outside the path (VideoRenderer). */
YCGE_API int ycge_texture_upload(ycge_ctx *ctx, int32_t id, int32_t w, int32_t h, const uint32_t *rgba);
YCGE_API int ycge_scene_upload(ycge_ctx *ctx, const ycge_scene *scene);                          /* Scene.RebuildBVH  Scenes/Scene.cs:66-69 */
YCGE_API int ycge_lights_update(ycge_ctx *ctx, int32_t n, const ycge_light *lights);             /* DayNightCycle.cs:80-83 */
YCGE_API int ycge_globals_update(ycge_ctx *ctx, const float bg_top[3], const float bg_bottom[3],
                                 const float ambient_color[3], float ambient_intensity);     /* DayNightCycle.cs:86-88 */

/* ---- per frame -------------------------------------------------------------------------- */
YCGE_API int ycge_set_camera(ycge_ctx *ctx, const float pos[3], float yaw, float pitch);         /* SetCamera  RaytraceRenderer.cs:140-148 */
YCGE_API int ycge_set_fov(ycge_ctx *ctx, float fov_deg);                                          /* SetFov     RaytraceRenderer.cs:150-153 */
YCGE_API int ycge_reset_history(ycge_ctx *ctx);                                                   /* scene.HasDynamicTextures / scene switch */
/* Two forms of the trace kernel with bit-identical results: 0 = one thread per pixel path, 1 (default) = ray stream (every
 * lane carries one ray per round through a single traversal; a warp takes its next 8x4 pixel tile when all of its lanes
 * are done -- refilling single lanes was measured slower, see csrc/trace_stream.cuh). */
YCGE_API int ycge_set_trace_variant(ycge_ctx *ctx, int32_t variant);
/* Two forms of the in-place a-trous iteration (RaytraceRenderer.cs:718) with bit-identical results: 0 = one warp per chain,
 * rows handed over through L2 (csrc/post.cuh); 1 (default) = systolic bands: one warp per 4 rows in lock step plus a halo
 * warp that brings the rows above (csrc/wavefront.cuh). */
YCGE_API int ycge_set_inplace_variant(ycge_ctx *ctx, int32_t variant);
/* TryFlipAndBlit (RaytraceRenderer.cs:157-267): synchronous; writes tile_rows*fb_w cells (the ctx's tile;
 * the whole frame when unsharded) into caller-owned host memory, row stride `stride_cells` (0 => fb_w). */
YCGE_API int ycge_render_frame(ycge_ctx *ctx, ycge_cell *out, int32_t stride_cells);
/* Same frame, but with the reference-defined traversal event counters filled (slower kernel variant). */
YCGE_API int ycge_render_frame_stats(ycge_ctx *ctx, ycge_cell *out, int32_t stride_cells);
/* Headless path: enqueue n frames on the ctx's stream without host synchronisation; cells of the
 * last frame stay in device memory (ycge_device_ptr(YCGE_PTR_CELLS)).  ycge_wait blocks until done. */
YCGE_API int ycge_render_frames_async(ycge_ctx *ctx, int32_t n);
YCGE_API int ycge_wait(ycge_ctx *ctx);
/* Frame pipelining on one GPU.  The reference computes a frame strictly after the previous one
 * (RaytraceRenderer.cs:157-267), but only three things actually flow from frame N to frame N+1: the TAA history with its
 * guides (:207-216), ToneMapper.aeExposure (ToneMapper.cs:85-90) and the camera memory (:266); the denoised image never
 * re-enters the history (:221-224).  With n_slots >= 2 the asynchronous entry points (ycge_render_frames_async,
 * ycge_submit_frame) therefore run the à-trous passes of up to n_slots frames concurrently with the trace/TAA of the
 * following ones (each slot owns its scratch buffers and a stream); the ordered exposure sum and the cell conversion
 * are chained frame to frame.  Results are bit-identical to n_slots = 1.  Default 1; costs ~0.45 KB per pixel and slot. */
YCGE_API int ycge_pipeline_config(ycge_ctx *ctx, int32_t n_slots);
/* SetCamera + TryFlipAndBlit without the wait: enqueues one frame with the current camera and copies its cells into `out`
 * (caller-owned, page-locked for a truly asynchronous copy) as part of that frame's work; *frame_id receives the frame
 * number.  ycge_frame_wait(id) returns once the cells of that frame have landed.  At most n_slots frames may be un-waited. */
YCGE_API int ycge_submit_frame(ycge_ctx *ctx, ycge_cell *out, int32_t stride_cells, int64_t *frame_id);
YCGE_API int ycge_frame_wait(ycge_ctx *ctx, int64_t frame_id);
/* Copy the device-resident cells of the last frame to host memory (what ycge_render_frame does at its end). */
YCGE_API int ycge_read_cells(ycge_ctx *ctx, ycge_cell *out, int32_t stride_cells);

/* ---- row-tile sharding (one process per GPU; the collectives are the caller's, e.g. torch.distributed/NCCL).
 * A ctx created with tile_row0/tile_rows renders one row tile (cell rows) of the frame plus the pixel-row halo its
 * image passes need (recomputed locally, bit-identical).  Per frame:
 *   ycge_frame_begin     trace + TAA + à-trous passes up to the first in-place pass (RaytraceRenderer.cs:718 makes
 *                        iteration 1 run in place: every row then depends on the rows above it, across tiles)
 *   while (ycge_frame_halo(ctx, &h) == 1) {
 *       <caller: receive h.recv_bytes into h.recv_ptr from the rank above  (the rows just above the tile's range)>
 *       ycge_frame_inplace   the wavefront over this tile's rows, then the following passes up to the next in-place
 *                            pass, or to the end: log-luminance samples of the tile's rows written into the full-frame
 *                            sample array (zeros elsewhere)
 *       <caller: send h.send_bytes from h.send_ptr to the rank below>
 *   }
 *   <caller: sum-all-reduce of YCGE_PTR_LOG_SAMPLES over ranks (each slot is owned by one rank)>
 *   ycge_frame_finish    ordered exposure sum (identical on every rank), cell conversion of the tile
 *   <caller: gather YCGE_PTR_CELLS tiles to rank 0>
 * An unsharded ctx runs everything inside ycge_frame_begin (ycge_frame_halo returns 0). All work is enqueued on the
 * ctx's stream (ycge_set_stream), so a transfer enqueued on the same stream is ordered with it. */
typedef struct ycge_halo {
    void *recv_ptr;  size_t recv_bytes;  int32_t recv_row0, recv_rows;  /* NULL/0 for the top tile */
    void *send_ptr;  size_t send_bytes;  int32_t send_row0, send_rows;  /* NULL/0 for the bottom tile; valid after ycge_frame_inplace */
} ycge_halo;
/* Peer hand-off (optional, replaces the caller's send/recv of ycge_frame_halo): every rank exports its two à-trous
 * scratch buffers and a flag word (raw device pointers for ranks in one process, CUDA IPC handles across processes) and
 * attaches to its neighbours.  The wavefront kernel of rank g then stores its boundary rows straight into rank g+1's
 * output buffer over NVLink, pixel by pixel, and rank g+1's wavefront polls them like any other row: the wavefront
 * crosses GPU boundaries without a kernel boundary.  With peers attached ycge_frame_halo reports 0 bytes. */
typedef struct ycge_peer {
    void *sa, *sb, *flags;                                   /* device pointers (valid inside the exporting process) */
    unsigned char sa_ipc[64], sb_ipc[64], flags_ipc[64];     /* cudaIpcMemHandle_t of the same allocations */
} ycge_peer;
YCGE_API int ycge_peer_export(ycge_ctx *ctx, ycge_peer *out);
YCGE_API int ycge_peer_attach(ycge_ctx *ctx, const ycge_peer *above, const ycge_peer *below, int32_t via_ipc); /* NULL: no neighbour */
/* Frame pipelining across ranks (asynchronous path): instead of ycge_frame_finish, park the finished tile (its denoised
 * rows and exposure samples) in a slot; the all-reduce of the slot's samples, ycge_frame_finish_stashed (ordered exposure
 * sum + cells, enqueued on the GIVEN stream) and the gather run later on a side stream, in frame order, while the ctx's
 * stream already renders the next frames.  With the peer hand-off this turns the serial wavefront of the in-place pass
 * into a pipeline: rank g renders frame f while rank g-1 renders frame f+1 (sharding.py: render_pipelined). */
YCGE_API int ycge_stash_config(ycge_ctx *ctx, int32_t n_slots);
YCGE_API int ycge_frame_stash(ycge_ctx *ctx, int32_t slot);
YCGE_API int ycge_frame_finish_stashed(ycge_ctx *ctx, int32_t slot, void *cuda_stream);
YCGE_API int ycge_stash_logs_ptr(ycge_ctx *ctx, int32_t slot, void **ptr, size_t *bytes);
YCGE_API int ycge_frame_begin(ycge_ctx *ctx);
YCGE_API int ycge_frame_halo(ycge_ctx *ctx, ycge_halo *out);  /* 1: an in-place pass is pending, 0: none (negative: error) */
YCGE_API int ycge_frame_inplace(ycge_ctx *ctx);
YCGE_API int ycge_frame_finish(ycge_ctx *ctx);
/* Frame-parallel sharding (asynchronous path, sharding.FrameParallelRenderer).  Row tiles cannot shorten the wavefront
 * of the in-place à-trous pass (every row depends on the rows above), but nothing of a frame's à-trous passes feeds the
 * next frame (see ycge_pipeline_config).  So only the FRONT of a frame (trace + TAA, RaytraceRenderer.cs:181-216) is cut
 * into row tiles -- ycge_frame_front on a row-tile ctx: the tile plus ONE halo row for the 3x3 luma clamp, no à-trous
 * halo -- and the BACK (à-trous passes, exposure samples) and the FINISH (ordered exposure sum, cells) of whole frames go
 * round-robin over the ranks: frame f's tiles of history + guides (YCGE_PTR_HIST/GND/GAS rows of every rank) are sent into
 * a back slot of rank f mod N (ycge_back_ptr), which runs ycge_back_denoise and, once the exposure state of frame f-1 has
 * arrived in YCGE_PTR_EXPOSURE, ycge_back_finish, then passes the state on.  Bit-identical to the unsharded frame. */
YCGE_API int ycge_frame_front(ycge_ctx *ctx);
YCGE_API int ycge_back_config(ycge_ctx *ctx, int32_t n_slots);                       /* whole-frame ctx only */
YCGE_API int ycge_back_ptr(ycge_ctx *ctx, int32_t slot, int32_t kind, void **ptr, size_t *bytes);
YCGE_API int ycge_back_denoise(ycge_ctx *ctx, int32_t slot, void *cuda_stream);
YCGE_API int ycge_back_finish(ycge_ctx *ctx, int32_t slot, void *cuda_stream);
/* ANSITerminalRenderer.Render's byte stream (ANSITerminalRenderer.cs:86-153: cursor address per row, colour escapes only
 * where the 8-bit indices change, UTF-8 glyphs, final reset) for the cells of the last frame, produced on the device and
 * copied to `out` (at most `cap` bytes; *n_bytes receives the length).  The resize prologue ("ESC[2J ESC[H", :103-106)
 * is the host's.  For a row-tile ctx the stream covers the tile's rows and starts with unknown colour state. */
YCGE_API int ycge_ansi_emit(ycge_ctx *ctx, uint8_t *out, size_t cap, size_t *n_bytes);
typedef enum ycge_ptr_kind {
    YCGE_PTR_CELLS = 0, YCGE_PTR_LOG_SAMPLES = 1,
    YCGE_PTR_HIST = 2,      /* taaHistory of the last frame: float4 rgb + luma, full-frame layout x + y*hiW */
    YCGE_PTR_GND = 3,       /* guides of the last frame: normalised normal + depth */
    YCGE_PTR_GAS = 4,       /* guides of the last frame: albedo + sky flag */
    YCGE_PTR_EXPOSURE = 5   /* ToneMapper state (16 bytes: aeExposure, effective exposure, logSum, count) */
} ycge_ptr_kind;
YCGE_API int ycge_device_ptr(ycge_ctx *ctx, int32_t kind, void **ptr, size_t *bytes);
YCGE_API int ycge_set_stream(ycge_ctx *ctx, void *cuda_stream); /* run on the caller's stream (e.g. torch's current stream) */

/* ---- introspection ---------------------------------------------------------------------- */
YCGE_API int ycge_debug_read(ycge_ctx *ctx, int32_t kind, void *dst, size_t bytes);
YCGE_API int ycge_get_stats(ycge_ctx *ctx, ycge_stats *out);
YCGE_API int ycge_get_frame_counter(ycge_ctx *ctx, int64_t *frame);
/* RaytraceSampler.cs:36-80 and Rng.cs:3-29, evaluated on the device for known-answer tests:
 * out[i] = first `n_draws` NextUnit() bit patterns of stream i. which: 0 = RaytraceSampler.Rng seeded by
 * PerFrameSeed(x[i],y[i],frame[i]); 1 = ConsoleRayTracing.Rng(seed = ((u64)x[i]<<32)|(u32)y[i]). */
YCGE_API int ycge_rng_kat(ycge_ctx *ctx, int32_t which, int32_t n, const int32_t *x, const int32_t *y,
                          const int64_t *frame, int32_t n_draws, uint32_t *out_bits, uint64_t *out_seed);

#ifdef __cplusplus
}
#endif
#endif /* YCGE_H */
