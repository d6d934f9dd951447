// device_types.h — HBM data layout shared by the host API (ycge_lib.cu) and the kernels.
// See DESIGN.md "Data layout in HBM".
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ycge {

// ---- acceleration structures -------------------------------------------------------------------------
// The reference keeps a 40-byte SoA node (own box + left/right/start/count, BVH.cs:11-20) and, per traversal
// step, pops a node, re-tests its own box, then fetches both children's boxes (3 dependent node fetches).
// Device layout: one 64-byte "pair node" per INTERNAL node holding both children's boxes and references, so
// one 4x16-byte fetch decides both children; leaves are folded into the reference that points at them.
//   q0 = (Lmin.x, Lmin.y, Lmin.z, Lmax.x)  q1 = (Lmax.y, Lmax.z, Rmin.x, Rmin.y)
//   q2 = (Rmin.z, Rmax.x, Rmax.y, Rmax.z)  q3 = (Lref, Rref, 0, 0) as int bits
// Child reference: >= 0 internal pair-node index; YCGE_REF_NONE = missing child; other negative values are leaves:
//   v = ~ref, count = (v >> 26) + 1 (1..32), start = v & 0x03FFFFFF into the leaf-ordered primitive array.
struct PairNode { float4 q0, q1, q2, q3; };
static constexpr int YCGE_REF_NONE = (int)0x80000000;
static constexpr int YCGE_LEAF_MAX_COUNT = 32;
static constexpr int YCGE_LEAF_MAX_START = 0x03FFFFFF;

struct TreeRoot { // the root's own box is tested once per query (BVH.cs:124-133 on the first pop)
    float lo[3], hi[3];
    int ref; // root reference (a leaf reference when the whole tree is one leaf)
    int pad;
};

// Triangles in leaf order, 48 bytes: t0 = (A.x, A.y, A.z, e1.x) t1 = (e1.y, e1.z, e2.x, e2.y) t2 = (e2.z, n.x, n.y, n.z)
struct DevTri { float4 t0, t1, t2; };

struct DevMesh {
    const PairNode *nodes;
    const DevTri *tris;       // leaf order
    const int *tri_id;        // leaf slot -> original triangle index (MeshLoader face order) = subId
    TreeRoot root;
    int material;             // index into the scene material table
    int n_tris;
};

// Voxel grid: one byte per voxel in the reference's bricked-Morton order (VolumeGrid.cs:235-252):
// 0 = empty (matId <= 0), else 1 + material index, i.e. materialLookup(mat, meta) resolved at upload.
struct DevVolume {
    const uint8_t *vox;
    const uint8_t *occ; // one byte per 8^3 brick: bit o = octant o (4^3 voxels, x>>2 | y>>2 << 1 | z>>2 << 2) holds a solid voxel
    int nx, ny, nz, nbx, nby, nbz;
    float min_corner[3];
    float voxel_size[3];
    int wireframe;
    float wire_width_frac;
    float wire_max_distance;
    int pad;
};

// Top-level object, 128 bytes = 8 x 16 B. Mirrors ycge_object plus the values the C# constructors derive.
struct DevObject {
    int kind, mat_a, mat_b, override_sr;                 // h0
    float checker_scale, specular, reflectivity; int ref; // h1  (ref: mesh / volume slot)
    float p[12];                                          // p0..p2
    float d[12];                                          // d0..d2: PLANE d0=ndotPoint; DISK d0=ndotCenter,d1=radius^2;
                                                          // CYLINDER d0=radius^2; TRIANGLE e1(0..2) e2(3..5) n(6..8)
};

struct DevLight { float pos[3]; float color[3]; float intensity; float pad; };

// Renderer/Texture.cs (static image): RGBA bytes, row-major, row 0 first (Texture.cs:81-90)
struct DevTexture { const uchar4 *px; int w, h, pad; };

struct DevScene {
    const PairNode *nodes;
    const int *leaf_obj;       // leaf slot -> object index (BVH.leafObjIndex order)
    const DevObject *objects;
    const float4 *materials;   // 4 x float4 per material (ycge_material layout)
    const DevMesh *meshes;
    const DevVolume *volumes;
    const DevLight *lights;
    const DevTexture *textures; // indexed by the slot stored in the material's 4th vector; n_textures == 0: no material is textured
    int n_textures;
    TreeRoot root;
    int n_lights;
    int is_volume_scene;
    float bg_top[3], bg_bottom[3];
    float ambient[3]; // colour * intensity is NOT pre-multiplied: the reference multiplies per hit (RaytraceRenderer.cs:573)
    float ambient_intensity;
};

// Per-frame constants derived on the host exactly as MakeJitteredRay (RaytraceRenderer.cs:419-437) derives them.
struct FrameConsts {
    float cam[3];
    float fwd[3], right[3], up[3];
    float half_w, half_h;
    float rot0, rot1;           // Frac((frameIdx+1)*0.7548776662466927f), ...0.5698402909980532f
    float jitter_rot_x, jitter_rot_y;
    long long frame;            // PerFrameSeed's frame argument
    int W, H;                   // hiW, hiH
    int y0, y1;                 // pixel rows [y0,y1) this launch covers
};

struct TraceParams {
    int diffuse_bounces, max_mirror_bounces, max_refractions;
    float mirror_threshold, eps, sigma_rad;
    unsigned long long seed_salt;
    int *host_err; // mapped host memory: a pixel whose traversal overflowed the device stack sets bit 1 (the frame call then fails)
};

struct TraceCounters { // device-side, zeroed at the start of every frame
    unsigned long long rays, top_nodes, mesh_nodes, leaf_refs, tris, prims, dda, stack_overflow;
    unsigned long long next_tile; // trace_stream_kernel: the next 8x4 pixel tile to hand out (low 32 bits)
};
struct TraceTotals { unsigned long long rays_total; }; // never reset: lets a benchmark difference it around a timed region

// Per-pixel image planes, all float4, row-major x + y*W over the FULL frame (every GPU allocates the full frame
// and only touches its tile + halo; see DESIGN.md "Multi-GPU").
struct ImagePlanes {
    float4 *cur;        // currentHdr.rgb, luma
    float4 *gnd[2];     // normalised gNormal.xyz, gDepth   (ping-pong: [frame&1] = this frame, other = prev*)
    float4 *gas[2];     // gAlbedo.rgb, sky flag (0/1)
    float4 *hist;       // taaHistory.rgb, luma
    float4 *sa, *sb;    // à-trous ping-pong: rgb, luma
    int2 *prim;         // primary objId, subId
    float *rays;        // optional debug: 6 floats per pixel (nullptr unless debug taps are enabled)
};

} // namespace ycge
