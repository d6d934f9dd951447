// ycge_host.hpp — C++ mirror of the reference's C# host side of the ray tracing path.
//
// The reference host is C# (.NET 8) and no C# toolchain exists in this image, so the classes a C# maintainer would
// keep (scene objects, the BuildSceneTable() scene factories, MeshLoader, Framebuffer/Chexel, the IConsoleRenderer
// seam, ANSITerminalRenderer's byte-stream emission) are mirrored here with the same names, argument meaning and
// error behaviour, sitting ABOVE the C ABI of include/ycge.h exactly as host_cs/CudaRaytraceRenderer.cs does.
// Nothing here renders: frames come from libycge.so (CUDA) only.
// File references are relative to /root/reference/ConsoleGame/.
#pragma once
#include "../../include/ycge.h"
#include "../csrc/bvh_build_parallel.hpp"

#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ycge_host {

struct Vec3 { // RayTracing/Vec3.cs
    float X = 0, Y = 0, Z = 0;
    Vec3() {}
    Vec3(float x, float y, float z) : X(x), Y(y), Z(z) {}
    Vec3(double x, double y, double z) : X((float)x), Y((float)y), Z((float)z) {}
    Vec3 operator+(const Vec3 &b) const { return Vec3(X + b.X, Y + b.Y, Z + b.Z); }
    Vec3 operator-(const Vec3 &b) const { return Vec3(X - b.X, Y - b.Y, Z - b.Z); }
    Vec3 operator*(float s) const { return Vec3(X * s, Y * s, Z * s); }
    float Dot(const Vec3 &b) const { return X * b.X + Y * b.Y + Z * b.Z; }
    Vec3 Normalized() const;
};

struct Material { // RayTracing/Material.cs:5-61
    Vec3 Albedo;
    double Specular = 0, Reflectivity = 0;
    Vec3 Emission;
    double Transparency = 0, IndexOfRefraction = 1.5;
    Vec3 TransmissionColor = Vec3(1.0, 1.0, 1.0);
    int DiffuseTexture = -1;
    double TextureWeight = 1.0, UVScale = 1.0;
    Material() {}
    Material(Vec3 albedo, double specular, double reflectivity, Vec3 emission) : Albedo(albedo), Specular(specular), Reflectivity(reflectivity), Emission(emission) {}
    Material(Vec3 albedo, double specular, double reflectivity, Vec3 emission, double transparency, double ior, Vec3 tint)
        : Albedo(albedo), Specular(specular), Reflectivity(reflectivity), Emission(emission), Transparency(transparency), IndexOfRefraction(ior), TransmissionColor(tint) {}
    ycge_material ToAbi() const;
};

// Renderer/Texture.cs, static images only: int[] pixels = RGBA bytes (byte 0 = R), row-major, row 0 first (:81-90).
// Sampling (SampleBilinear :143-162) runs on the GPU; Material.DiffuseTexture is the texture's index in Scene::Textures.
class Texture {
  public:
    int width = 0, height = 0;
    std::vector<uint32_t> pixels;
    Texture() {}
    Texture(int w, int h, const uint32_t *rgba) : width(w), height(h), pixels(rgba, rgba + (size_t)w * h) {}
    explicit Texture(const std::string &pngPath);       // Cv2.ImRead(Color) + BGR2RGBA (:25-49): 8-bit non-interlaced PNG, alpha := 255
    static Texture Procedural(int w, int h);            // deterministic stand-in when the asset is missing (tests, labelled in Scene::Name)
};

// Func<Vec3, Vec3, float, Material> as data: the closed set used by the reference (Scenes/Scenes.cs:408-428)
struct MaterialFunc {
    Material a, b;
    float scale = 0.0f; // 0: constant `a`
};
MaterialFunc Solid(Vec3 albedo);
MaterialFunc Emissive(Vec3 emission);
MaterialFunc Checker(Vec3 a, Vec3 b, float scale);
MaterialFunc Constant(const Material &m);

struct PointLight { Vec3 Position, Color; float Intensity; PointLight(Vec3 p, Vec3 c, float i) : Position(p), Color(c), Intensity(i) {} };
struct AmbientLight { Vec3 Color = Vec3(1.0, 1.0, 1.0); float Intensity = 0.075f; AmbientLight() {} AmbientLight(Vec3 c, float i) : Color(c), Intensity(i) {} };

class SceneExport;

class Hittable { // RayTracing/Objects/Hittable.cs (Hit itself runs on the GPU)
  public:
    virtual ~Hittable() {}
    virtual bool TryGetBounds(float &minX, float &minY, float &minZ, float &maxX, float &maxY, float &maxZ, float &cx, float &cy, float &cz) const = 0;
    virtual void Export(SceneExport &out) const = 0;
};

class Sphere : public Hittable { // BoundedObjects.cs:6-70
  public:
    Vec3 Center; float Radius; Material Mat;
    Sphere(Vec3 c, float r, Material m) : Center(c), Radius(r), Mat(m) {}
    bool TryGetBounds(float &, float &, float &, float &, float &, float &, float &, float &, float &) const override;
    void Export(SceneExport &out) const override;
};
class Plane : public Hittable { // Surfaces.cs:7-72
  public:
    Vec3 Point, Normal; MaterialFunc MatFunc; float Specular, Reflectivity;
    Plane(Vec3 p, Vec3 n, MaterialFunc f, float specular, float reflectivity) : Point(p), Normal(n.Normalized()), MatFunc(f), Specular(specular), Reflectivity(reflectivity) {}
    bool TryGetBounds(float &, float &, float &, float &, float &, float &, float &, float &, float &) const override;
    void Export(SceneExport &out) const override;
};
class Disk : public Hittable { // Surfaces.cs:73-143
  public:
    Vec3 Center, Normal; float Radius; MaterialFunc MatFunc; float Specular, Reflectivity;
    Disk(Vec3 c, Vec3 n, float r, MaterialFunc f, float specular, float reflectivity) : Center(c), Normal(n.Normalized()), Radius(r), MatFunc(f), Specular(specular), Reflectivity(reflectivity) {}
    bool TryGetBounds(float &, float &, float &, float &, float &, float &, float &, float &, float &) const override;
    void Export(SceneExport &out) const override;
};
class AxisRect : public Hittable { // XYRect / XZRect / YZRect, Surfaces.cs:144-358
  public:
    int Kind; // YCGE_XYRECT / YCGE_XZRECT / YCGE_YZRECT
    float A0, A1, B0, B1, K; MaterialFunc MatFunc; float Specular, Reflectivity;
    AxisRect(int kind, float a0, float a1, float b0, float b1, float k, MaterialFunc f, float specular, float reflectivity)
        : Kind(kind), A0(a0), A1(a1), B0(b0), B1(b1), K(k), MatFunc(f), Specular(specular), Reflectivity(reflectivity) {}
    bool TryGetBounds(float &, float &, float &, float &, float &, float &, float &, float &, float &) const override;
    void Export(SceneExport &out) const override;
};
inline std::shared_ptr<Hittable> XYRect(float x0, float x1, float y0, float y1, float z, MaterialFunc f, float s, float r) { return std::make_shared<AxisRect>(YCGE_XYRECT, x0, x1, y0, y1, z, f, s, r); }
inline std::shared_ptr<Hittable> XZRect(float x0, float x1, float z0, float z1, float y, MaterialFunc f, float s, float r) { return std::make_shared<AxisRect>(YCGE_XZRECT, x0, x1, z0, z1, y, f, s, r); }
inline std::shared_ptr<Hittable> YZRect(float y0, float y1, float z0, float z1, float x, MaterialFunc f, float s, float r) { return std::make_shared<AxisRect>(YCGE_YZRECT, y0, y1, z0, z1, x, f, s, r); }
class Box : public Hittable { // BoundedObjects.cs:72-116
  public:
    Vec3 Min, Max; MaterialFunc MatFunc; float Specular, Reflectivity;
    Box(Vec3 mn, Vec3 mx, MaterialFunc f, float specular, float reflectivity) : Min(mn), Max(mx), MatFunc(f), Specular(specular), Reflectivity(reflectivity) {}
    bool TryGetBounds(float &, float &, float &, float &, float &, float &, float &, float &, float &) const override;
    void Export(SceneExport &out) const override;
};
class CylinderY : public Hittable { // BoundedObjects.cs:118-248
  public:
    Vec3 Center; float Radius, YMin, YMax; bool Capped; Material Mat;
    CylinderY(Vec3 c, float r, float yMin, float yMax, bool capped, Material m);
    bool TryGetBounds(float &, float &, float &, float &, float &, float &, float &, float &, float &) const override;
    void Export(SceneExport &out) const override;
};
class Triangle : public Hittable { // Triangle.cs
  public:
    Vec3 A, B, C; Material Mat;
    Triangle(Vec3 a, Vec3 b, Vec3 c, Material m) : A(a), B(b), C(c), Mat(m) {}
    bool TryGetBounds(float &, float &, float &, float &, float &, float &, float &, float &, float &) const override;
    void Export(SceneExport &out) const override;
};

// MeshBVH (MeshBVH.cs): triangle SoA + SAH tree built on the host, exported flat.
class MeshBVH {
  public:
    static int counter; // MeshBVH.cs:13 (counts triangles loaded)
    std::vector<float> ax, ay, az, e1x, e1y, e1z, e2x, e2y, e2z, nx, ny, nz;
    std::vector<float> abc; // the triangles' A,B,C as given (9 floats each), for the ycge_mesh_upload_triangles form
    ycge::FlatTree tree;
    Material triMat;
    explicit MeshBVH(const std::vector<Triangle> &tris);
    int TriangleCount() const { return (int)ax.size(); }
};
class Mesh : public Hittable { // Mesh.cs
  public:
    Vec3 BoundsMin, BoundsMax;
    std::shared_ptr<MeshBVH> bvh;
    Mesh(const std::vector<Triangle> &triangles, Vec3 mn, Vec3 mx) : BoundsMin(mn), BoundsMax(mx), bvh(std::make_shared<MeshBVH>(triangles)) {}
    bool TryGetBounds(float &, float &, float &, float &, float &, float &, float &, float &, float &) const override;
    void Export(SceneExport &out) const override;
};

struct ObjData { std::vector<Vec3> positions; std::vector<int> faces; /* 3 per triangle */ };
class MeshLoader { // MeshLoader.cs
  public:
    static ObjData ParseObj(const std::string &path); // throws std::runtime_error("OBJ not found") like FileNotFoundException
    static std::shared_ptr<Mesh> FromObj(const std::string &path, Material defaultMaterial, float scale = 1.0f, Vec3 translate = Vec3(), bool normalize = true, float targetSize = 1.0f);
    static std::shared_ptr<Mesh> FromData(ObjData data, Material defaultMaterial, float scale = 1.0f, Vec3 translate = Vec3(), bool normalize = true, float targetSize = 1.0f);
    static void NormalizeAllUsedVertices(std::vector<Vec3> &pos, const std::vector<int> &faces, float targetSize);
    static bool BoundsNormalizedLargestComponent(const ObjData &d, Vec3 &mn, Vec3 &mx); // MeshScenes.TryReadObjBoundsNormalized :186-331
};

// VolumeGrid (VolumeGrid.cs): bricked-Morton int arrays + palette table.
struct VoxelPalette { int n_ids = 0, meta_levels = 1, def = 0; std::vector<Material> materials; std::vector<int> table; /* n_ids*levels -> materials[] */ };
class VolumeGrid : public Hittable {
  public:
    int nx, ny, nz, nbx, nby, nbz;
    std::vector<int32_t> mat, meta;
    Vec3 minCorner, voxelSize;
    std::shared_ptr<VoxelPalette> palette;
    bool wireframe; float wireWidthFrac, wireMaxDistance;
    // cells(ix,iy,iz) -> (matId, metaId); mirrors `new VolumeGrid((int,int)[,,] cells, ...)` VolumeGrid.cs:55-93
    VolumeGrid(int nx, int ny, int nz, const std::function<void(int, int, int, int &, int &)> &cells, Vec3 minCorner, Vec3 voxelSize,
               std::shared_ptr<VoxelPalette> materialLookup, bool enableWireframe = true, float wireWidthFraction = 0.06f, float wireMaxDistance = 16.0f);
    int IndexOf(int ix, int iy, int iz) const;
    static int Morton3_3bits(int x, int y, int z);
    bool AnySolid() const;
    bool TryGetBounds(float &, float &, float &, float &, float &, float &, float &, float &, float &) const override;
    void Export(SceneExport &out) const override;
};

class BVH { // Objects/BVH.cs: top-level tree over Scene.Objects
  public:
    ycge::FlatTree tree;
    explicit BVH(const std::vector<std::shared_ptr<Hittable>> &objects); // throws std::runtime_error("Unbounded Hittable") :39
};

class Scene;
// Scene entities (Scenes/Scene.cs:528-533 ISceneEntity): the ones that change the path's inputs between frames.  GetHittables and
// the entity -> Objects sync (:518-534) are folded into Scene::Objects, which keeps the same order.
class ISceneEntity {
  public:
    bool Enabled = true;
    virtual ~ISceneEntity() {}
    virtual void Update(float dt, Scene &scene) = 0; // dt in seconds
};
// Scenes/DayNightCycle.cs: sun / moon PointLights and the sky gradient of the voxel worlds.
class DayNightEntity : public ISceneEntity {
  public:
    explicit DayNightEntity(float cycleSeconds = 120.0f, float sunRadius = 2000.0f) : cycleSeconds(std::max(1.0f, cycleSeconds)), sunRadius(sunRadius) {}
    void Update(float dt, Scene &scene) override; // DayNightCycle.cs:41-91
    float Time() const { return time; }
  private:
    float time = 0.0f, cycleSeconds, sunRadius;
    int sun = -1, moon = -1; // indices into Scene::Lights (the reference keeps the PointLight objects)
};
// Scenes/TestScenesRandom.cs:688-798, the animated exhibits' entities.  (UVWobbleEntity :800-828 is not mirrored: Material is a
// struct, the entity wobbles its own copy and the scene never sees it.)
class BobbingSphereEntity : public ISceneEntity { // :688-720: moves the sphere and requests a geometry rebuild every update
  public:
    BobbingSphereEntity(std::shared_ptr<Sphere> sphere, float amplitude, float speed, float phase)
        : sphere(sphere), baseY((float)sphere->Center.Y), amplitude(amplitude), speed(speed), phase(phase) {}
    void Update(float dt, Scene &scene) override;
  private:
    std::shared_ptr<Sphere> sphere; float baseY, amplitude, speed, phase, t = 0.0f;
};
class OrbitingLightEntity : public ISceneEntity { // :722-757
  public:
    OrbitingLightEntity(int light, Vec3 pivot, float radius, float height, float speed, float phase) : light(light), pivot(pivot), radius(radius), height(height), speed(speed), phase(phase) {}
    void Update(float dt, Scene &scene) override;
  private:
    int light; Vec3 pivot; float radius, height, speed, phase, angle = 0.0f;
};
class PulsingLightEntity : public ISceneEntity { // :759-798
  public:
    PulsingLightEntity(const Scene &scene, int light, float baseScale, float ampFraction, float speed);
    void Update(float dt, Scene &scene) override;
  private:
    int light; float initialIntensity, minMult, maxMult, speed, t = 0.0f;
};

class Scene { // Scenes/Scene.cs
  public:
    std::vector<std::shared_ptr<Hittable>> Objects;
    std::vector<PointLight> Lights;
    std::vector<std::shared_ptr<Texture>> Textures; // Material.DiffuseTexture indexes this list
    int AddTexture(std::shared_ptr<Texture> t) { Textures.push_back(t); return (int)Textures.size() - 1; }
    Vec3 BackgroundTop = Vec3(0.6, 0.8, 1.0), BackgroundBottom = Vec3(1.0, 1.0, 1.0);
    AmbientLight Ambient;
    float DefaultFovDeg = 45.0f;
    Vec3 DefaultCameraPos = Vec3(0.0, 1.0, 0.0);
    float DefaultYaw = 0.0f, DefaultPitch = 0.0f;
    Vec3 CameraPos = Vec3(0.0, 1.0, 0.0);
    float Yaw = 0.0f, Pitch = 0.0f;
    bool HasDynamicTextures = false;
    bool IsVolumeScene = false; // `scene is VolumeScene`
    std::string Name;
    std::shared_ptr<BVH> bvh;
    void Add(std::shared_ptr<Hittable> h) { Objects.push_back(h); }
    void RebuildBVH() { bvh = std::make_shared<BVH>(Objects); } // Scene.cs:66-69
    void ResetCamera() { CameraPos = DefaultCameraPos; Yaw = DefaultYaw; Pitch = DefaultPitch; }
    std::vector<std::shared_ptr<ISceneEntity>> Entities;
    bool GeometryDirty = false;
    unsigned LightsVersion = 0;   // bumped when entities ran (they may have rewritten Lights / Background*): CudaRaytraceRenderer::SyncLights
    unsigned GeometryVersion = 0; // bumped when the tree was rebuilt: CudaRaytraceRenderer::SyncGeometry
    void AddEntity(std::shared_ptr<ISceneEntity> e) { if (e) Entities.push_back(e); }
    void RequestGeometryRebuild() { GeometryDirty = true; } // Scene.cs:508-511
    void Update(float deltaTimeMS) { // Scene.cs:100-127: milliseconds in, entities get seconds; then the tree if geometry changed
        float dt = deltaTimeMS * 0.001f;
        if (dt < 0.0f) dt = 0.0f;
        for (auto &e : Entities) if (e && e->Enabled) e->Update(dt, *this);
        if (!Entities.empty()) LightsVersion++;
        if (GeometryDirty || !bvh) { GeometryDirty = false; RebuildBVH(); GeometryVersion++; }
    }
};

// Flat arrays handed to the C ABI (what host_cs/CudaRaytraceRenderer.cs marshals)
class SceneExport {
  public:
    std::vector<ycge_material> materials;
    std::vector<ycge_object> objects;
    std::vector<ycge_light> lights;
    std::vector<std::shared_ptr<MeshBVH>> meshes;      // id = index
    std::vector<const VolumeGrid *> volumes;           // id = index
    std::vector<std::vector<int32_t>> volume_palettes; // per volume, indices into `materials`
    int AddMaterial(const Material &m);
    void AddFunc(ycge_object &o, const MaterialFunc &f, float specular, float reflectivity);
};

// ---- SceneSyncProtocol (Scenes/SyncScene.cs:267-569): the engine's 'SCNE' v1 scene snapshot, used here as a scene
// interchange format for the harness.  Quirks kept: material functions are baked by ONE call at a fixed point (a checker
// floor becomes the colour under that point), a Box is always written as the writer's grey stand-in material (:350-359),
// textures are not serialised (:541), meshes and volume grids are skipped (:384-387).
namespace SceneSyncProtocol {
std::vector<uint8_t> WriteSnapshot(const Scene &scene);
std::shared_ptr<Scene> ReadSnapshot(const std::vector<uint8_t> &bytes);
} // namespace SceneSyncProtocol

// ---- scene factories (BuildSceneTable(), RaytraceEntity.cs:319-344) --------------------------------------------
namespace Scenes { // Scenes/Scenes.cs
std::shared_ptr<Scene> BuildTestScene();
std::shared_ptr<Scene> BuildCornellBox();
std::shared_ptr<Scene> BuildMirrorSpheresOnChecker();
std::shared_ptr<Scene> BuildCylindersDisksAndTriangles();
std::shared_ptr<Scene> BuildBoxesShowcase();
std::shared_ptr<Scene> BuildVolumeGridTestScene();
std::shared_ptr<Scene> BuildTextureTestScene(); // assets/image.png if present, else a procedural stand-in (labelled in Scene::Name)
std::shared_ptr<Scene> BuildEntitiesDemo();     // NOT in the reference: bobbing spheres, an orbiting and a pulsing light (tests)
std::shared_ptr<Scene> BuildTextureGallery();   // NOT in the reference: every primitive kind that carries U,V, textured (tests)
} // namespace Scenes
namespace TestScenes { // Scenes/TestScenes.cs
std::shared_ptr<Scene> BuildTestScene(); // the "museum": three Cornell rooms, a mesh gallery, pedestals, textures, two voxel dioramas
} // namespace TestScenes
namespace MeshScenes { // Scenes/MeshScenes.cs
extern std::string AssetDir; // where cow.obj / stanford-bunny.obj / teapot.obj / xyzrgb_dragon.obj are looked up
std::shared_ptr<Scene> BuildCowScene();
std::shared_ptr<Scene> BuildBunnyScene();
std::shared_ptr<Scene> BuildTeapotScene();
std::shared_ptr<Scene> BuildDragonScene(); // real xyzrgb_dragon.obj if present, else the procedural stand-in (labelled in Scene::Name)
std::shared_ptr<Scene> BuildAllMeshesScene(int knotU = 0, int knotV = 0); // cow, bunny, teapot, dragon (or its stand-in: bunny x4; a knot when knotU > 0) in one scene
ObjData SubdivideMidpoint(const ObjData &in);
ObjData DragonStandin(); // stanford-bunny with one level of midpoint subdivision, 277 804 triangles
std::shared_ptr<Scene> BuildMeshScene(const ObjData &mesh, Material mat, const std::string &name, Vec3 targetPos = Vec3(0.0f, 0.5f, 1.0f));
ObjData ProceduralKnot(int segU, int segV); // procedural mesh of a chosen size (tests): 2*segU*segV triangles
} // namespace MeshScenes
namespace VolumeScenes { // structure of Scenes/VolumeScenes.cs:569-627 over a synthetic heightfield (generator is out of scope)
std::shared_ptr<Scene> BuildSyntheticWorld(int worldSize, int worldHeight, int chunkSize, float daySeconds);
std::shared_ptr<Scene> BuildIslandWorld(int worldSize, int worldHeight, int chunkSize, float daySeconds); // the reference's generator, seed 0
void WriteIslandWorldFile(const std::string &path, int worldSize, int worldHeight);
float SyntheticHeight(int x, int z, int worldHeight);
} // namespace VolumeScenes
std::shared_ptr<Scene> BuildSceneByName(const std::string &name);

// ---- output side ----------------------------------------------------------------------------------------------
struct ChexelColor { int color_16 = 0; Vec3 color_f32; int ansi_256 = 16; }; // Renderer/Chexel.cs:6-97 (+ cached ANSI index)
struct Chexel { char16_t Char = u' '; ChexelColor ForegroundColor, BackgroundColor; }; // Chexel.cs:99-125
class Framebuffer { // Renderer/Framebuffer.cs
  public:
    int Width, Height;
    Framebuffer(int width, int height);
    Chexel GetChexel(int x, int y) const { return chexels[(size_t)x * Height + y]; }          // Chexel[x,y]: x is the slow index
    void SetChexel(int x, int y, const Chexel &c) { chexels[(size_t)x * Height + y] = c; }
  private:
    std::vector<Chexel> chexels;
};

class IConsoleRenderer { // RaytraceEntity.cs:12-18
  public:
    virtual ~IConsoleRenderer() {}
    virtual void SetCamera(Vec3 pos, float yaw, float pitch) = 0;
    virtual void SetFov(float fovDeg) = 0;
    virtual void TryFlipAndBlit(Framebuffer &fb) = 0;
    virtual void Resize(Framebuffer &fb, int superSample) = 0;
};

// The drop-in: same constructor arguments as RaytraceRenderer (RaytraceRenderer.cs:74), frames from libycge.so.
class CudaRaytraceRenderer : public IConsoleRenderer {
  public:
    CudaRaytraceRenderer(Framebuffer &framebuffer, Scene &scene, float fovDeg, int pxW, int pxH, int superSample, int device = 0, int tileRow0 = 0, int tileRows = 0,
                         const std::vector<int> &devices = std::vector<int>()); // two or more devices: the library renders frames in parallel over them (ycge_config.n_devices)
    ~CudaRaytraceRenderer() override;
    void SetCamera(Vec3 pos, float yaw, float pitch) override;
    void SetFov(float fovDeg) override;
    void TryFlipAndBlit(Framebuffer &fb) override;
    void Resize(Framebuffer &fb, int superSample) override;
    void UploadTexture(int id, const Texture &t);   // new Texture(path) -> ycge_texture_upload
    void UploadScene(Scene &scene);                 // scene switch (RaytraceEntity.SwitchToScene :234-246)
    void SyncLights(const Scene &scene);            // after scene.Update(ms): ycge_lights_update + ycge_globals_update when an entity moved them
    void SyncGeometry(Scene &scene);                // after scene.Update(ms) rebuilt the tree: objects + top-level BVH again, history kept
    void RenderCells(ycge_cell *out);               // TryFlipAndBlit without the Chexel unpack (headless)
    ycge_ctx *Context() { return ctx; }
    int fbW, fbH, ss;
  private:
    ycge_ctx *ctx = nullptr;
    std::vector<ycge_cell> staging;
    void Check(int rc, const char *what);
};

class ANSITerminalRenderer { // Renderer/ANSITerminalRenderer.cs:86-153 (byte stream only; no console I/O)
  public:
    static std::vector<uint8_t> Render(const Framebuffer &fb, int consoleWidth, int consoleHeight);
    static std::vector<uint8_t> RenderCells(const ycge_cell *cells, int fbW, int fbH);
};
class Win32TerminalRenderer { // Renderer/Win32TerminalRenderer.cs:68-112: CHAR_INFO {char, attr}
  public:
    static std::vector<uint32_t> BuildCharInfo(const Framebuffer &fb);
};

} // namespace ycge_host
