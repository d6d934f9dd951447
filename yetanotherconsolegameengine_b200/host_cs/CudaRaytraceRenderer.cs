// CudaRaytraceRenderer.cs — the reference-side binding of libycge (include/ycge.h).
//
// This is the one file a maintainer of YetAnotherConsoleGameEngine adds to ConsoleGame/RayTracing/ to make the
// B200 library the frame producer behind RaytraceEntity.IConsoleRenderer (ConsoleGame/RaytraceEntity.cs:12-18).
// It follows the house conventions of the engine's existing P/Invoke code (Win32TerminalRenderer.cs:119-151):
// [DllImport] static extern, blittable [StructLayout(LayoutKind.Sequential)] structs, failure ->
// InvalidOperationException carrying the native message, IDisposable owner.
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no .NET toolchain.  Every struct below mirrors include/ycge.h
// field for field (tests/test_abi.py pins the C layout against the ctypes mirror; the same sizes are asserted in the
// static constructor here).  The C++ mirror of this host logic that IS compiled and tested is
// yetanotherconsolegameengine_b200/host/ycge_host.cpp (Flatten(), CudaRaytraceRenderer::TryFlipAndBlit).
//
// Host-side change footprint in the engine (INTEGRATION.md): this file, three `internal` accessors that expose the
// private flat arrays of BVH / MeshBVH / VolumeGrid (BVH.cs:11-25, MeshBVH.cs:18-39, VolumeGrid.cs:25-32), and the
// three construction sites in RaytraceEntity (RaytraceEntity.cs:97-98, 240-241, 262-263).
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;
using ConsoleGame.RayTracing.Objects;
using ConsoleGame.RayTracing.Scenes;
using ConsoleGame.Renderer;

namespace ConsoleGame.RayTracing
{
    public sealed class CudaRaytraceRenderer : IDisposable
    {
        private const string Lib = "ycge"; // libycge.so / ycge.dll

        // ---- include/ycge.h, mirrored ---------------------------------------------------------------------------
        [StructLayout(LayoutKind.Sequential)]
        public struct YMaterial
        {
            public float AlbedoX, AlbedoY, AlbedoZ, Reflectivity;
            public float EmissionX, EmissionY, EmissionZ, Transparency;
            public float TransmissionX, TransmissionY, TransmissionZ, Ior;
            public float Specular; public int TexId; public float TexWeight, UvScale;
        } // 64 bytes

        [StructLayout(LayoutKind.Sequential)]
        public unsafe struct YObject
        {
            public int Kind, MatA, MatB; public float CheckerScale; public int OverrideSr;
            public float Specular, Reflectivity; public int RefId;
            public fixed float P[12];
        } // 80 bytes

        [StructLayout(LayoutKind.Sequential)]
        public struct YLight { public float Px, Py, Pz, Cr, Cg, Cb, Intensity; }

        [StructLayout(LayoutKind.Sequential)]
        public struct YBvh
        {
            public int NNodes, Root, NLeafRefs;
            public IntPtr MinX, MinY, MinZ, MaxX, MaxY, MaxZ, Left, Right, Start, Count, LeafIndex;
        }

        [StructLayout(LayoutKind.Sequential)]
        public struct YScene
        {
            public float BgTopX, BgTopY, BgTopZ, BgBotX, BgBotY, BgBotZ, AmbR, AmbG, AmbB, AmbientIntensity;
            public int IsVolumeScene, NLights; public IntPtr Lights;
            public int NMaterials; public IntPtr Materials;
            public int NObjects; public IntPtr Objects;
            public IntPtr Bvh;
        }

        [StructLayout(LayoutKind.Sequential)]
        public struct YMeshSoa
        {
            public int NTris;
            public IntPtr Ax, Ay, Az, E1x, E1y, E1z, E2x, E2y, E2z, Nx, Ny, Nz;
            public YMaterial Material; public IntPtr Bvh;
        }

        [StructLayout(LayoutKind.Sequential)]
        public struct YVolume
        {
            public int Nx, Ny, Nz; public float MinX, MinY, MinZ, SizeX, SizeY, SizeZ;
            public IntPtr Mat, Meta; public int Wireframe; public float WireWidthFrac, WireMaxDistance;
            public int PaletteNIds, PaletteMetaLevels; public IntPtr Palette; public int PaletteDefault;
        }

        [StructLayout(LayoutKind.Sequential)]
        public struct YParams
        {
            public int DiffuseBounces, MaxMirrorBounces, MaxRefractions, AtrousIterations;
            public float MirrorThreshold, Eps, TaaAlpha, MotionTransReset, MotionRotReset, DiffuseSigmaDeg, LuminancePad;
            public float CPhi, NPhi, ZPhi, APhi, ToneExposure, ToneGamma, AeKey, AeSpeed, AeMin, AeMax, Saturation, Vibrance;
            public int AutoExposure; public ulong SeedSalt;
        }

        [StructLayout(LayoutKind.Sequential)]
        public unsafe struct YConfig
        {
            public int FbW, FbH, Ss, Device, TileRow0, TileRows; public YParams Params;
            public int NDevices; public fixed int Devices[8]; // >= 2: the library drives these GPUs of the box behind this one context
        }

        [StructLayout(LayoutKind.Sequential, Pack = 1)]
        public struct YCell
        {
            public ushort Glyph; public byte Fg16, Bg16, FgAnsi, BgAnsi; public ushort Attr;
            public float FgR, FgG, FgB, BgR, BgG, BgB;
        } // 32 bytes

        public enum Kind { Sphere = 0, Plane = 1, Disk = 2, XYRect = 3, XZRect = 4, YZRect = 5, Box = 6, CylinderY = 7, Triangle = 8, Mesh = 9, Volume = 10 }

        [DllImport(Lib)] private static extern void ycge_default_params(out YParams p);
        [DllImport(Lib)] private static extern int ycge_create(ref YConfig cfg, out IntPtr ctx);
        [DllImport(Lib)] private static extern void ycge_destroy(IntPtr ctx);
        [DllImport(Lib)] private static extern IntPtr ycge_last_error(IntPtr ctx);
        [DllImport(Lib)] private static extern int ycge_resize(IntPtr ctx, int fbW, int fbH, int ss);
        [DllImport(Lib)] private static extern int ycge_mesh_upload_soa(IntPtr ctx, int id, ref YMeshSoa mesh);
        // optional (SURVEY 8f-2): the same MeshBVH, node for node, built on the GPU from the raw triangle list (n x 9 floats: A, B, C);
        // a loader that knows this renderer is active may skip `new MeshBVH(tris)` and call this instead of ycge_mesh_upload_soa
        [DllImport(Lib)] private static extern int ycge_mesh_build_device(IntPtr ctx, int id, int nTris, [In] float[] abc, ref YMaterial material);
        [DllImport(Lib)] private static extern int ycge_volume_upload(IntPtr ctx, int id, ref YVolume vol);
        [DllImport(Lib)] private static extern int ycge_scene_upload(IntPtr ctx, ref YScene scene);
        [DllImport(Lib)] private static extern int ycge_lights_update(IntPtr ctx, int n, [In] YLight[] lights);
        [DllImport(Lib)] private static extern int ycge_globals_update(IntPtr ctx, [In] float[] bgTop, [In] float[] bgBottom, [In] float[] ambient, float ambientIntensity);
        [DllImport(Lib)] private static extern int ycge_set_camera(IntPtr ctx, [In] float[] pos, float yaw, float pitch);
        [DllImport(Lib)] private static extern int ycge_set_fov(IntPtr ctx, float fovDeg);
        [DllImport(Lib)] private static extern int ycge_reset_history(IntPtr ctx);
        [DllImport(Lib)] private static extern int ycge_render_frame(IntPtr ctx, IntPtr cells, int strideCells);
        [DllImport(Lib)] private static extern int ycge_texture_upload(IntPtr ctx, int id, int w, int h, [In] int[] rgba);
        [DllImport(Lib)] private static extern int ycge_pipeline_config(IntPtr ctx, int nSlots);
        [DllImport(Lib)] private static extern int ycge_submit_frame(IntPtr ctx, IntPtr cells, int strideCells, out long frameId);
        [DllImport(Lib)] private static extern int ycge_frame_wait(IntPtr ctx, long frameId);

        static CudaRaytraceRenderer()
        {
            if (Marshal.SizeOf<YMaterial>() != 64 || Marshal.SizeOf<YObject>() != 80 || Marshal.SizeOf<YCell>() != 32)
                throw new InvalidOperationException("ycge.h struct layout mismatch");
        }

        // ---- the renderer -----------------------------------------------------------------------------------------
        private IntPtr ctx;
        private readonly Scene scene;
        private YCell[] cells;
        private GCHandle cellsPin; // long-lived pinned buffer, as VolumeGrid pins its arrays (VolumeGrid.cs:70-73)
        private int fbW, fbH, ss;
        private readonly float[] camTmp = new float[3];
        private BVH uploadedBvh; // the tree object the device copy was made from

        /// Same arguments as RaytraceRenderer's ctor (RaytraceRenderer.cs:74): hiW = fbW*ss, hiH = fbH*2*ss.
        /// `devices`: two or more CUDA ordinals of one box -> the library renders frames in parallel over them behind this one renderer.
        public unsafe CudaRaytraceRenderer(Framebuffer framebuffer, Scene scene, float fovDeg, int pxW, int pxH, int superSample, int device = 0, int[] devices = null)
        {
            this.scene = scene ?? throw new ArgumentNullException(nameof(scene));
            ss = Math.Max(1, superSample); fbW = framebuffer.Width; fbH = framebuffer.Height;
            var cfg = new YConfig { FbW = fbW, FbH = fbH, Ss = ss, Device = device };
            if (devices != null && devices.Length >= 2)
            {
                if (devices.Length > 8) throw new ArgumentException("at most 8 devices");
                cfg.NDevices = devices.Length;
                for (int k = 0; k < devices.Length; k++) cfg.Devices[k] = devices[k];
            }
            ycge_default_params(out cfg.Params); // the reference's compile-time constants (RaytraceRenderer.cs:31-43,65)
            Check(ycge_create(ref cfg, out ctx));
            AllocCells();
            scene.RebuildBVH(); // RaytraceRenderer.cs:107
            UploadScene();
            Check(ycge_set_fov(ctx, fovDeg));
        }

        public void SetCamera(Vec3 pos, float yaw, float pitch) // RaytraceRenderer.cs:140-148
        {
            camTmp[0] = pos.X; camTmp[1] = pos.Y; camTmp[2] = pos.Z;
            Check(ycge_set_camera(ctx, camTmp, yaw, pitch));
        }

        public void SetFov(float fovDeg) { Check(ycge_set_fov(ctx, fovDeg)); } // :150-153

        public void Resize(Framebuffer fb, int superSample) // :110-138 (keeps frame counter and exposure)
        {
            ss = Math.Max(1, superSample); fbW = fb.Width; fbH = fb.Height;
            Check(ycge_resize(ctx, fbW, fbH, ss));
            AllocCells();
        }

        /// Synchronous, like the reference (:157-267): on return fb holds the finished frame.
        public void TryFlipAndBlit(Framebuffer fb)
        {
            if (scene.HasDynamicTextures) Check(ycge_reset_history(ctx)); // :171
            // The reference's renderer reads scene.Objects / scene.Lights afresh every frame.  Scene.Update (Scene.cs:100-127) may have
            // run entities since the last one: a rebuilt tree is a new BVH object (Scene.cs:66-69) -> flatten again, history kept;
            // otherwise lights and sky may still have moved (DayNightCycle.cs:80-89, Orbiting/PulsingLightEntity) -> two small calls.
            if (!ReferenceEquals(scene.Bvh, uploadedBvh)) UploadScene();
            else if (scene.Entities.Count > 0) UpdateLightsAndGlobals();
            Check(ycge_render_frame(ctx, cellsPin.AddrOfPinnedObject(), fbW));
            for (int cy = 0; cy < fbH; cy++)
            {
                for (int cx = 0; cx < fbW; cx++)
                {
                    ref YCell c = ref cells[cy * fbW + cx];
                    // new Chexel('▀', topSDR, botSDR) (:260): the ChexelColor(Vec3) ctor re-derives color_16 from color_f32
                    // (Chexel.cs:37-41); the library's Fg16/Bg16/FgAnsi/BgAnsi equal what the engine's renderers derive.
                    fb.SetChexel(cx, cy, new Chexel((char)c.Glyph, new Vec3(c.FgR, c.FgG, c.FgB), new Vec3(c.BgR, c.BgG, c.BgB)));
                }
            }
        }

        /// Frames in flight (include/ycge.h, ycge_pipeline_config): SubmitFrame enqueues SetCamera's pose and returns at once; the
        /// cells land in `dst` (pinned) when WaitFrame(id) returns.  A host loop that can show frame N-2 while frame N renders
        /// gets the GPU's throughput instead of one frame's latency; TryFlipAndBlit above stays the strict drop-in.
        public void ConfigurePipeline(int framesInFlight) { Check(ycge_pipeline_config(ctx, framesInFlight)); }
        public long SubmitFrame(IntPtr dst) { Check(ycge_submit_frame(ctx, dst, fbW, out long id)); return id; }
        public void WaitFrame(long id) { Check(ycge_frame_wait(ctx, id)); }

        /// Per-frame light / sky changes without re-uploading geometry (DayNightCycle.cs:80-89).
        public void UpdateLightsAndGlobals()
        {
            var l = new YLight[scene.Lights.Count];
            for (int i = 0; i < l.Length; i++) l[i] = ToLight(scene.Lights[i]);
            Check(ycge_lights_update(ctx, l.Length, l));
            Check(ycge_globals_update(ctx, V(scene.BackgroundTop), V(scene.BackgroundBottom), V(scene.Ambient.Color), scene.Ambient.Intensity));
        }

        // What survives a re-sync of the objects (the C++ mirror's SceneExport does the same, host/ycge_host.cpp): the material table
        // is append-only and de-duplicated by value -- a VolumeGrid's palette (37 materials, VoxelMaterialPalette.cs:48-98) and the
        // one or two materials of every object would otherwise pass the library's limit of 255 voxel materials, and an index, once
        // handed to ycge_volume_upload as part of a palette, must keep its meaning --; meshes, voxel grids and textures are uploaded
        // once per object and found again by reference.
        private readonly List<YMaterial> materials = new List<YMaterial>();
        private readonly Dictionary<(float, float, float, float, float, float, float, float, float, float, float, float, float, int, float, float), int> materialIndex
            = new Dictionary<(float, float, float, float, float, float, float, float, float, float, float, float, float, int, float, float), int>();
        private readonly Dictionary<ConsoleGame.Renderer.Texture, int> texIds = new Dictionary<ConsoleGame.Renderer.Texture, int>();
        private readonly Dictionary<Mesh, int> meshIds = new Dictionary<Mesh, int>();
        private readonly Dictionary<VolumeGrid, int> volIds = new Dictionary<VolumeGrid, int>();

        // new Texture(path) (Renderer/Texture.cs:25-49): the int[] pixels (RGBA bytes, row 0 first, :81-90) go to the device once per
        // distinct Texture object; Material.DiffuseTexture becomes its id.  Needs one internal accessor, Texture.Pixels.
        private int TexId(Material m)
        {
            if (m.DiffuseTexture == null) return -1;
            if (!texIds.TryGetValue(m.DiffuseTexture, out int id))
            {
                id = texIds.Count;
                Check(ycge_texture_upload(ctx, id, m.DiffuseTexture.width, m.DiffuseTexture.height, m.DiffuseTexture.Pixels));
                texIds[m.DiffuseTexture] = id;
            }
            return id;
        }
        private YMaterial ToMaterialWithTexture(Material m) { var ym = ToMaterial(m); ym.TexId = TexId(m); return ym; }
        private int AddMat(Material m)
        {
            YMaterial ym = ToMaterialWithTexture(m);
            var key = (ym.AlbedoX, ym.AlbedoY, ym.AlbedoZ, ym.Reflectivity, ym.EmissionX, ym.EmissionY, ym.EmissionZ, ym.Transparency,
                       ym.TransmissionX, ym.TransmissionY, ym.TransmissionZ, ym.Ior, ym.Specular, ym.TexId, ym.TexWeight, ym.UvScale);
            if (!materialIndex.TryGetValue(key, out int idx)) { idx = materials.Count; materials.Add(ym); materialIndex[key] = idx; }
            return idx;
        }

        /// Scene.RebuildBVH happened (geometry changed / scene switch, RaytraceEntity.cs:234-246; or an entity moved an object,
        /// Scene.cs:121-127): the object records and the top-level tree are flattened again -- the cheap part.  A mesh, a voxel grid
        /// or a texture the renderer has seen before is NOT uploaded again (a bobbing sphere rebuilds the tree every frame).
        public unsafe void UploadScene()
        {
            var objects = new List<YObject>();
            var pins = new List<GCHandle>();
            IntPtr Pin(Array a) { var h = GCHandle.Alloc(a, GCHandleType.Pinned); pins.Add(h); return h.AddrOfPinnedObject(); }
            try
            {
                foreach (Hittable h in scene.Objects) // enumeration order = primary-hit objId = BVH item index (BVH.cs:34-50)
                {
                    var o = new YObject { MatA = 0, MatB = 0, RefId = -1 };
                    switch (h)
                    {
                        case Sphere s: o.Kind = (int)Kind.Sphere; Set(o.P, s.Center.X, s.Center.Y, s.Center.Z, s.Radius); o.MatA = o.MatB = AddMat(s.Mat); break;
                        case Plane p: o.Kind = (int)Kind.Plane; Set(o.P, p.Point.X, p.Point.Y, p.Point.Z, p.Normal.X, p.Normal.Y, p.Normal.Z); MatFunc(ref o, p.MaterialFunc, p.Specular, p.Reflectivity, AddMat); break;
                        case Disk d: o.Kind = (int)Kind.Disk; Set(o.P, d.Center.X, d.Center.Y, d.Center.Z, d.Normal.X, d.Normal.Y, d.Normal.Z, d.Radius); MatFunc(ref o, d.MaterialFunc, d.Specular, d.Reflectivity, AddMat); break;
                        case XYRect r: o.Kind = (int)Kind.XYRect; Set(o.P, r.X0, r.X1, r.Y0, r.Y1, r.Z); MatFunc(ref o, r.MaterialFunc, r.Specular, r.Reflectivity, AddMat); break;
                        case XZRect r: o.Kind = (int)Kind.XZRect; Set(o.P, r.X0, r.X1, r.Z0, r.Z1, r.Y); MatFunc(ref o, r.MaterialFunc, r.Specular, r.Reflectivity, AddMat); break;
                        case YZRect r: o.Kind = (int)Kind.YZRect; Set(o.P, r.Y0, r.Y1, r.Z0, r.Z1, r.X); MatFunc(ref o, r.MaterialFunc, r.Specular, r.Reflectivity, AddMat); break;
                        case Box b: o.Kind = (int)Kind.Box; Set(o.P, b.Min.X, b.Min.Y, b.Min.Z, b.Max.X, b.Max.Y, b.Max.Z); MatFunc(ref o, b.MaterialFunc, b.Specular, b.Reflectivity, AddMat); break;
                        case CylinderY c: o.Kind = (int)Kind.CylinderY; Set(o.P, c.Center.X, c.Center.Y, c.Center.Z, c.Radius, c.YMin, c.YMax, c.Capped ? 1f : 0f); o.MatA = o.MatB = AddMat(c.Mat); break;
                        case Triangle t: o.Kind = (int)Kind.Triangle; Set(o.P, t.A.X, t.A.Y, t.A.Z, t.B.X, t.B.Y, t.B.Z, t.C.X, t.C.Y, t.C.Z); o.MatA = o.MatB = AddMat(t.Mat); break;
                        case Mesh m when meshIds.TryGetValue(m, out int knownMesh): o.Kind = (int)Kind.Mesh; o.RefId = knownMesh; break;
                        case Mesh m:
                        {   // MeshBVH's private SoA + tree through the `internal` accessor MeshBVH.ExportFlat (INTEGRATION.md)
                            MeshBVH.Flat f = m.Bvh.ExportFlat();
                            int meshId = meshIds.Count;
                            var tree = new YBvh { NNodes = f.NodeCount, Root = f.Root, NLeafRefs = f.LeafTriIndex.Length,
                                MinX = Pin(f.NodeMinX), MinY = Pin(f.NodeMinY), MinZ = Pin(f.NodeMinZ), MaxX = Pin(f.NodeMaxX), MaxY = Pin(f.NodeMaxY), MaxZ = Pin(f.NodeMaxZ),
                                Left = Pin(f.NodeLeft), Right = Pin(f.NodeRight), Start = Pin(f.NodeStart), Count = Pin(f.NodeCount_), LeafIndex = Pin(f.LeafTriIndex) };
                            var treeBox = new[] { tree };
                            var soa = new YMeshSoa { NTris = f.Ax.Length, Ax = Pin(f.Ax), Ay = Pin(f.Ay), Az = Pin(f.Az), E1x = Pin(f.E1x), E1y = Pin(f.E1y), E1z = Pin(f.E1z),
                                E2x = Pin(f.E2x), E2y = Pin(f.E2y), E2z = Pin(f.E2z), Nx = Pin(f.Nx), Ny = Pin(f.Ny), Nz = Pin(f.Nz),
                                Material = ToMaterialWithTexture(f.TriMat[0]), Bvh = Pin(treeBox) }; // MeshLoader gives every triangle the same material (MeshLoader.cs:58-97)
                            Check(ycge_mesh_upload_soa(ctx, meshId, ref soa));
                            meshIds[m] = meshId;
                            o.Kind = (int)Kind.Mesh; o.RefId = meshId;
                            break;
                        }
                        case VolumeGrid g when volIds.TryGetValue(g, out int knownVol): o.Kind = (int)Kind.Volume; o.RefId = knownVol; break;
                        case VolumeGrid g:
                        {   // the grid's pinned bricked-Morton arrays go over unchanged (VolumeGrid.cs:70-73, 235-252)
                            VolumeGrid.Flat f = g.ExportFlat();
                            int volId = volIds.Count;
                            var palette = new List<int>(); // materialLookup(id, meta) tabulated: a closed table (VoxelMaterialPalette.cs:48-98)
                            for (int id = 0; id < f.PaletteIds; id++) for (int meta = 0; meta < f.PaletteMetaLevels; meta++) palette.Add(AddMat(f.MaterialLookup(id, meta)));
                            int def = AddMat(f.MaterialLookup(int.MaxValue, 0));
                            var vol = new YVolume { Nx = f.Nx, Ny = f.Ny, Nz = f.Nz, MinX = f.MinCorner.X, MinY = f.MinCorner.Y, MinZ = f.MinCorner.Z,
                                SizeX = f.VoxelSize.X, SizeY = f.VoxelSize.Y, SizeZ = f.VoxelSize.Z, Mat = f.MatPtr, Meta = f.MetaPtr,
                                Wireframe = f.Wireframe ? 1 : 0, WireWidthFrac = f.WireWidthFraction, WireMaxDistance = f.WireMaxDistance,
                                PaletteNIds = f.PaletteIds, PaletteMetaLevels = f.PaletteMetaLevels, Palette = Pin(palette.ToArray()), PaletteDefault = def };
                            // palette entries are indices into the append-only material table: they stay valid over later re-syncs
                            Check(ycge_volume_upload(ctx, volId, ref vol));
                            volIds[g] = volId;
                            o.Kind = (int)Kind.Volume; o.RefId = volId;
                            break;
                        }
                        default: throw new InvalidOperationException("Unbounded Hittable"); // BVH.cs:39
                    }
                    objects.Add(o);
                }
                BVH.Flat top = scene.Bvh.ExportFlat(); // the host's own tree is authoritative (tie-breaking depends on it)
                var topTree = new[] { new YBvh { NNodes = top.NodeCount, Root = top.Root, NLeafRefs = top.LeafObjIndex.Length,
                    MinX = Pin(top.NodeMinX), MinY = Pin(top.NodeMinY), MinZ = Pin(top.NodeMinZ), MaxX = Pin(top.NodeMaxX), MaxY = Pin(top.NodeMaxY), MaxZ = Pin(top.NodeMaxZ),
                    Left = Pin(top.NodeLeft), Right = Pin(top.NodeRight), Start = Pin(top.NodeStart), Count = Pin(top.NodeCount_), LeafIndex = Pin(top.LeafObjIndex) } };
                var lights = new YLight[scene.Lights.Count];
                for (int i = 0; i < lights.Length; i++) lights[i] = ToLight(scene.Lights[i]);
                var ys = new YScene
                {
                    BgTopX = scene.BackgroundTop.X, BgTopY = scene.BackgroundTop.Y, BgTopZ = scene.BackgroundTop.Z,
                    BgBotX = scene.BackgroundBottom.X, BgBotY = scene.BackgroundBottom.Y, BgBotZ = scene.BackgroundBottom.Z,
                    AmbR = scene.Ambient.Color.X, AmbG = scene.Ambient.Color.Y, AmbB = scene.Ambient.Color.Z, AmbientIntensity = scene.Ambient.Intensity,
                    IsVolumeScene = scene is VolumeScene ? 1 : 0, // RaytraceRenderer.cs:761
                    NLights = lights.Length, Lights = Pin(lights), NMaterials = materials.Count, Materials = Pin(materials.ToArray()),
                    NObjects = objects.Count, Objects = Pin(objects.ToArray()), Bvh = Pin(topTree)
                };
                Check(ycge_scene_upload(ctx, ref ys));
                uploadedBvh = scene.Bvh;
            }
            finally { foreach (var h in pins) h.Free(); } // the library copies everything during the call
        }

        // Material functions are data, not delegates: every lambda in the engine is Solid/Emissive (constant) or
        // Checker(a, b, scale) (Scenes.cs:408-428).  The factories tag their closures (INTEGRATION.md) so the kind is known.
        private static void MatFunc(ref YObject o, Func<Vec3, Vec3, float, Material> f, float specular, float reflectivity, Func<Material, int> addMat)
        {
            o.OverrideSr = 1; o.Specular = specular; o.Reflectivity = reflectivity; // Surfaces.cs:64-66
            if (f.Target is Scenes.Scenes.CheckerClosure ck) { o.MatA = addMat(ck.A); o.MatB = addMat(ck.B); o.CheckerScale = ck.Scale; }
            else { o.MatA = o.MatB = addMat(f(Vec3.Zero, Vec3.Zero, 0f)); o.CheckerScale = 0f; }
        }

        private static YMaterial ToMaterial(Material m) => new YMaterial
        {   // binary64 scalars are only ever read through (float) casts on the path (RaytraceRenderer.cs:500-559,776)
            AlbedoX = m.Albedo.X, AlbedoY = m.Albedo.Y, AlbedoZ = m.Albedo.Z, Reflectivity = (float)m.Reflectivity,
            EmissionX = m.Emission.X, EmissionY = m.Emission.Y, EmissionZ = m.Emission.Z, Transparency = (float)m.Transparency,
            TransmissionX = m.TransmissionColor.X, TransmissionY = m.TransmissionColor.Y, TransmissionZ = m.TransmissionColor.Z, Ior = (float)m.IndexOfRefraction,
            Specular = (float)m.Specular, TexId = -1, TexWeight = (float)m.TextureWeight, UvScale = (float)m.UVScale
        };
        private static YLight ToLight(PointLight l) => new YLight { Px = l.Position.X, Py = l.Position.Y, Pz = l.Position.Z, Cr = l.Color.X, Cg = l.Color.Y, Cb = l.Color.Z, Intensity = l.Intensity };
        private static float[] V(Vec3 v) => new[] { v.X, v.Y, v.Z };
        private static unsafe void Set(float* p, params float[] v) { for (int i = 0; i < v.Length; i++) p[i] = v[i]; }

        private void AllocCells()
        {
            if (cellsPin.IsAllocated) cellsPin.Free();
            cells = new YCell[fbW * fbH];
            cellsPin = GCHandle.Alloc(cells, GCHandleType.Pinned);
        }

        private void Check(int rc)
        {
            if (rc == 0) return;
            string msg = Marshal.PtrToStringAnsi(ycge_last_error(ctx)) ?? "";
            throw new InvalidOperationException($"ycge error {rc}: {msg}"); // as Win32TerminalRenderer.cs:99-104 does for native failures
        }

        public void Dispose()
        {
            if (ctx != IntPtr.Zero) { ycge_destroy(ctx); ctx = IntPtr.Zero; }
            if (cellsPin.IsAllocated) cellsPin.Free();
        }
    }
}

// In RaytraceEntity.cs (a partial class), next to RaytraceWrapper (:20-28):
//
//     private sealed class CudaRaytraceWrapper : IConsoleRenderer
//     {
//         private readonly CudaRaytraceRenderer inner;
//         public CudaRaytraceWrapper(CudaRaytraceRenderer inner) { this.inner = inner; }
//         public void SetCamera(Vec3 pos, float yaw, float pitch) { inner.SetCamera(pos, yaw, pitch); }
//         public void SetFov(float fovDeg) { inner.SetFov(fovDeg); }
//         public void TryFlipAndBlit(Framebuffer fb) { inner.TryFlipAndBlit(fb); }
//         public void Resize(Framebuffer fb, int superSample) { inner.Resize(fb, superSample); }
//     }
//
// and at the three construction sites (:97-98, :240-241, :262-263):
//
//     this.renderer = new CudaRaytraceWrapper(new CudaRaytraceRenderer(fb, this.activeScene, activeScene.DefaultFovDeg, rtWidth, rtHeight, rtSuperSample));
