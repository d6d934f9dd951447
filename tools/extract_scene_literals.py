"""Generates tests/golden/scene_literals.json from the reference's C# scene factories (RayTracing/Scenes/Scenes.cs), so that the
host mirror's factories (host/ycge_host.cpp) can be checked against the source mechanically instead of by reading.

The factories are straight-line C#: declarations with literal arithmetic, `new X(...)`, `s.Add(...)`, `s.Lights.Add(...)`.  Each
statement is rewritten into Python syntax (type prefixes dropped, `new` dropped, `1.5f` -> binary32, lambdas that return a captured
material -> Constant(m)) and executed against small recording classes that apply the reference's own conversions (Vec3(double,
double,double) casts to float, `float` arithmetic stays binary32).  Needs /root/reference; the JSON it writes is the fixture
tests/test_host.py::test_scene_factories_match_the_reference_source reads.
    python tools/extract_scene_literals.py [path/to/ConsoleGame]
"""
import json
import os
import re
import sys

import numpy as np

F = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FUNCS = {"test": "BuildTestScene", "cornell": "BuildCornellBox", "mirror_spheres": "BuildMirrorSpheresOnChecker",
         "cylinders_disks_triangles": "BuildCylindersDisksAndTriangles", "boxes": "BuildBoxesShowcase", "texture_test": "BuildTextureTestScene"}


def f32(x):
    return float(F(x))


class Vec3:
    def __init__(self, x, y, z):
        self.v = [f32(x), f32(y), f32(z)]  # Vec3(double, double, double) casts each component to float (Vec3.cs:21-26)

    X = property(lambda self: F(self.v[0]))
    Y = property(lambda self: F(self.v[1]))
    Z = property(lambda self: F(self.v[2]))

    def __add__(self, o):  # Vec3.cs:31-35, component-wise binary32
        return Vec3(F(self.v[0]) + F(o.v[0]), F(self.v[1]) + F(o.v[1]), F(self.v[2]) + F(o.v[2]))

    def __sub__(self, o):  # :37-41
        return Vec3(F(self.v[0]) - F(o.v[0]), F(self.v[1]) - F(o.v[1]), F(self.v[2]) - F(o.v[2]))


class Material:  # Material.cs:5-61: scalars are doubles; the path reads them through (float) casts
    def __init__(self, albedo, specular, reflectivity, emission, transparency=0.0, ior=1.5, tint=None):
        self.albedo, self.specular, self.reflectivity, self.emission = albedo.v, float(specular), float(reflectivity), emission.v
        self.transparency, self.ior, self.tint = float(transparency), float(ior), (tint.v if tint else [1.0, 1.0, 1.0])
        self.DiffuseTexture, self.TextureWeight, self.UVScale = None, 1.0, 1.0

    def dump(self):
        return dict(albedo=self.albedo, specular=f32(self.specular), reflectivity=f32(self.reflectivity), emission=self.emission, transparency=f32(self.transparency),
                    ior=f32(self.ior), tint=self.tint, textured=self.DiffuseTexture is not None, tex_weight=f32(self.TextureWeight), uv_scale=f32(self.UVScale))


class MatFunc:
    def __init__(self, a, b, scale):
        self.a, self.b, self.scale = a, b, f32(scale)


def Solid(albedo):  # Scenes.cs:408-411
    m = Material(albedo, 0.0, 0.0, Vec3(0, 0, 0))
    return MatFunc(m, m, 0.0)


def Emissive(emission):  # :413-416
    m = Material(Vec3(0.0, 0.0, 0.0), 0.0, 0.0, emission)
    return MatFunc(m, m, 0.0)


def Checker(a, b, scale):  # :418-428
    return MatFunc(Material(a, 0.0, 0.0, Vec3(0, 0, 0)), Material(b, 0.0, 0.0, Vec3(0, 0, 0)), scale)


def Constant(m):
    return MatFunc(m, m, 0.0)


class Texture:
    def __init__(self, path):
        self.path = path


class Scene:
    def __init__(self):
        self.objects, self.lights = [], []
        self.Lights = self.Objects = self
        self.DefaultCameraPos = None
        self.Ambient, self.BackgroundTop, self.BackgroundBottom = None, None, None

    def Add(self, o):
        (self.lights if o["kind"] == "light" else self.objects).append(o)

    def Update(self, dt):
        pass

    def RebuildBVH(self):
        pass


def obj(kind, p, func=None, specular=None, reflectivity=None, mat=None):
    d = dict(kind=kind, p=[f32(v) for v in p])
    if func is not None:  # flat primitives and Box overwrite Specular / Reflectivity of the function's result (Surfaces.cs:64-66)
        d.update(a=func.a.dump(), b=func.b.dump(), checker_scale=func.scale, override_sr=True, specular=f32(specular), reflectivity=f32(reflectivity))
    else:
        d.update(a=mat.dump(), b=mat.dump(), checker_scale=0.0, override_sr=False)
    return d


def normalized(v):  # Vec3.Normalized (Vec3.cs:98-107) as the Plane / Disk ctors apply it
    l2 = F(F(F(F(v[0]) * F(v[0])) + F(F(v[1]) * F(v[1]))) + F(F(v[2]) * F(v[2])))
    if l2 <= 0:
        return v
    inv = F(F(1) / np.sqrt(l2, dtype=F))
    return [f32(F(v[0]) * inv), f32(F(v[1]) * inv), f32(F(v[2]) * inv)]


NS = dict(
    F=F, Vec3=Vec3, Material=Material, Solid=Solid, Emissive=Emissive, Checker=Checker, Constant=Constant, Scene=Scene, Texture=Texture, true=True, false=False,
    AmbientLight=lambda c, i: dict(color=c.v, intensity=f32(i)),
    PointLight=lambda p, c, i: dict(kind="light", pos=p.v, color=c.v, intensity=f32(i)),
    Sphere=lambda c, r, m: obj("sphere", c.v + [r], mat=m),
    Plane=lambda p, n, f, s, r: obj("plane", p.v + normalized(n.v), f, s, r),
    Disk=lambda c, n, rad, f, s, r: obj("disk", c.v + normalized(n.v) + [rad], f, s, r),
    XYRect=lambda a0, a1, b0, b1, k, f, s, r: obj("xyrect", [a0, a1, b0, b1, k], f, s, r),
    XZRect=lambda a0, a1, b0, b1, k, f, s, r: obj("xzrect", [a0, a1, b0, b1, k], f, s, r),
    YZRect=lambda a0, a1, b0, b1, k, f, s, r: obj("yzrect", [a0, a1, b0, b1, k], f, s, r),
    Box=lambda mn, mx, f, s, r: obj("box", mn.v + mx.v, f, s, r),
    CylinderY=lambda c, rad, y0, y1, capped, m: obj("cylinder_y", c.v + [rad, y0, y1, 1.0 if capped else 0.0], mat=m),
    Triangle=lambda a, b, c, m: obj("triangle", a.v + b.v + c.v, mat=m),
)


def function_body(src, name):
    at = re.search(r"(?:public|private) static Scene " + name + r"\(\)", src).start()
    i = src.index("{", at)
    depth, j = 0, i
    while True:
        depth += {"{": 1, "}": -1}.get(src[j], 0)
        if depth == 0:
            return src[i + 1:j]
        j += 1


def to_python(stmt):
    s = stmt.strip()
    if not s or s.startswith("return"):
        return None
    s = re.sub(r"^(?:[\w\.]+(?:<[^=]*>)?)\s+(\w+)\s*=", r"\1 =", s)             # `Type name = ...` -> `name = ...`
    s = re.sub(r"\(\s*\w+\s*,\s*\w+\s*,\s*\w+\s*\)\s*=>\s*(\w+)", r"Constant(\1)", s)   # (pos, n, u) => capturedMaterial
    s = s.replace("ConsoleGame.Renderer.Texture", "Texture").replace("Vec3.Zero", "Vec3(0, 0, 0)").replace("new ", "")
    s = re.sub(r'@"([^"]*)"', r'r"\1"', s)
    s = re.sub(r"(?<![\w.])(\d+\.\d+|\d+)f\b", r"F(\1)", s)                        # binary32 literals
    return s


CONSOLE_COLORS = ["Black", "DarkBlue", "DarkGreen", "DarkCyan", "DarkRed", "DarkMagenta", "DarkYellow", "Gray", "DarkGray", "Blue", "Green", "Cyan", "Red", "Magenta",
                  "Yellow", "White"]  # System.ConsoleColor, values 0..15


def mesh_swatches(src):
    """MeshSwatches (MeshScenes.cs:12-103): the Palette16 table and the named swatches, evaluated from the source text."""
    table = re.search(r"Palette16 = new Vec3\[\]\s*\{(.*?)\};", src, re.S).group(1)
    palette = [Vec3(*(F(v) for v in m)) for m in re.findall(r"new Vec3\(([\d.]+)f,([\d.]+)f,([\d.]+)f\)", table)]
    assert len(palette) == 16

    def from_console(name):
        return palette[CONSOLE_COLORS.index(name)]

    def scale(name, k):  # :43-49
        k = min(max(F(k), F(0)), F(1))
        v = from_console(name)
        return Vec3(F(v.v[0]) * k, F(v.v[1]) * k, F(v.v[2]) * k)

    out = {}
    for name, expr in re.findall(r"public static readonly Vec3 (\w+) = ([^;]+);", src):
        m = re.match(r"FromConsole\(ConsoleColor\.(\w+)\)", expr)
        if m:
            out[name] = from_console(m.group(1))
            continue
        m = re.match(r"Scale\(ConsoleColor\.(\w+), ([\d.]+)f\)", expr)
        if m:
            out[name] = scale(m.group(1), m.group(2))
            continue
        m = re.match(r"new Vec3\(([\d.]+)f, ([\d.]+)f, ([\d.]+)f\)", expr)
        out[name] = Vec3(*(F(v) for v in m.groups()))
    return out


def mesh_scene_materials(src):
    """Build{Cow,Bunny,Teapot,Dragon}Scene (:108-143): the mesh material (MeshSwatches.Matte / Mirror of a swatch) and the camera."""
    sw = mesh_swatches(src)
    out = {}
    for scene, fn in (("cow", "BuildCowScene"), ("bunny", "BuildBunnyScene"), ("teapot", "BuildTeapotScene"), ("dragon", "BuildDragonScene")):
        body = function_body(src, fn)
        kind, swatch, args = re.search(r"MeshSwatches\.(Matte|Mirror)\(MeshSwatches\.(\w+)((?:, [\d.]+)*)\)", body).groups()
        nums = [float(v) for v in re.findall(r"[\d.]+", args)]
        if kind == "Matte":   # Matte(albedo, specular = 0.10, reflectivity = 0.00) :92-95
            spec, refl = (nums + [0.10, 0.00][len(nums):])[:2]
        else:                 # Mirror(tint, reflectivity = 0.85) :96-99
            spec, refl = 0.0, (nums + [0.85])[0]
        cam = re.search(r"s\.DefaultCameraPos = new Vec3\(([^)]*)\)", body)
        target = re.search(r"targetPos: new Vec3\(([^)]*)\)", body).group(1)
        out[scene] = dict(material=Material(sw[swatch], spec, refl, Vec3(0, 0, 0)).dump(),
                          camera=[f32(v) for v in cam.group(1).split(",")] if cam else None,
                          target_pos=[f32(v.strip().rstrip("f")) for v in target.split(",")])
    return out


def all_meshes_scene(src):
    """BuildAllMeshesScene (:145-158): per mesh the asset, its material (a named local built from a swatch) and targetPos, in order."""
    sw = mesh_swatches(src)
    body = function_body(src, "BuildAllMeshesScene")
    mats = {}
    for var, kind, swatch, args in re.findall(r"Material (\w+) = MeshSwatches\.(Matte|Mirror)\(MeshSwatches\.(\w+)((?:, [\d.]+)*)\);", body):
        nums = [float(v) for v in re.findall(r"[\d.]+", args)]
        spec, refl = ((nums + [0.10, 0.00][len(nums):])[:2]) if kind == "Matte" else (0.0, (nums + [0.85])[0])
        mats[var] = Material(sw[swatch], spec, refl, Vec3(0, 0, 0)).dump()
    out = []
    for asset, var, scale, target in re.findall(r'AddMeshAutoGround\(s, @"assets\\([\w.-]+)", (\w+), scale: ([\d.]+)f, targetPos: new Vec3\(([^)]*)\)\);', body):
        out.append(dict(asset=asset, material=mats[var], scale=f32(scale), target_pos=[f32(v.strip().rstrip("f")) for v in target.split(",")]))
    assert len(out) == 4
    return out


def any_function_body(src, name):
    at = re.search(r"(?:public|private) static \w+ " + name + r"\(", src).start()
    i = src.index("{", at)
    depth, j = 0, i
    while True:
        depth += {"{": 1, "}": -1}.get(src[j], 0)
        if depth == 0:
            return src[i + 1:j]
        j += 1


def drop_guarded_blocks(body, guard):
    """Removes `if(<guard>) { ... }` blocks: the museum's video exhibits exist only when Assets/TestVideo.mp4 does (it does not)."""
    while True:
        m = re.search(r"if\s*\(\s*" + guard + r"\s*\)\s*\{", body)
        if not m:
            return body
        depth, j = 0, m.end() - 1
        while True:
            depth += {"{": 1, "}": -1}.get(body[j], 0)
            if depth == 0:
                break
            j += 1
        body = body[:m.start()] + body[j + 1:]


def inline_material_lambdas(stmt):
    """`(p, n, u) => new Material(...)` -> `Constant(Material(...))` (the lambda ignores its arguments)."""
    while True:
        m = re.search(r"\(\s*\w+\s*,\s*\w+\s*,\s*\w+\s*\)\s*=>\s*new Material\(", stmt)
        if not m:
            return stmt
        depth, j = 0, m.end() - 1
        while True:
            depth += {"(": 1, ")": -1}.get(stmt[j], 0)
            if depth == 0:
                break
            j += 1
        stmt = stmt[:m.start()] + "Constant(new Material(" + stmt[m.end():j + 1] + ")" + stmt[j + 1:]


def run_statements(body, ns):
    """Straight-line C# (declarations, calls, nested plain blocks) executed statement by statement; log / timing calls dropped."""
    body = re.sub(r"//[^\n]*", "", body).replace("{", " ").replace("}", " ")
    for stmt in body.split(";"):
        st = stmt.strip()
        if not st or re.match(r"(Console\.|Stopwatch |\w*[sS]w\.)", st):
            continue
        py = to_python(inline_material_lambdas(st))
        if py:
            exec(py, ns)


def museum_scene(src):
    """TestScenes.BuildTestScene (TestScenes.cs:16-159) with AddCornellBoxRoom (:161-213), TryAddMeshAutoGround (:363-379) and the
    statements of the two voxel dioramas after their cell loops (:256-277, :310-330), all executed from the source text.  The cell
    loops themselves are compared separately (tests/test_host.py re-states them in numpy); the `switch (id)` material lookups are
    parsed case by case."""
    s, base = Scene(), {}
    meshes_present = {"cow.obj", "stanford-bunny.obj", "teapot.obj"}  # the assets shipped with the reference; the dragon is not

    def try_add_mesh(scene, path, mat, scale, target):
        asset = path.replace("\\", "/").split("/")[-1]
        if asset not in meshes_present:
            return
        y = F(F(target.Y + F(0.5)) + F(0.01))
        scene.Add(dict(kind="mesh", asset=asset, material=mat.dump(), scale=f32(scale), translate=[f32(target.X), f32(y), f32(target.Z)]))

    def room(scene, anchor, width, height, left, right, white, power, emissive):
        ns = dict(NS, s=scene, anchor=anchor, width=F(width), height=F(height), leftColor=left, rightColor=right, whiteColor=white, lightPower=F(power), emissive=emissive)
        run_statements(any_function_body(src, "AddCornellBoxRoom"), ns)

    def diorama(name):
        def build(scene, min_corner, *mats):
            body = any_function_body(src, name)
            params = re.search(name + r"\(Scene s, Vec3 minCorner, ([^)]*)\)", src).group(1)
            ns = dict(base, s=scene, minCorner=min_corner, cells=name, **{p.split()[-1]: m for p, m in zip(params.split(","), mats)})
            lookup_src = body[body.index("Func<int, int, Material> materialLookup"):]
            lookup_src = lookup_src[:lookup_src.index("};") + 2]
            table = {}
            for key, expr in re.findall(r"(case \d+|default):\s*return ([^;]+);", lookup_src):
                exec("_m = " + to_python(expr.strip() + " "), ns)
                table["default" if key == "default" else key.split()[1]] = ns["_m"].dump()
            ns["materialLookup"] = table
            tail = body[body.index("Vec3 voxelSize"):].replace(lookup_src, "")
            run_statements(tail, ns)
        return build

    base.update(NS, TryAddMeshAutoGround=try_add_mesh,
                VolumeGrid=lambda cells, mn, size, lookup: dict(kind="volume", cells=cells, min_corner=mn.v, voxel_size=size.v, lookup=lookup))
    ns = dict(base, s=s, AddCornellBoxRoom=room, BuildVolumeDioramaA=diorama("BuildVolumeDioramaA"), BuildVolumeDioramaB=diorama("BuildVolumeDioramaB"),
              BuildVideoDiorama=lambda *a: None)
    body = drop_guarded_blocks(any_function_body(src, "BuildTestScene"), r'File\.Exists\("Assets/TestVideo\.mp4"\)')
    run_statements(body, ns)
    s = ns["s"]  # `Scene s = new Scene();` is the factory's first statement
    return dict(ambient=s.Ambient, bg_top=s.BackgroundTop.v, bg_bottom=s.BackgroundBottom.v, camera=s.DefaultCameraPos.v,
                lights=[{k: v for k, v in l.items() if k != "kind"} for l in s.lights], objects=s.objects)


def renderer_constants(ref):
    """The compile-time constants of the path (RaytraceRenderer.cs:31-43,:65,:222,:281; ToneMapper.cs:8-21) = ycge_default_params."""
    rr = open(os.path.join(ref, "RayTracing", "RaytraceRenderer.cs"), encoding="utf-8-sig").read()
    tm = open(os.path.join(ref, "RayTracing", "ToneMapper.cs"), encoding="utf-8-sig").read()

    def num(src, name):
        m = re.search(r"\b" + name + r"\s*[=:]\s*([-\d.e]+|0x[0-9A-Fa-f]+|true|false)(?:f|UL)?\b", src)
        v = m.group(1)
        return int(v, 16) if v.startswith("0x") else (v == "true") if v in ("true", "false") else (int(v) if re.fullmatch(r"-?\d+", v) else f32(v))

    return dict(diffuse_bounces=num(rr, "DiffuseBounces"), max_mirror_bounces=num(rr, "MaxMirrorBounces"), max_refractions=num(rr, "MaxRefractions"),
                mirror_threshold=num(rr, "MirrorThreshold"), eps=num(rr, "Eps"), seed_salt=num(rr, "SeedSalt"), taa_alpha=num(rr, "taaAlpha"),
                motion_trans_reset=num(rr, "MotionTransReset"), motion_rot_reset=num(rr, "MotionRotReset"), diffuse_sigma_deg=num(rr, "DiffuseSigmaDeg"),
                luminance_pad=num(rr, "luminancePad"), atrous_iterations=num(rr, "iterations"), c_phi=num(rr, "cPhi"), n_phi=num(rr, "nPhi"), z_phi=num(rr, "zPhi"),
                a_phi=num(rr, "aPhi"), tone_exposure=num(tm, "toneExposure"), tone_gamma=num(tm, "toneGamma"), auto_exposure=num(tm, "autoExposure"),
                ae_key=num(tm, "aeKey"), ae_speed=num(tm, "aeSpeed"), ae_min=num(tm, "aeMin"), ae_max=num(tm, "aeMax"), saturation=num(tm, "toneSaturation"),
                vibrance=num(tm, "toneVibrance"))


def voxel_palette(ref):
    """VoxelMaterialPalette (Scenes/VoxelMaterialPalette.cs:8-98): MaterialLookup(id, meta) for every block id and meta 0..2, evaluated
    from the two switch statements (Normalize, CreateMaterial), the Palette16 table and PalMat."""
    src = open(os.path.join(ref, "RayTracing", "Scenes", "VoxelMaterialPalette.cs"), encoding="utf-8-sig").read()
    blocks = dict(re.findall(r"public const int (\w+) = (\d+);", open(os.path.join(ref, "RayTracing", "Scenes", "WorldGeneration", "WorldGenSettings.cs"), encoding="utf-8-sig").read().split("class Blocks")[1].split("}")[0]))
    table = re.search(r"Palette16 = new Vec3\[\]\s*\{(.*?)\};", src, re.S).group(1)
    palette = [[f32(v) for v in m] for m in re.findall(r"new Vec3\(([\d.]+),([\d.]+),([\d.]+)\)", table)]
    spec, refl = re.search(r"new Material\(c, ([\d.]+), ([\d.]+),", src).groups()
    norm_src = src[src.index("Normalize(int id, int meta)"):src.index("CreateMaterial((int id, int meta) key)")]
    normalize = {}
    for name, nid, meta in re.findall(r"case WorldGenSettings\.Blocks\.(\w+): return \((\d+), (0|Clamp\(meta, 0, 2\))\);", norm_src):
        normalize[int(blocks[name])] = (int(nid), meta != "0")
    norm_default = tuple(int(v) for v in re.search(r"default: return \((\d+), (\d+)\);", norm_src).groups())
    create_src = src[src.index("CreateMaterial((int id, int meta) key)"):src.index("private static void Prewarm()")]
    create = {}
    for m in re.finditer(r"case (\d+):\s*(?:return PalMat\((\d+)\);|switch \(key\.meta\)\s*\{(.*?)\})", create_src, re.S):
        if m.group(2) is not None:
            create[int(m.group(1))] = {None: int(m.group(2))}
        else:
            inner = {int(a): int(b) for a, b in re.findall(r"case (\d+): return PalMat\((\d+)\);", m.group(3))}
            inner[None] = int(re.search(r"default: return PalMat\((\d+)\);", m.group(3)).group(1))
            create[int(m.group(1))] = inner
    lookup = {}
    for bid in range(0, 13):
        for meta in range(3):
            if bid in normalize:
                nid, nmeta = normalize[bid][0], (min(max(meta, 0), 2) if normalize[bid][1] else 0)
            else:
                nid, nmeta = norm_default
            entry = create[nid]
            lookup[f"{bid},{meta}"] = palette[entry.get(nmeta, entry[None])]
    return dict(lookup=lookup, specular=f32(spec), reflectivity=f32(refl), default=palette[create[norm_default[0]].get(norm_default[1], create[norm_default[0]][None])])


def tables(ref):
    """Small tables of the path: BlueNoise8x8 (RaytraceSampler.cs:9-19), the 16-colour palette (Renderer/Chexel.cs:11-29), the colour
    cube thresholds (ANSITerminalRenderer.cs:288-296)."""
    rs = open(os.path.join(ref, "RayTracing", "RaytraceSampler.cs"), encoding="utf-8-sig").read()
    blue = [[int(v) for v in row.split(",")] for row in re.findall(r"\{\s*((?:\d+\s*,\s*){7}\d+)\s*\}", rs)]
    ch = open(os.path.join(ref, "Renderer", "Chexel.cs"), encoding="utf-8-sig").read()
    pal = re.search(r"s_Palette16 = new Vec3\[\]\s*\{(.*?)\};", ch, re.S).group(1)
    palette = [[f32(v) for v in m] for m in re.findall(r"new Vec3\(([\d.]+)f,([\d.]+)f,([\d.]+)f\)", pal)]
    an = open(os.path.join(ref, "Renderer", "ANSITerminalRenderer.cs"), encoding="utf-8-sig").read()
    cube = an[an.index("ToCubeLevelSrgb(byte v)"):]
    thresholds = [[int(a), int(b)] for a, b in re.findall(r"if \(v < (\d+)\) return (\d+);", cube[:cube.index("}")])]
    assert len(blue) == 8 and len(palette) == 16 and len(thresholds) == 5
    return dict(blue_noise=blue, palette16=palette, cube_thresholds=thresholds)


def extract(src, name):
    body = re.sub(r"//[^\n]*", "", function_body(src, name))
    ns = dict(NS, FloorMat=Solid, MeshBVH=type("MeshBVH", (), {}))  # FloorMat (MeshScenes.cs:372-375) is Solid by another name
    for stmt in body.split(";"):
        py = to_python(stmt)
        if py:
            exec(py, ns)
    s = ns["s"]
    return dict(ambient=s.Ambient, bg_top=s.BackgroundTop.v, bg_bottom=s.BackgroundBottom.v, lights=[{k: v for k, v in l.items() if k != "kind"} for l in s.lights], objects=s.objects)


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/ConsoleGame"
    src = open(os.path.join(ref, "RayTracing", "Scenes", "Scenes.cs"), encoding="utf-8-sig").read()
    out = {scene: extract(src, fn) for scene, fn in FUNCS.items()}
    msrc = open(os.path.join(ref, "RayTracing", "Scenes", "MeshScenes.cs"), encoding="utf-8-sig").read()
    out["mesh_base"] = extract(msrc, "NewBaseScene")
    out["mesh_scenes"] = mesh_scene_materials(msrc)
    out["all_meshes"] = all_meshes_scene(msrc)
    out["museum"] = museum_scene(open(os.path.join(ref, "RayTracing", "Scenes", "TestScenes.cs"), encoding="utf-8-sig").read())
    out["params"] = renderer_constants(ref)
    out["voxel_palette"] = voxel_palette(ref)
    out["tables"] = tables(ref)
    dst = os.path.join(ROOT, "tests", "golden", "scene_literals.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(dst, {k: (len(v["objects"]), len(v["lights"])) for k, v in out.items() if "objects" in v}, out["params"])
