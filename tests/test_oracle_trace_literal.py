"""The trace stage pinned by a second, independent restatement: MakeJitteredRay, the per-pixel RNG, TraceFull with its explicit
work stack, ComputeTransmittanceToLight, OrenNayarBRDF, CosineSampleHemisphere, Scene.Hit -> BVH.Hit / BoxHitFast, Sphere.Hit,
Plane.Hit, Disk.Hit, CylinderY.Hit, Triangle.Hit (scalar path), the three axis rects, Box.Hit (six rects, shrinking closest),
Mesh.Hit -> MeshBVH.Hit with its sign-indexed
BoxHitFast and the deferred-division TriHit (MeshBVH.cs:132-332), VolumeGrid.Hit (the DDA with its entry-axis / tie rules, bricked
Morton addressing and the binary64 wireframe test, VolumeGrid.cs:99-355), Scene.Occluded for volume scenes and the Checker
material function, transcribed from
the C# source into numpy binary32 scalars, one operation at a time (RayTracing/RaytraceRenderer.cs:413-620,:757-831, RaytraceSampler.cs, Objects/BVH.cs:99-236, BoundedObjects.cs:31-69,
:78-115, Surfaces.cs:184-358, Scenes/Scenes.cs:418-428).  It shares nothing with the oracle's C++ but the transcendental functions
of include/ycge_detmath.h (sin, cos, tan) and the scene description; the tree it walks is the HOST mirror's, built by a
third implementation.  Radiance, G-buffer, sky flag and primary ids of every pixel must equal the oracle's bit for bit,
on a scene with a true mirror (reflectivity 0.9 = MirrorThreshold), one with a checker floor, the Cornell box (emissive
rect, closed room, boxes), the boxes showcase (plane + boxes) and a mesh scene (120-triangle knot over the ground plane, the
path of the Dragon workload), the voxel test grid and a 32x32x32 voxel world of chunk grids (a VolumeScene: binary shadow
rays; its test grid holds a transparent block, so the Fresnel split with Refract / FresnelSchlick runs too) and the
cylinder / disk / triangle showcase, and the texture gallery (SampleAlbedo + Texture.SampleBilinear with U, V from rects, box
faces, a triangle and a mesh, blended weights, tiling, a textured glass pane), over two frames.
"""
import numpy as np
import pytest

from yetanotherconsolegameengine_b200 import api
from oracle_binding import Oracle, load_oracle

F = np.float32
M64 = (1 << 64) - 1
FLT_MAX = F(3.4028234663852886e38)
BLUE = [[0, 32, 8, 40, 2, 34, 10, 42], [48, 16, 56, 24, 50, 18, 58, 26], [12, 44, 4, 36, 14, 46, 6, 38], [60, 28, 52, 20, 62, 30, 54, 22],
        [3, 35, 11, 43, 1, 33, 9, 41], [51, 19, 59, 27, 49, 17, 57, 25], [15, 47, 7, 39, 13, 45, 5, 37], [63, 31, 55, 23, 61, 29, 53, 21]]  # RaytraceSampler.cs:9-19

EPS, MIRROR_THRESHOLD, MAX_MIRROR, MAX_REFR, DIFFUSE_BOUNCES = F(1e-4), F(0.9), 2, 2, 1  # RaytraceRenderer.cs:31-36
PI, SIGMA_DEG = F(3.14159265358979323846), F(25.0)                                       # :63-65
MATHF_PI = F(3.14159274)


def v3(x, y, z):
    return np.array([x, y, z], F)


def dot(a, b):  # Vec3.Dot: (x*x' + y*y') + z*z'
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def cross(a, b):
    return v3(F(F(a[1] * b[2]) - F(a[2] * b[1])), F(F(a[2] * b[0]) - F(a[0] * b[2])), F(F(a[0] * b[1]) - F(a[1] * b[0])))


def normalized(v):  # Vec3.cs:98-107
    l2 = F(F(F(v[0] * v[0]) + F(v[1] * v[1])) + F(v[2] * v[2]))
    if l2 <= 0:
        return v
    inv = F(F(1) / np.sqrt(l2, dtype=F))
    return v3(v[0] * inv, v[1] * inv, v[2] * inv)


def vdiv(v, s):  # Vec3 operator / (Vec3.cs:68-71): multiply by the reciprocal
    inv = F(F(1) / s)
    return v3(v[0] * inv, v[1] * inv, v[2] * inv)


def max_f(a, b):  # MathF.Max: NaN propagates
    if a != a:
        return a
    if b != b:
        return b
    return a if a > b else b


def min_f(a, b):
    if a != a:
        return a
    if b != b:
        return b
    return a if a < b else b


def frac(v):
    return F(v - np.floor(v))


def sm64(z):  # RaytraceSampler.cs:71-80
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


class Rng:  # RaytraceSampler.cs:36-53
    def __init__(self, seed):
        self.state = seed if seed != 0 else 0x9E3779B97F4A7C15

    def next_unit(self):
        self.state = sm64(self.state)
        m24 = self.state >> 40
        return F(F(F(m24) + F(0.5)) * F(1.0 / 16777216.0))


def per_frame_seed(x, y, frame, salt=0x9E3779B97F4A7C15):  # :56-68 with jx = jy = 0
    h = 1469598103934665603
    for v, k in ((x, 0x9E3779B97F4A7C15), (y, 0xC2B2AE3D27D4EB4F), (frame, 0x165667B19E3779F9)):
        h ^= ((v & M64) * k) & M64
        h = sm64(h)
    h = sm64(h)  # h ^= 0
    return sm64(h ^ salt)


class LiteralTracer:
    def __init__(self, scene: api.HostScene, lib):
        flat = scene.flat.contents
        self.mats = [flat.materials[i] for i in range(flat.n_materials)]
        self.objs = [flat.objects[i] for i in range(flat.n_objects)]
        self.lights = [(v3(*flat.lights[i].pos), v3(*flat.lights[i].color), F(flat.lights[i].intensity)) for i in range(flat.n_lights)]
        self.bg_top, self.bg_bottom = v3(*flat.bg_top), v3(*flat.bg_bottom)
        self.amb_c, self.amb_i = v3(*flat.ambient_color), F(flat.ambient_intensity)
        self.tree = scene.bvh_arrays(-1)
        self.meshes = []
        for i in range(scene.n_meshes):  # MeshBVH.cs:18-39: triangle SoA + the host mirror's tree
            m = scene.mesh(i).contents
            soa = {k: np.ctypeslib.as_array(getattr(m, k), shape=(m.n_tris,)).copy() for k in ("ax", "ay", "az", "e1x", "e1y", "e1z", "e2x", "e2y", "e2z", "nx", "ny", "nz")}
            mm = m.material
            mat = dict(albedo=v3(*mm.albedo), refl=F(mm.reflectivity), emission=v3(*mm.emission), transparency=F(mm.transparency), tint=v3(*mm.transmission), ior=F(mm.ior),
                       tex=mm.tex_id, tex_weight=F(mm.tex_weight), uv_scale=F(mm.uv_scale))
            self.meshes.append(dict(soa=soa, tree=scene.bvh_arrays(i), mat=mat))
        self.textures = [scene.texture(i) for i in range(scene.n_textures)]  # (h, w) uint32, byte 0 = R (Texture.cs:81-90)
        self.volumes = []
        for i in range(scene.n_volumes):  # VolumeGrid.cs:25-32: the int arrays in bricked-Morton order + the material lookup as a table
            v = scene.volume(i).contents
            nb = ((v.nx + 7) >> 3) * ((v.ny + 7) >> 3) * ((v.nz + 7) >> 3) * 512
            levels = max(1, v.palette_meta_levels)
            self.volumes.append(dict(n=(v.nx, v.ny, v.nz), mn=v3(*v.min_corner), size=v3(*v.voxel_size),
                                     mat=np.ctypeslib.as_array(v.mat, shape=(nb,)).copy(), meta=np.ctypeslib.as_array(v.meta, shape=(nb,)).copy(),
                                     wire=bool(v.wireframe), wire_frac=F(v.wire_width_frac), wire_max=F(v.wire_max_distance),
                                     palette=np.ctypeslib.as_array(v.palette, shape=(v.palette_n_ids * levels,)).copy(), n_ids=v.palette_n_ids, levels=levels,
                                     default=v.palette_default))
        self.is_volume_scene = bool(flat.is_volume_scene)
        self.sin = lambda x: F(lib.yo_math(3, float(x), 0.0))
        self.cos = lambda x: F(lib.yo_math(4, float(x), 0.0))
        self.tan = lambda x: F(lib.yo_math(5, float(x), 0.0))
        self.pow = lambda x, y: F(lib.yo_math(2, float(x), float(y)))
        self.rays = 0

    # ---- materials: constant or Checker(a, b, scale) (Scenes.cs:418-428), then the object's Specular / Reflectivity (Surfaces.cs:279-281)
    def material(self, o, pos):
        m = self.mats[o.mat_a]
        if o.checker_scale != 0.0:
            cx = int(np.floor(F(pos[0] / F(o.checker_scale))))
            cz = int(np.floor(F(pos[2] / F(o.checker_scale))))
            m = self.mats[o.mat_a if ((cx + cz) & 1) == 0 else o.mat_b]
        refl = F(o.reflectivity) if o.override_sr else F(m.reflectivity)
        return dict(albedo=v3(*m.albedo), refl=refl, emission=v3(*m.emission), transparency=F(m.transparency), tint=v3(*m.transmission), ior=F(m.ior),
                    tex=m.tex_id, tex_weight=F(m.tex_weight), uv_scale=F(m.uv_scale))

    # ---- SampleAlbedo (RaytraceRenderer.cs:724-735) + Texture.SampleBilinear (Renderer/Texture.cs:143-162)
    def sample_albedo(self, m, u, v):
        if m.get("tex", -1) < 0 or m["tex_weight"] <= 0:
            return m["albedo"]
        tex = self.textures[m["tex"]]
        h, w = tex.shape
        tiles = F(max(1e-6, float(m["uv_scale"])))
        u, v = F(u * tiles), F(v * tiles)
        u, v = F(u - np.floor(u)), F(v - np.floor(v))
        fx, fy = F(u * F(w - 1)), F(v * F(h - 1))
        x0, y0 = int(np.floor(fx)), int(np.floor(fy))
        x1, y1 = (x0 + 1) % w, (y0 + 1) % h
        tx, ty = F(fx - F(x0)), F(fy - F(y0))
        texel = lambda x, y: [F(F((int(tex[y, x]) >> s) & 255) / F(255)) for s in (0, 8, 16)]
        lerp = lambda a, b, t: [F(F(a[k] * F(F(1) - t)) + F(b[k] * t)) for k in range(3)]
        c = lerp(lerp(texel(x0, y0), texel(x1, y0), tx), lerp(texel(x0, y1), texel(x1, y1), tx), ty)
        c = [min(max(x, F(0)), F(1)) for x in c]
        t = min(max(m["tex_weight"], F(0)), F(1))
        return v3(*(min(max(F(F(m["albedo"][k] * F(F(1) - t)) + F(c[k] * t)), F(0)), F(1)) for k in range(3)))

    # ---- Sphere.Hit BoundedObjects.cs:31-69
    def sphere_hit(self, o, ro, rd, t_min, t_max):
        c, radius = v3(*o.p[0:3]), F(o.p[3])
        ox, oy, oz = F(ro[0] - c[0]), F(ro[1] - c[1]), F(ro[2] - c[2])
        dx, dy, dz = rd
        a = F(F(F(dx * dx) + F(dy * dy)) + F(dz * dz))
        half_b = F(F(F(ox * dx) + F(oy * dy)) + F(oz * dz))
        cc = F(F(F(F(ox * ox) + F(oy * oy)) + F(oz * oz)) - F(radius * radius))
        disc = F(F(half_b * half_b) - F(a * cc))
        if disc < 0:
            return None
        s = np.sqrt(disc, dtype=F)
        inv_a = F(F(1) / a)
        t = F(F(-half_b - s) * inv_a)
        if t < t_min or t > t_max:
            t = F(F(-half_b + s) * inv_a)
            if t < t_min or t > t_max:
                return None
        p = v3(F(ro[0] + F(t * dx)), F(ro[1] + F(t * dy)), F(ro[2] + F(t * dz)))
        inv_r = F(F(1) / radius)
        n = v3(F(F(p[0] - c[0]) * inv_r), F(F(p[1] - c[1]) * inv_r), F(F(p[2] - c[2]) * inv_r))
        return dict(t=t, P=p, N=n, mat=self.material(o, p))

    # ---- XYRect / XZRect / YZRect.Hit (Surfaces.cs:184-214, :256-286, :328-358); k = index of the constant coordinate,
    #      (a, b) = the two free coordinates in the order the class names them
    def rect_hit(self, o, k, a, b, a0, a1, b0, b1, c, ro, rd, t_min, t_max):
        dir_k = rd[k]
        adir = abs(dir_k)
        safe = np.copysign(max_f(adir, F(1e-8)), dir_k)
        t = F(F(c - ro[k]) / safe)
        pa, pb = F(ro[a] + F(t * rd[a])), F(ro[b] + F(t * rd[b]))
        ok = adir >= F(1e-8) and t_min <= t <= t_max and a0 <= pa <= a1 and b0 <= pb <= b1
        if not ok:
            return None
        p, n = v3(0, 0, 0), v3(0, 0, 0)
        p[k], p[a], p[b] = c, pa, pb
        n[k] = np.copysign(F(1), -dir_k)
        return dict(t=t, P=p, N=n, mat=self.material(o, p), U=F(F(pa - a0) * F(F(1) / F(a1 - a0))), V=F(F(pb - b0) * F(F(1) / F(b1 - b0))))

    # ---- Plane.Hit (Surfaces.cs:39-71); p = point.xyz, normal.xyz (normalised by the ctor, :21), ndotPoint :25
    def plane_hit(self, o, ro, rd, t_min, t_max):
        pt, n = v3(*o.p[0:3]), v3(*o.p[3:6])
        ndot_point = F(F(F(n[0] * pt[0]) + F(n[1] * pt[1])) + F(n[2] * pt[2]))
        denom = F(F(F(n[0] * rd[0]) + F(n[1] * rd[1])) + F(n[2] * rd[2]))
        if F(-1e-6) < denom < F(1e-6):
            return None
        t = F(F(ndot_point - F(F(F(n[0] * ro[0]) + F(n[1] * ro[1])) + F(n[2] * ro[2]))) / denom)
        if t < t_min or t > t_max:
            return None
        p = v3(F(ro[0] + F(t * rd[0])), F(ro[1] + F(t * rd[1])), F(ro[2] + F(t * rd[2])))
        return dict(t=t, P=p, N=n if denom < 0 else v3(-n[0], -n[1], -n[2]), mat=self.material(o, p))

    # ---- CylinderY.Hit (BoundedObjects.cs:148-247); p = center.xyz, radius, yMin, yMax, capped
    def cylinder_hit(self, o, ro, rd, t_min, t_max):
        cx, cz, radius, y_min, y_max, capped = F(o.p[0]), F(o.p[2]), F(o.p[3]), F(o.p[4]), F(o.p[5]), o.p[6] != 0.0
        radius2 = F(radius * radius)
        ox, oy, oz = F(ro[0] - cx), ro[1], F(ro[2] - cz)
        dx, dy, dz = rd
        a = F(F(dx * dx) + F(dz * dz))
        hit_t, hit_n, hit = FLT_MAX, v3(0, 0, 0), False
        if a > F(1e-12):
            half_b = F(F(ox * dx) + F(oz * dz))
            c = F(F(F(ox * ox) + F(oz * oz)) - radius2)
            disc = F(F(half_b * half_b) - F(a * c))
            if disc >= 0:
                s = np.sqrt(disc, dtype=F)
                inv_a = F(F(1) / a)
                for root in (F(F(-half_b - s) * inv_a), F(F(-half_b + s) * inv_a)):
                    if hit:
                        break
                    if t_min < root < t_max:
                        y = F(oy + F(root * dy))
                        if y_min <= y <= y_max:
                            hit_t, hit = root, True
                            hit_n = v3(F(F(ox + F(root * dx)) / radius), F(0), F(F(oz + F(root * dz)) / radius))
        if capped and abs(dy) > F(1e-8):
            for y_cap, ny in ((y_max, F(1)), (y_min, F(-1))):
                t_cap = F(F(y_cap - oy) / dy)
                if t_min < t_cap < t_max:
                    rx, rz = F(ox + F(t_cap * dx)), F(oz + F(t_cap * dz))
                    if F(F(rx * rx) + F(rz * rz)) <= radius2 and t_cap < hit_t:
                        hit_t, hit_n, hit = t_cap, v3(0, ny, 0), True
        if not hit:
            return None
        p = v3(F(ro[0] + F(hit_t * dx)), F(ro[1] + F(hit_t * dy)), F(ro[2] + F(hit_t * dz)))
        n = hit_n if dot(hit_n, rd) < 0 else v3(-hit_n[0], -hit_n[1], -hit_n[2])
        return dict(t=hit_t, P=p, N=n, mat=self.material(o, p))

    # ---- Disk.Hit (Surfaces.cs:108-142): the radius test uses dx, dz only; p = center.xyz, normal.xyz (normalised), radius
    def disk_hit(self, o, ro, rd, t_min, t_max):
        c, n, radius = v3(*o.p[0:3]), v3(*o.p[3:6]), F(o.p[6])
        denom = dot(n, rd)
        adenom = abs(denom)
        safe = np.copysign(max_f(adenom, F(1e-8)), denom)
        t = F(F(dot(n, c) - dot(n, ro)) / safe)
        p = v3(F(ro[0] + F(t * rd[0])), F(ro[1] + F(t * rd[1])), F(ro[2] + F(t * rd[2])))
        ddx, ddz = F(p[0] - c[0]), F(p[2] - c[2])
        rr = F(F(ddx * ddx) + F(ddz * ddz))
        if not (adenom >= F(1e-6) and t_min <= t <= t_max and rr <= F(radius * radius)):
            return None
        return dict(t=t, P=p, N=n if denom < 0 else v3(-n[0], -n[1], -n[2]), mat=self.material(o, p))

    # ---- Triangle.Hit, scalar path (Triangle.cs:130-175); edges and normal as the ctor derives them (:36-44)
    def triangle_hit(self, o, ro, rd, t_min, t_max):
        a, b, c = v3(*o.p[0:3]), v3(*o.p[3:6]), v3(*o.p[6:9])
        e1, e2 = [F(b[k] - a[k]) for k in range(3)], [F(c[k] - a[k]) for k in range(3)]
        nn = [F(F(e1[1] * e2[2]) - F(e1[2] * e2[1])), F(F(e1[2] * e2[0]) - F(e1[0] * e2[2])), F(F(e1[0] * e2[1]) - F(e1[1] * e2[0]))]
        inv_len = F(F(1) / max_f(F(1e-20), np.sqrt(F(F(F(nn[0] * nn[0]) + F(nn[1] * nn[1])) + F(nn[2] * nn[2])), dtype=F)))
        n = v3(nn[0] * inv_len, nn[1] * inv_len, nn[2] * inv_len)
        px = F(F(rd[1] * e2[2]) - F(rd[2] * e2[1]))
        py = F(F(rd[2] * e2[0]) - F(rd[0] * e2[2]))
        pz = F(F(rd[0] * e2[1]) - F(rd[1] * e2[0]))
        det = F(F(F(e1[0] * px) + F(e1[1] * py)) + F(e1[2] * pz))
        if abs(det) < F(1e-8):
            return None
        inv_det = F(F(1) / det)
        sx, sy, sz = F(ro[0] - a[0]), F(ro[1] - a[1]), F(ro[2] - a[2])
        u = F(F(F(F(sx * px) + F(sy * py)) + F(sz * pz)) * inv_det)
        if u < 0 or u > 1:
            return None
        qx = F(F(sy * e1[2]) - F(sz * e1[1]))
        qy = F(F(sz * e1[0]) - F(sx * e1[2]))
        qz = F(F(sx * e1[1]) - F(sy * e1[0]))
        v = F(F(F(F(rd[0] * qx) + F(rd[1] * qy)) + F(rd[2] * qz)) * inv_det)
        if v < 0 or F(u + v) > 1:
            return None
        t = F(F(F(F(e2[0] * qx) + F(e2[1] * qy)) + F(e2[2] * qz)) * inv_det)
        if t < t_min or t > t_max:
            return None
        p = v3(F(ro[0] + F(t * rd[0])), F(ro[1] + F(t * rd[1])), F(ro[2] + F(t * rd[2])))
        nd = F(F(F(n[0] * rd[0]) + F(n[1] * rd[1])) + F(n[2] * rd[2]))
        return dict(t=t, P=p, N=n if nd < 0 else v3(-n[0], -n[1], -n[2]), mat=self.material(o, p), U=u, V=v)

    # ---- MeshBVH.TriHit (MeshBVH.cs:239-304): division deferred, bounds scaled by |det|
    @staticmethod
    def tri_hit(soa, i, ro, rd, t_min, t_max):
        e1, e2, a = [soa[k][i] for k in ("e1x", "e1y", "e1z")], [soa[k][i] for k in ("e2x", "e2y", "e2z")], [soa[k][i] for k in ("ax", "ay", "az")]
        px = F(F(rd[1] * e2[2]) - F(rd[2] * e2[1]))
        py = F(F(rd[2] * e2[0]) - F(rd[0] * e2[2]))
        pz = F(F(rd[0] * e2[1]) - F(rd[1] * e2[0]))
        det = F(F(F(e1[0] * px) + F(e1[1] * py)) + F(e1[2] * pz))
        if F(-1e-8) < det < F(1e-8):
            return None
        sx, sy, sz = F(ro[0] - a[0]), F(ro[1] - a[1]), F(ro[2] - a[2])
        u_num = F(F(F(sx * px) + F(sy * py)) + F(sz * pz))
        sgn = F(1) if det > 0 else F(-1)
        det_abs, u_s = F(det * sgn), F(u_num * sgn)
        if u_s < 0 or u_s > det_abs:
            return None
        qx = F(F(sy * e1[2]) - F(sz * e1[1]))
        qy = F(F(sz * e1[0]) - F(sx * e1[2]))
        qz = F(F(sx * e1[1]) - F(sy * e1[0]))
        v_num = F(F(F(rd[0] * qx) + F(rd[1] * qy)) + F(rd[2] * qz))
        v_s = F(v_num * sgn)
        if v_s < 0 or F(u_s + v_s) > det_abs:
            return None
        t_num = F(F(F(e2[0] * qx) + F(e2[1] * qy)) + F(e2[2] * qz))
        t_s = F(t_num * sgn)
        if t_s < F(t_min * det_abs) or t_s > F(t_max * det_abs):
            return None
        inv_det = F(F(1) / det)
        return F(t_num * inv_det), F(u_num * inv_det), F(v_num * inv_det)

    # ---- MeshBVH.BoxHitFast (MeshBVH.cs:308-332): sign-indexed slabs with early outs
    @staticmethod
    def mesh_box_hit(box, ro, inv, sign, t_min, t_max):
        for k in range(3):
            lo, hi = (box[k], box[3 + k]) if sign[k] == 0 else (box[3 + k], box[k])
            t_en, t_ex = F(F(lo - ro[k]) * inv[k]), F(F(hi - ro[k]) * inv[k])
            if t_en > t_min:
                t_min = t_en
            if t_ex < t_max:
                t_max = t_ex
            if k < 2 and t_max < t_min:
                return False, t_min
        return bool(t_max >= t_min), t_min

    # ---- Mesh.Hit -> MeshBVH.Hit (Mesh.cs:23-26, MeshBVH.cs:132-236)
    def mesh_hit(self, o, ro, rd, t_min, t_max):
        mesh = self.meshes[o.ref_id]
        tr, soa = mesh["tree"], mesh["soa"]
        if tr["root"] < 0:
            return None
        inv = [F(F(1) / rd[0]), F(F(1) / rd[1]), F(F(1) / rd[2])]
        sign = [1 if v < 0 else 0 for v in inv]
        closest, best = t_max, None
        stack = [tr["root"]]
        while stack:
            ni = stack.pop()
            hit, _ = self.mesh_box_hit(tr["boxes"][ni], ro, inv, sign, t_min, closest)
            if not hit:
                continue
            left, right, start, count = (int(v) for v in tr["lrsc"][ni])
            if count > 0:
                for i in range(count):
                    tri = int(tr["leaf"][start + i])
                    tuv = self.tri_hit(soa, tri, ro, rd, t_min, closest)
                    if tuv is not None:
                        t = closest = tuv[0]
                        p = v3(F(ro[0] + F(t * rd[0])), F(ro[1] + F(t * rd[1])), F(ro[2] + F(t * rd[2])))
                        n = v3(soa["nx"][tri], soa["ny"][tri], soa["nz"][tri])
                        ndotd = F(F(F(n[0] * rd[0]) + F(n[1] * rd[1])) + F(n[2] * rd[2]))
                        best = dict(t=t, P=p, N=n if ndotd < 0 else v3(-n[0], -n[1], -n[2]), mat=mesh["mat"], sub=tri, U=tuv[1], V=tuv[2])
            else:
                hit_l = hit_r = False
                l_near = r_near = F(0)
                if left >= 0:
                    hit_l, l_near = self.mesh_box_hit(tr["boxes"][left], ro, inv, sign, t_min, closest)
                if right >= 0:
                    hit_r, r_near = self.mesh_box_hit(tr["boxes"][right], ro, inv, sign, t_min, closest)
                if hit_l and hit_r:
                    stack += [right, left] if l_near < r_near else [left, right]
                elif hit_l:
                    stack.append(left)
                elif hit_r:
                    stack.append(right)
        return best

    # ---- VolumeGrid.Hit (VolumeGrid.cs:99-231) with RayAabb / Slab (:319-355), IndexOf / Morton3_3bits (:235-252), IsWireOnFace
    #      (:256-283).  The racy centre-block highlight (:181-186) needs |u - 0.5| <= 1e-6 and is unreachable at even resolutions.
    def volume_hit(self, o, ro, rd, t_min, t_max):
        g = self.volumes[o.ref_id]
        nx, ny, nz = g["n"]
        mn, size = g["mn"], g["size"]
        mx = [F(mn[k] + F(F(g["n"][k]) * size[k])) for k in range(3)]
        t_enter, t_exit, enter_axis = F(-np.inf), F(np.inf), -1
        for k in range(3):  # Slab
            if abs(rd[k]) < F(1e-12):
                if ro[k] < mn[k] or ro[k] > mx[k]:
                    return None
                continue
            inv = F(F(1) / rd[k])
            t0, t1 = F(F(mn[k] - ro[k]) * inv), F(F(mx[k] - ro[k]) * inv)
            if t0 > t1:
                t0, t1 = t1, t0
            if t0 > t_enter:
                t_enter, enter_axis = t0, k
            if t1 < t_exit:
                t_exit = t1
            if not t_exit >= t_enter:
                return None
        if not t_exit >= max_f(F(0), t_enter):
            return None
        t = t_enter
        if t < t_min:
            t = t_min
        if t > t_max or t > t_exit:
            return None
        t = F(t + F(1e-6))
        p = [F(ro[k] + F(rd[k] * t)) for k in range(3)]
        idx = [min(max(int(np.floor(F(F(p[k] - mn[k]) / size[k]))), 0), g["n"][k] - 1) for k in range(3)]
        step = [1 if rd[k] > 0 else (-1 if rd[k] < 0 else 0) for k in range(3)]
        inv_d = [F(0) if step[k] == 0 else F(F(1) / rd[k]) for k in range(3)]
        next_v = [F(mn[k] + (F(F(idx[k] + 1) * size[k]) if step[k] > 0 else F(F(idx[k]) * size[k]))) for k in range(3)]
        t_max_a = [F(np.inf) if step[k] == 0 else F(F(next_v[k] - ro[k]) * inv_d[k]) for k in range(3)]
        t_delta = [F(np.inf) if step[k] == 0 else abs(F(size[k] * inv_d[k])) for k in range(3)]
        if enter_axis < 0:
            last_axis = 0 if (t_max_a[0] <= t_max_a[1] and t_max_a[0] <= t_max_a[2]) else (1 if t_max_a[1] <= t_max_a[2] else 2)
        else:
            last_axis = enter_axis
        wire_max2 = F(-1) if g["wire_max"] <= 0 else F(g["wire_max"] * g["wire_max"])
        dir_len2 = F(F(F(rd[0] * rd[0]) + F(rd[1] * rd[1])) + F(rd[2] * rd[2]))
        nbx, nby = (nx + 7) >> 3, (ny + 7) >> 3
        while t <= t_exit and t <= t_max:
            ix, iy, iz = idx
            if 0 <= ix < nx and 0 <= iy < ny and 0 <= iz < nz:
                lx, ly, lz = ix & 7, iy & 7, iz & 7
                morton = ((lx & 1) << 0) | ((ly & 1) << 1) | ((lz & 1) << 2) | ((lx & 2) << 2) | ((ly & 2) << 3) | ((lz & 2) << 4) | ((lx & 4) << 4) | ((ly & 4) << 5) | ((lz & 4) << 6)
                at = ((((iz >> 3) * nby) + (iy >> 3)) * nbx + (ix >> 3)) * 512 + morton
                mat_id = int(g["mat"][at])
                if mat_id > 0:
                    meta_id = int(g["meta"][at])
                    axis = last_axis  # never < 0 here: resolved above
                    hit_t = max_f(t, t_min)
                    n = v3(0, 0, 0)
                    n[axis] = F(-1) if step[axis] > 0 else F(1)
                    hp = v3(*(F(ro[k] + F(rd[k] * hit_t)) for k in range(3)))  # Ray.At
                    within = False
                    if g["wire"] and wire_max2 >= 0:
                        within = F(F(hit_t * hit_t) * dir_len2) <= wire_max2
                    mi = g["default"] if mat_id >= g["n_ids"] else int(g["palette"][mat_id * g["levels"] + min(max(meta_id, 0), g["levels"] - 1)])
                    m = self.mats[mi]
                    albedo = v3(*m.albedo)
                    if g["wire"] and within:  # IsWireOnFace: binary32 products widened to binary64
                        lo = [float(F(mn[k] + F(F(idx[k]) * size[k]))) for k in range(3)]
                        hi = [lo[k] + float(size[k]) for k in range(3)]
                        edge = lambda k: min(max(float(hp[k]) - lo[k], 0.0), max(hi[k] - float(hp[k]), 0.0))
                        a, b = [k for k in range(3) if k != axis]
                        w = float(F(g["wire_frac"] * min_f(size[a], size[b])))
                        if edge(a) <= w or edge(b) <= w:
                            albedo = v3(0, 0, 0)  # WireColor
                    mat = dict(albedo=albedo, refl=F(m.reflectivity), emission=v3(*m.emission), transparency=F(m.transparency), tint=v3(*m.transmission), ior=F(m.ior))
                    return dict(t=hit_t, P=hp, N=n, mat=mat, sub=ix + nx * (iy + ny * iz))
            if t_max_a[0] <= t_max_a[1] and t_max_a[0] <= t_max_a[2]:
                k = 0
            elif t_max_a[1] <= t_max_a[2]:
                k = 1
            else:
                k = 2
            idx[k] += step[k]
            t = t_max_a[k]
            t_max_a[k] = F(t_max_a[k] + t_delta[k])
            last_axis = k
            if not (0 <= idx[0] < nx and 0 <= idx[1] < ny and 0 <= idx[2] < nz):
                break
        return None

    # ---- Box.Hit (BoundedObjects.cs:78-115): six rects in a fixed order, the accepted hit shrinks `closest`
    def box_hit_obj(self, o, ro, rd, t_min, t_max):
        mnx, mny, mnz, mxx, mxy, mxz = (F(v) for v in o.p[0:6])
        faces = [(2, 0, 1, mnx, mxx, mny, mxy, mxz), (2, 0, 1, mnx, mxx, mny, mxy, mnz), (1, 0, 2, mnx, mxx, mnz, mxz, mxy),
                 (1, 0, 2, mnx, mxx, mnz, mxz, mny), (0, 1, 2, mny, mxy, mnz, mxz, mxx), (0, 1, 2, mny, mxy, mnz, mxz, mnx)]
        best, closest = None, t_max
        for i, (k, a, b, a0, a1, b0, b1, c) in enumerate(faces):
            tmp = self.rect_hit(o, k, a, b, a0, a1, b0, b1, c, ro, rd, t_min, closest)
            if tmp is not None:
                closest, best = tmp["t"], dict(tmp, sub=i)
        return best

    def object_hit(self, obj_id, ro, rd, t_min, t_max):
        o = self.objs[obj_id]
        q = [F(v) for v in o.p[0:5]]
        if o.kind == 0:
            return self.sphere_hit(o, ro, rd, t_min, t_max)
        if o.kind == 1:
            return self.plane_hit(o, ro, rd, t_min, t_max)
        if o.kind == 3:  # XYRect(x0, x1, y0, y1, z)
            return self.rect_hit(o, 2, 0, 1, q[0], q[1], q[2], q[3], q[4], ro, rd, t_min, t_max)
        if o.kind == 4:  # XZRect(x0, x1, z0, z1, y)
            return self.rect_hit(o, 1, 0, 2, q[0], q[1], q[2], q[3], q[4], ro, rd, t_min, t_max)
        if o.kind == 5:  # YZRect(y0, y1, z0, z1, x)
            return self.rect_hit(o, 0, 1, 2, q[0], q[1], q[2], q[3], q[4], ro, rd, t_min, t_max)
        if o.kind == 2:
            return self.disk_hit(o, ro, rd, t_min, t_max)
        if o.kind == 6:
            return self.box_hit_obj(o, ro, rd, t_min, t_max)
        if o.kind == 7:
            return self.cylinder_hit(o, ro, rd, t_min, t_max)
        if o.kind == 8:
            return self.triangle_hit(o, ro, rd, t_min, t_max)
        if o.kind == 9:
            return self.mesh_hit(o, ro, rd, t_min, t_max)
        if o.kind == 10:
            return self.volume_hit(o, ro, rd, t_min, t_max)
        raise NotImplementedError(f"object kind {o.kind} is outside this transcription")

    # ---- BVH.BoxHitFast BVH.cs:201-236
    @staticmethod
    def box_hit(box, ro, inv, t_min, t_max):
        ent, ext = [], []
        for k in range(3):
            e, x = F(F(box[k] - ro[k]) * inv[k]), F(F(box[3 + k] - ro[k]) * inv[k])
            if e > x:
                e, x = x, e
            ent.append(e)
            ext.append(x)
        t_enter = max_f(ent[0], max_f(ent[1], ent[2]))
        t_exit = min_f(ext[0], min_f(ext[1], ext[2]))
        if t_enter < t_min:
            t_enter = t_min
        if t_exit > t_max:
            t_exit = t_max
        return bool(t_exit >= t_enter), t_enter

    # ---- Scene.Hit -> BVH.Hit BVH.cs:99-198
    def scene_hit(self, ro, rd, t_min, t_max):
        self.rays += 1
        tr = self.tree
        if tr["root"] < 0:
            return None
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = [F(F(1) / rd[0]), F(F(1) / rd[1]), F(F(1) / rd[2])]
            closest, best = t_max, None
            stack = [tr["root"]]
            while stack:
                ni = stack.pop()
                hit, _ = self.box_hit(tr["boxes"][ni], ro, inv, t_min, closest)
                if not hit:
                    continue
                left, right, start, count = (int(v) for v in tr["lrsc"][ni])
                if count > 0:
                    for i in range(count):
                        obj_id = int(tr["leaf"][start + i])
                        tmp = self.object_hit(obj_id, ro, rd, t_min, closest)
                        if tmp is not None:
                            closest, best = tmp["t"], dict(tmp, obj=obj_id, sub=tmp.get("sub", 0))
                else:
                    hit_l = hit_r = False
                    l_near = r_near = F(0)
                    if left >= 0:
                        hit_l, l_near = self.box_hit(tr["boxes"][left], ro, inv, t_min, closest)
                    if right >= 0:
                        hit_r, r_near = self.box_hit(tr["boxes"][right], ro, inv, t_min, closest)
                    if hit_l and hit_r:
                        if l_near < r_near:
                            stack += [right, left]
                        else:
                            stack += [left, right]
                    elif hit_l:
                        stack.append(left)
                    elif hit_r:
                        stack.append(right)
        return best

    # ---- OrenNayarBRDF :810-831
    @staticmethod
    def oren_nayar(albedo, n, wo, wi, sigma):
        cos_i, cos_o = max_f(F(0), dot(n, wi)), max_f(F(0), dot(n, wo))
        if cos_i <= 0 or cos_o <= 0:
            return v3(0, 0, 0)
        sin_i = np.sqrt(max_f(F(0), F(F(1) - F(cos_i * cos_i))), dtype=F)
        sin_o = np.sqrt(max_f(F(0), F(F(1) - F(cos_o * cos_o))), dtype=F)
        proj_i = normalized(v3(*(F(wi[k] - F(n[k] * cos_i)) for k in range(3))))
        proj_o = normalized(v3(*(F(wo[k] - F(n[k] * cos_o)) for k in range(3))))
        cos_phi = max_f(F(0), dot(proj_i, proj_o))
        s2 = F(sigma * sigma)
        a = F(F(1) - F(s2 / F(F(2) * F(s2 + F(0.33)))))
        b = F(F(F(0.45) * s2) / F(s2 + F(0.09)))
        sin_alpha = max_f(sin_i, sin_o)
        tan_beta = min_f(F(sin_i / max_f(F(1e-6), cos_i)), F(sin_o / max_f(F(1e-6), cos_o)))
        on = F(a + F(F(F(b * cos_phi) * sin_alpha) * tan_beta))
        k = F(on * F(F(1) / PI))
        return v3(*(min(max(F(albedo[c] * k), F(0)), F(1)) for c in range(3)))

    # ---- CosineSampleHemisphere RaytraceSampler.cs:83-111
    def cosine_sample(self, n, rng):
        u1, u2 = rng.next_unit(), rng.next_unit()
        r = np.sqrt(u1, dtype=F)
        phi = F(F(6.2831853071795864769) * u2)
        x, y = F(r * self.cos(phi)), F(r * self.sin(phi))
        z = np.sqrt(F(F(1) - u1), dtype=F)
        w = n
        if w[2] < F(-0.999999):
            u, v = v3(0, -1, 0), v3(-1, 0, 0)
        else:
            a = F(F(1) / F(F(1) + w[2]))
            b = F(F(-w[0] * w[1]) * a)
            u = v3(F(1.0 - float(F(F(w[0] * w[0]) * a))), b, -w[0])
            v = v3(b, F(1.0 - float(F(F(w[1] * w[1]) * a))), -w[1])
        return v3(*(F(F(F(u[k] * x) + F(v[k] * y)) + F(w[k] * z)) for k in range(3)))

    # ---- ComputeTransmittanceToLight :757-798
    def transmittance(self, ro, rd, max_dist):
        if self.is_volume_scene:  # Scene.Occluded (Scene.cs:77-82): a nearest-hit query from 0.001, binary result
            return v3(0, 0, 0) if self.scene_hit(ro, rd, F(0.001), max_dist) is not None else v3(1, 1, 1)
        tr, tmin, counter = [F(1), F(1), F(1)], F(F(0) + EPS), 0
        while counter < MAX_REFR:
            block = self.scene_hit(ro, rd, tmin, max_dist)
            if block is None:
                break
            counter += 1
            m = block["mat"]
            if m["transparency"] <= 0:
                return v3(0, 0, 0)
            tr = [F(tr[k] * F(m["tint"][k] * m["transparency"])) for k in range(3)]
            if all(t <= F(1e-6) for t in tr):
                return v3(0, 0, 0)
            if block["t"] > max_dist:
                break
            tmin = F(block["t"] + EPS)
        return v3(*tr)

    # ---- MakeJitteredRay :419-437
    def make_ray(self, cam, yaw, pitch, fov, aspect, px, py, w, h, jrx, jry, frame_idx):
        def blue(channel):  # RaytraceSampler.cs:27-34
            base = F(F(F(BLUE[py & 7][px & 7]) + F(0.5)) * F(1.0 / 64.0))
            rot = frac(F(F(frame_idx + 1) * (F(0.7548776662466927) if channel == 0 else F(0.5698402909980532))))
            return frac(F(base + rot))
        jx, jy = F(frac(F(blue(0) + jrx)) - F(0.5)), F(frac(F(blue(1) + jry)) - F(0.5))
        u = F(F(F(F(F(F(px) + F(0.5)) + jx) / F(w)) * F(2)) - F(1))
        v = F(F(1) - F(F(F(F(F(py) + F(0.5)) + jy) / F(h)) * F(2)))
        fov_rad = F(fov * F(MATHF_PI / F(180)))
        half_h = self.tan(F(F(0.5) * fov_rad))
        half_w = F(half_h * aspect)
        cp = self.cos(pitch)
        fwd = normalized(v3(F(self.sin(yaw) * cp), self.sin(pitch), F(-self.cos(yaw) * cp)))
        right = normalized(cross(fwd, v3(0, 1, 0)))
        up = normalized(cross(right, fwd))
        su, sv = F(u * half_w), F(v * half_h)
        d = normalized(v3(*(F(F(fwd[k] + F(right[k] * su)) + F(up[k] * sv)) for k in range(3))))
        return cam, normalized(d)  # new Ray(...) normalises again (Ray.cs:8-12)

    # ---- TraceFull :448-620
    def trace_full(self, ro, rd, rng):
        stack = [dict(ro=ro, rd=rd, beta=v3(1, 1, 1), mirror=0, diffuse=0, primary=True)]
        radiance = v3(0, 0, 0)
        primary_hit, is_sky, gbuf_valid = False, False, False
        g = dict(albedo=v3(0, 0, 0), normal=v3(0, 0, 0), depth=FLT_MAX, obj=-1, sub=-1)
        sigma = F(SIGMA_DEG * F(MATHF_PI / F(180)))
        add = lambda acc, beta, c: v3(*(F(acc[k] + F(beta[k] * c[k])) for k in range(3)))
        while stack:
            item = stack.pop()
            ro, rd, beta, mirror, diffuse = item["ro"], item["rd"], item["beta"], item["mirror"], item["diffuse"]
            while True:
                rec = self.scene_hit(ro, rd, F(0.001), FLT_MAX)
                if rec is None:
                    tbg = F(F(0.5) * F(rd[1] + F(1)))
                    sky = v3(*(F(F(self.bg_bottom[k] * F(F(1) - tbg)) + F(self.bg_top[k] * tbg)) for k in range(3)))  # Lerp :805-808
                    if item["primary"] and not primary_hit:
                        is_sky = True
                        if not gbuf_valid:
                            g = dict(albedo=v3(0, 0, 0), normal=v3(0, 0, 0), depth=FLT_MAX, obj=-1, sub=-1)
                            gbuf_valid = True
                    radiance = add(radiance, beta, sky)
                    break
                m = rec["mat"]
                albedo = self.sample_albedo(m, rec.get("U", F(0)), rec.get("V", F(0)))
                if item["primary"]:
                    primary_hit, is_sky = True, False
                    if not gbuf_valid:
                        g = dict(albedo=albedo, normal=rec["N"], depth=rec["t"], obj=rec["obj"], sub=rec["sub"])
                        gbuf_valid = True
                    item["primary"] = False
                if m["emission"].any():
                    radiance = add(radiance, beta, m["emission"])
                base = albedo
                if m["transparency"] > 0:  # :506-558: Fresnel split into two deferred work items (reflection pushed first, refraction popped first)
                    if mirror >= MAX_MIRROR:
                        break
                    n, wo = rec["N"], rd
                    front = dot(n, wo) < 0
                    nl = n if front else v3(*(F(n[k] * F(-1)) for k in range(3)))
                    eta_i, eta_t = (F(1), m["ior"]) if front else (m["ior"], F(1))
                    eta = F(eta_i / eta_t)
                    k2 = F(F(2) * dot(wo, nl))
                    refl_dir = normalized(v3(*(F(wo[k] - F(nl[k] * k2)) for k in range(3))))
                    cosi = -max_f(F(-1), min_f(F(1), dot(wo, nl)))                       # Refract :737-748
                    kk = F(F(1) - F(F(eta * eta) * F(F(1) - F(cosi * cosi))))
                    has_refract = not kk < 0
                    refr_dir = v3(0, 0, 0)
                    if has_refract:
                        c2 = F(F(eta * cosi) - np.sqrt(kk, dtype=F))
                        refr_dir = v3(*(F(F(wo[k] * eta) + F(nl[k] * c2)) for k in range(3)))
                    cos_theta = abs(dot(nl, v3(*(F(wo[k] * F(-1)) for k in range(3)))))
                    r0 = F(F(eta_i - eta_t) / F(eta_i + eta_t))                          # FresnelSchlick :750-755
                    r0 = F(r0 * r0)
                    R = F(r0 + F(F(F(1) - r0) * self.pow(F(F(1) - cos_theta), F(5))))
                    tr_c = min(max(m["transparency"], F(0)), F(1))
                    T = F(F(F(1) - R) * tr_c) if has_refract else F(0)
                    R = min(max(F(R + F(m["refl"] * F(F(1) - R))), F(0)), F(1))
                    if R > 0 and len(stack) < 16:
                        o2 = v3(*(F(rec["P"][k] + F(nl[k] * EPS)) for k in range(3)))
                        stack.append(dict(ro=o2, rd=normalized(refl_dir), beta=v3(*(F(F(beta[k] * base[k]) * R) for k in range(3))), mirror=mirror + 1, diffuse=diffuse, primary=False))
                    if T > 0 and len(stack) < 16:
                        o2 = v3(*(F(rec["P"][k] - F(nl[k] * EPS)) for k in range(3)))
                        stack.append(dict(ro=o2, rd=normalized(normalized(refr_dir)), beta=v3(*(F(F(beta[k] * m["tint"][k]) * T) for k in range(3))), mirror=mirror + 1, diffuse=diffuse, primary=False))
                    break
                if m["refl"] >= MIRROR_THRESHOLD:
                    if mirror >= MAX_MIRROR:
                        break
                    k2 = F(F(2) * dot(rd, rec["N"]))
                    refl = normalized(v3(*(F(rd[k] - F(rec["N"][k] * k2)) for k in range(3))))  # Reflect :800-803
                    ro = v3(*(F(rec["P"][k] + F(rec["N"][k] * EPS)) for k in range(3)))
                    rd = normalized(refl)
                    beta = v3(*(F(beta[k] * base[k]) for k in range(3)))
                    mirror += 1
                    continue
                if self.amb_i > 0:
                    amb = v3(*(F(F(self.amb_c[k] * self.amb_i) * base[k]) for k in range(3)))
                    radiance = add(radiance, beta, amb)
                wo = normalized(v3(*(F(rd[k] * F(-1)) for k in range(3))))
                for lpos, lcol, lint in self.lights:
                    to_l = v3(*(F(lpos[k] - rec["P"][k]) for k in range(3)))
                    dist2 = dot(to_l, to_l)
                    dist = np.sqrt(dist2, dtype=F)
                    ldir = vdiv(to_l, dist)
                    n_dot_l = max_f(F(0), dot(rec["N"], ldir))
                    if n_dot_l <= 0:
                        continue
                    so = v3(*(F(rec["P"][k] + F(rec["N"][k] * EPS)) for k in range(3)))
                    trans = self.transmittance(so, normalized(ldir), F(dist - EPS))
                    if all(t <= F(1e-6) for t in trans):
                        continue
                    atten = F(lint / dist2)
                    f_d = self.oren_nayar(base, rec["N"], wo, ldir, sigma)
                    contrib = v3(*(F(F(F(f_d[k] * n_dot_l) * F(lcol[k] * atten)) * trans[k]) for k in range(3)))
                    radiance = add(radiance, beta, contrib)
                if diffuse < DIFFUSE_BOUNCES:
                    bounce = self.cosine_sample(rec["N"], rng)
                    f_on = self.oren_nayar(base, rec["N"], wo, bounce, sigma)
                    ro = v3(*(F(rec["P"][k] + F(rec["N"][k] * EPS)) for k in range(3)))
                    rd = normalized(bounce)
                    beta = v3(*(F(beta[k] * F(f_on[k] * PI)) for k in range(3)))
                    diffuse += 1
                    continue
                break
        return radiance, is_sky, g


@pytest.mark.parametrize("scene_name,fb_w,fb_h,ss", [("test", 10, 4, 2), ("mirror_spheres", 10, 4, 2), ("cornell", 9, 4, 2), ("boxes", 5, 9, 1), ("knot:12x5", 8, 3, 2),
                                                     ("volume_grid_test", 10, 4, 2), ("voxel_world:32x32", 8, 4, 2), ("cylinders_disks_triangles", 10, 4, 2),
                                                     ("texture_gallery", 14, 5, 2), ("voxel_island:96x128", 8, 4, 2), ("all_meshes:40x10", 16, 3, 2), ("museum", 14, 5, 2), ("museum-diorama", 12, 4, 2), ("museum-texture", 12, 4, 2)])
def test_trace_stage_matches_a_literal_python_transcription(scene_name, fb_w, fb_h, ss):
    lib = load_oracle()
    lib.yo_set_math_mode(0)
    scene = api.HostScene(scene_name.split("-")[0])
    o = Oracle(scene, fb_w, fb_h, ss)
    lt = LiteralTracer(scene, lib)
    pos, yaw, pitch, fov = scene.default_camera()
    mesh_in_view = scene.n_meshes > 0
    if scene_name.startswith("knot"):  # the default pose of the mesh scenes looks away from the mesh
        pos, yaw, pitch = api.BENCH_POSE
        o.set_camera(pos, yaw, pitch)
    if scene_name == "museum":         # from the entrance the mesh gallery (x = 9, z = -40) is out of sight: stand in front of it
        pos, yaw, pitch = (9.0, 3.0, -35.5), 0.0, -0.35
        o.set_camera(pos, yaw, pitch)
    if scene_name == "museum-texture": # the textured sphere (Sphere.Hit leaves U = V = 0: one texel) in front of the textured end wall (a Plane: likewise)
        scene_name, mesh_in_view = "museum", False
        pos, yaw, pitch = (-1.6, 1.0, -7.5), 0.0, -0.1
        o.set_camera(pos, yaw, pitch)
    if scene_name == "museum-diorama": # voxel diorama B: 14 x 7 x 14 cells of 0.45 -- partial 8^3 bricks -- with the teapot on its stand
        scene_name = "museum"
        pos, yaw, pitch = (5.5, 2.6, -84.5), 1.45, -0.3
        o.set_camera(pos, yaw, pitch)
    cam, yaw, pitch, fov = v3(*pos), F(yaw), F(pitch), F(fov)
    w, h = fb_w * ss, fb_h * 2 * ss
    aspect = F(F(w) / F(h))
    saw_mirror, saw_mesh = False, False
    with np.errstate(over="ignore"):
        for frame in (1, 2):
            o.render_frame(threads=2)
            hdr, als, nd, prim = o.debug_read(api.DBG_HDR), o.debug_read(api.DBG_ALBEDO_SKY), o.debug_read(api.DBG_NORMAL_DEPTH), o.debug_read(api.DBG_PRIM_ID)
            raw_n, rays_ref = o.raw_normal(), o.debug_read(api.DBG_RAYS)
            frame_idx = frame & 0x7FFFFFFF
            jrx, jry = frac(F(F(frame_idx + 1) * F(0.61803398875))), frac(F(F(frame_idx + 1) * F(0.38196601125)))  # :178-179
            lt.rays = 0
            for py in range(h):
                for px in range(w):
                    ro, rd = lt.make_ray(cam, yaw, pitch, fov, aspect, px, py, w, h, jrx, jry, frame_idx)
                    assert np.array_equal(np.concatenate([ro, rd]).view(np.uint32), rays_ref[py, px].view(np.uint32)), (frame, px, py, "ray")
                    rad, is_sky, g = lt.trace_full(ro, rd, Rng(per_frame_seed(px, py, frame)))
                    where = (scene_name, frame, px, py)
                    assert np.array_equal(rad.view(np.uint32), hdr[py, px, :3].view(np.uint32)), where + ("radiance", rad, hdr[py, px, :3])
                    assert bool(als[py, px, 3]) == is_sky, where + ("sky",)
                    assert np.array_equal(g["albedo"].view(np.uint32), als[py, px, :3].view(np.uint32)), where + ("albedo",)
                    assert np.array_equal(g["normal"].view(np.uint32), raw_n[py, px].view(np.uint32)), where + ("normal",)
                    assert F(g["depth"]).view(np.uint32) == nd[py, px, 3].view(np.uint32), where + ("depth",)
                    assert (g["obj"], g["sub"]) == tuple(prim[py, px]), where + ("primary ids",)
                    saw_mesh |= g["obj"] >= 0 and lt.objs[g["obj"]].kind == 9
                    if g["obj"] >= 0 and lt.objs[g["obj"]].kind != 9 and F(lt.objs[g["obj"]].reflectivity if lt.objs[g["obj"]].override_sr else lt.mats[lt.objs[g["obj"]].mat_a].reflectivity) >= MIRROR_THRESHOLD:
                        saw_mirror = True
            assert lt.rays == o.stats()["rays"], (scene_name, frame, "Scene.Hit invocations")
    assert saw_mesh == mesh_in_view, "the mesh must be in view, or MeshBVH.Hit is not exercised on primary rays"
    if scene_name == "test":
        assert saw_mirror, "the mirror sphere must be in view, or the mirror branch is not exercised"
    o.close()
    scene.close()


@pytest.mark.parametrize("scene_name", ["cornell", "cylinders_disks_triangles", "volume_grid_test", "texture_gallery", "museum", "all_meshes:40x10", "voxel_island:96x128", "entities_demo"])
def test_trace_stage_matches_the_transcription_from_random_poses(scene_name):
    """The same comparison from camera poses drawn at random (seeded): inside and outside the geometry, looking up, down and along
    surfaces, so that grazing hits, back faces, the inside of boxes, misses of every slab and total internal reflection occur."""
    lib = load_oracle()
    lib.yo_set_math_mode(0)
    scene = api.HostScene(scene_name)
    fb_w, fb_h, ss = 6, 3, 1
    o = Oracle(scene, fb_w, fb_h, ss)
    lt = LiteralTracer(scene, lib)
    w, h = fb_w * ss, fb_h * 2 * ss
    aspect = F(F(w) / F(h))
    rng = np.random.default_rng(len(scene_name) * 7919)
    frame = 0
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for _ in range(5):
            pos = (float(rng.uniform(-2.5, 2.5)), float(rng.uniform(0.05, 3.0)), float(rng.uniform(-5.0, 1.0)))
            yaw, pitch = float(F(rng.uniform(-3.1, 3.1))), float(F(rng.uniform(-1.2, 1.2)))
            o.set_camera(pos, yaw, pitch)
            o.render_frame(threads=2)
            frame += 1
            hdr, als, nd, prim = o.debug_read(api.DBG_HDR), o.debug_read(api.DBG_ALBEDO_SKY), o.debug_read(api.DBG_NORMAL_DEPTH), o.debug_read(api.DBG_PRIM_ID)
            frame_idx = frame & 0x7FFFFFFF
            jrx, jry = frac(F(F(frame_idx + 1) * F(0.61803398875))), frac(F(F(frame_idx + 1) * F(0.38196601125)))
            lt.rays = 0
            for py in range(h):
                for px in range(w):
                    ro, rd = lt.make_ray(v3(*pos), F(yaw), F(pitch), F(45.0), aspect, px, py, w, h, jrx, jry, frame_idx)
                    rad, is_sky, g = lt.trace_full(ro, rd, Rng(per_frame_seed(px, py, frame)))
                    where = (scene_name, pos, yaw, pitch, px, py)
                    assert np.array_equal(rad.view(np.uint32), hdr[py, px, :3].view(np.uint32)), where + ("radiance",)
                    assert bool(als[py, px, 3]) == is_sky and (g["obj"], g["sub"]) == tuple(prim[py, px]), where + ("sky / ids",)
                    assert F(g["depth"]).view(np.uint32) == nd[py, px, 3].view(np.uint32), where + ("depth",)
            assert lt.rays == o.stats()["rays"], (scene_name, pos, "Scene.Hit invocations")
    o.close()
    scene.close()
