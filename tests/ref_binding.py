"""ctypes binding of oracle/_ref/libycge_ref.so: the reference's own C# sources, rewritten into C++ syntactically at build time
(oracle/ref_transpile.py) -- test infrastructure.  RefRenderer drives the verbatim head and tail of TryFlipAndBlit: a whole frame of
the reference, from the flat scene description the product consumes."""
import ctypes as C
import os

import numpy as np

from yetanotherconsolegameengine_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libycge_ref.so")


def available():
    return os.path.exists(REF_SO)


def load():
    lib = C.CDLL(REF_SO)
    vp = C.c_void_p
    lib.ref_renderer_create.restype = vp
    lib.ref_renderer_create.argtypes = [C.c_int] * 4
    lib.ref_renderer_destroy.argtypes = [vp]
    lib.ref_post_frame.argtypes = [vp] * 6 + [C.c_int] + [vp] * 10
    lib.ref_trace_create.restype = vp
    lib.ref_trace_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int] + [vp] * 10 + [C.c_int] + [vp] * 4 + [C.c_int, vp, C.c_int, vp, C.c_int]
    lib.ref_trace_destroy.argtypes = [vp]
    lib.ref_trace_frame.argtypes = [vp, vp, C.c_float, C.c_float] + [vp] * 6
    lib.ref_set_threads.argtypes = [C.c_int]
    lib.ref_taa_should_reset.argtypes = [vp, vp, C.c_float, C.c_float]
    lib.ref_taa_commit_camera.argtypes = [vp, vp, C.c_float, C.c_float]
    lib.ref_taa_resize.argtypes = [vp]
    lib.ref_renderer_resize.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.ref_trace_resize.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    return lib


def P(a):
    return C.c_void_p(a.ctypes.data)


def mat13(m):
    """one material row of the harness: 16 floats (ref_harness.cpp MS): the 13 scalars of Material + texture id (-1: none), weight, uv scale"""
    return list(m.albedo) + [m.specular, m.reflectivity] + list(m.emission) + [m.transparency, m.ior] + list(m.transmission) + [float(m.tex_id), m.tex_weight, m.uv_scale]


MAT_PAD = [[0.0] * 13 + [-1.0, 0.0, 0.0]]


def set_textures(lib, scene):
    """hands the scene's static images (Texture.cs:81-90 layout) to the harness; the arrays must stay alive while objects are created"""
    tex = [np.ascontiguousarray(scene.texture(i), np.uint32) for i in range(scene.n_textures)]
    w = np.array([t.shape[1] for t in tex] + [0], np.int32)
    h = np.array([t.shape[0] for t in tex] + [0], np.int32)
    ptrs = (C.c_void_p * max(1, len(tex)))(*[t.ctypes.data for t in tex])
    lib.ref_set_textures.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_set_textures(len(tex), P(w), P(h), ptrs)
    return tex


class RefRenderer:
    """One frame = TraceStage (RaytraceRenderer.cs:159, :175-216) + PostTail (:218-264) of the transpiled reference."""

    def __init__(self, scene: api.HostScene, fb_w: int, fb_h: int, ss: int, threads: int = 1):
        self.lib = load()
        self.lib.ref_set_threads(threads)
        self.s, self.fb_w, self.fb_h, self.ss = scene, fb_w, fb_h, ss
        f = scene.flat.contents
        n = f.n_objects
        assert all(f.objects[k].kind <= 10 for k in range(n))
        self.keep = []  # arrays the native side points into
        kind = np.array([f.objects[k].kind for k in range(n)], np.int32)
        p12 = np.array([list(f.objects[k].p) for k in range(n)], np.float32)
        ma = np.array([mat13(f.materials[max(0, f.objects[k].mat_a)]) for k in range(n)], np.float32)
        mb = np.array([mat13(f.materials[max(0, f.objects[k].mat_b)]) for k in range(n)], np.float32)
        cs = np.array([f.objects[k].checker_scale for k in range(n)], np.float32)
        sp = np.array([f.objects[k].specular for k in range(n)], np.float32)
        rf = np.array([f.objects[k].reflectivity for k in range(n)], np.float32)
        meshes = [np.ascontiguousarray(scene.mesh_triangles(i), np.float32) for i in range(scene.n_meshes)]
        mesh_n = np.array([len(m) for m in meshes] + [0], np.int32)
        mesh_ptrs = (C.c_void_p * max(1, len(meshes)))(*[m.ctypes.data for m in meshes])
        mesh_mat = np.array([mat13(scene.mesh(i).contents.material) for i in range(scene.n_meshes)] + MAT_PAD, np.float32)
        lights = np.array([list(f.lights[i].pos) + list(f.lights[i].color) + [f.lights[i].intensity] for i in range(f.n_lights)] + [[0] * 7], np.float32)
        top, bot = np.array(list(f.bg_top), np.float32), np.array(list(f.bg_bottom), np.float32)
        amb = np.array(list(f.ambient_color) + [f.ambient_intensity], np.float32)
        vols = (api.Volume * max(1, scene.n_volumes))(*[scene.volume(i).contents for i in range(scene.n_volumes)])
        all_mats = np.array([mat13(f.materials[i]) for i in range(f.n_materials)] + MAT_PAD, np.float32)
        self.keep += [meshes, vols]
        set_textures(self.lib, scene)  # copied by the harness; materials with a texture id point at them
        fov = scene.default_camera()[3]
        self.trace = self.lib.ref_trace_create(fb_w, fb_h, ss, fov, n, P(kind), P(p12), P(ma), P(mb), P(cs), P(sp), P(rf), P(mesh_n), mesh_ptrs, P(mesh_mat), f.n_lights, P(lights),
                                               P(top), P(bot), P(amb), scene.n_volumes, C.cast(vols, C.c_void_p), f.n_materials, P(all_mats), f.is_volume_scene)
        assert self.trace, "the transpiled reference could not build the scene"
        self.post = self.lib.ref_renderer_create(fb_w, fb_h, ss, 3)
        self.cam = scene.default_camera()[:3]
        self.frame = 0

    def set_camera(self, pos, yaw, pitch):
        self.cam = (pos, yaw, pitch)

    def render_frame(self, reset_history=None):
        """-> dict of planes (rays, hdr, albedo, normal, depth, sky, taa, den), ae_exposure and the cell fields.  `reset_history`: None =
        the reference's own decision, taa.ShouldResetHistory(camera) of the transpiled TemporalAA.cs (TryFlipAndBlit :171; the camera is
        committed after the frame, :266); True / False overrides it (`|| scene.HasDynamicTextures`)."""
        W, H, fb_w, fb_h = self.fb_w * self.ss, self.fb_h * 2 * self.ss, self.fb_w, self.fb_h
        o = dict(rays=np.empty((H, W, 6), np.float32), hdr=np.empty((H, W, 3), np.float32), albedo=np.empty((H, W, 3), np.float32), normal=np.empty((H, W, 3), np.float32),
                 depth=np.empty((H, W), np.float32), sky=np.empty((H, W), np.uint8), taa=np.empty((H, W, 3), np.float32), den=np.empty((H, W, 3), np.float32),
                 expo=np.empty(2, np.float32), glyph=np.empty((fb_h, fb_w), np.uint16), fg16=np.empty((fb_h, fb_w), np.uint8), bg16=np.empty((fb_h, fb_w), np.uint8),
                 fg_ansi=np.empty((fb_h, fb_w), np.uint8), bg_ansi=np.empty((fb_h, fb_w), np.uint8), fg=np.empty((fb_h, fb_w, 3), np.float32), bg=np.empty((fb_h, fb_w, 3), np.float32))
        c3 = np.array(self.cam[0], np.float32)
        rc = self.lib.ref_trace_frame(self.trace, P(c3), np.float32(self.cam[1]), np.float32(self.cam[2]), P(o["rays"]), P(o["hdr"]), P(o["albedo"]), P(o["normal"]), P(o["depth"]), P(o["sky"]))
        assert rc == 0
        self.frame += 1
        reset = bool(self.lib.ref_taa_should_reset(self.post, P(c3), np.float32(self.cam[1]), np.float32(self.cam[2]))) if reset_history is None else reset_history
        self.last_reset = reset
        rc = self.lib.ref_post_frame(self.post, P(o["hdr"]), P(o["albedo"]), P(o["normal"]), P(o["depth"]), P(o["sky"]), 1 if reset else 0, P(o["taa"]), P(o["den"]), P(o["expo"]),
                                     P(o["glyph"]), P(o["fg16"]), P(o["bg16"]), P(o["fg_ansi"]), P(o["bg_ansi"]), P(o["fg"]), P(o["bg"]))
        assert rc == 0
        self.lib.ref_taa_commit_camera(self.post, P(c3), np.float32(self.cam[1]), np.float32(self.cam[2]))
        return o

    def resize(self, fb_w, fb_h, ss):
        """RaytraceRenderer.Resize (:110-138), the reference's own text, on both halves of the harness."""
        assert self.lib.ref_trace_resize(self.trace, fb_w, fb_h, ss) == 0
        assert self.lib.ref_renderer_resize(self.post, fb_w, fb_h, ss) == 0
        self.fb_w, self.fb_h, self.ss = fb_w, fb_h, ss

    def close(self):
        if self.trace:
            self.lib.ref_trace_destroy(self.trace)
            self.lib.ref_renderer_destroy(self.post)
            self.trace = None
