"""ctypes binding of the CPU oracle (oracle/libycge_oracle.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The oracle consumes exactly the flat
scene description (include/ycge.h structs) that the host layer hands to the CUDA library: same inputs on both sides."""
import ctypes as C
import os
import subprocess

import numpy as np

from yetanotherconsolegameengine_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libycge_oracle.so")

_lib = None


def load_oracle():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            subprocess.check_call(["make"], cwd=ORACLE_DIR)
        o = C.CDLL(ORACLE_LIB)
        vp = C.c_void_p
        o.yo_create.argtypes = [C.POINTER(api.Config)]
        o.yo_create.restype = vp
        o.yo_destroy.argtypes = [vp]
        o.yo_destroy.restype = None
        o.yo_default_params.argtypes = [C.POINTER(api.Params)]
        o.yo_default_params.restype = None
        o.yo_resize.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        o.yo_mesh_upload_triangles.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        o.yo_mesh_upload_soa.argtypes = [vp, C.c_int, vp]
        o.yo_volume_upload.argtypes = [vp, C.c_int, vp]
        o.yo_scene_upload.argtypes = [vp, vp]
        o.yo_texture_upload.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        o.yo_texture_sample.argtypes = [C.c_int, C.c_int, vp, C.c_float, C.c_float, vp]
        o.yo_texture_sample.restype = None
        o.yo_lights_update.argtypes = [vp, C.c_int, vp]
        o.yo_globals_update.argtypes = [vp, vp, vp, vp, C.c_float]
        o.yo_set_camera.argtypes = [vp, vp, C.c_float, C.c_float]
        o.yo_set_fov.argtypes = [vp, C.c_float]
        o.yo_reset_history.argtypes = [vp]
        o.yo_render_frame.argtypes = [vp, vp, C.c_int, C.c_int]
        o.yo_get_stats.argtypes = [vp, C.POINTER(api.Stats)]
        o.yo_debug_read.argtypes = [vp, C.c_int, vp, C.c_size_t]
        o.yo_debug_raw_normal.argtypes = [vp, vp]
        o.yo_bvh_info.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        o.yo_bvh_read.argtypes = [vp, C.c_int, vp, vp, vp]
        o.yo_mesh_soa_read.argtypes = [vp, C.c_int, vp]
        o.yo_scene_hit.argtypes = [vp, C.c_int, vp, C.c_float, C.c_float, C.c_int, vp, vp, vp]
        o.yo_splitmix64.argtypes = [C.c_uint64]
        o.yo_splitmix64.restype = C.c_uint64
        o.yo_per_frame_seed.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_uint64]
        o.yo_per_frame_seed.restype = C.c_uint64
        o.yo_rng_draws.argtypes = [C.c_uint64, C.c_int, vp, vp]
        o.yo_rng_draws.restype = None
        o.yo_rng_cs_draws.argtypes = [C.c_uint64, C.c_int, vp]
        o.yo_rng_cs_draws.restype = None
        o.yo_blue_noise.argtypes = [C.c_int] * 4
        o.yo_blue_noise.restype = C.c_float
        o.yo_blue_noise_table.argtypes = [C.c_int, C.c_int]
        o.yo_morton3.argtypes = [C.c_int] * 3
        o.yo_volume_index_of.argtypes = [C.c_int] * 6
        o.yo_ansi256.argtypes = [C.c_float] * 3
        o.yo_linear_to_srgb8.argtypes = [C.c_double]
        o.yo_cube_level.argtypes = [C.c_int]
        o.yo_nearest16.argtypes = [C.c_float] * 3
        o.yo_tonemap.argtypes = [C.c_float] * 4 + [vp]
        o.yo_tonemap.restype = None
        o.yo_cosine_sample.argtypes = [C.c_float] * 3 + [C.c_uint64, vp]
        o.yo_cosine_sample.restype = None
        o.yo_dotnet_sort_floats.argtypes = [vp, vp, C.c_int]
        o.yo_dotnet_sort_floats.restype = None
        o.yo_math.argtypes = [C.c_int, C.c_float, C.c_float]
        o.yo_math.restype = C.c_float
        o.yo_set_math_mode.argtypes = [C.c_int]
        o.yo_set_math_mode.restype = None
        o.yo_set_sort_mode.argtypes = [C.c_int]
        o.yo_set_sort_mode.restype = None
        _lib = o
    return _lib


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class Oracle:
    """The CPU restatement driven with the same flat scene the CUDA library receives."""

    def __init__(self, scene, fb_w, fb_h, ss=1, params=None, use_host_trees=True, mesh_form="soa"):
        self.o = load_oracle()
        cfg = api.Config()
        cfg.fb_w, cfg.fb_h, cfg.ss = fb_w, fb_h, ss
        if params is None:
            self.o.yo_default_params(C.byref(cfg.params))
        else:
            cfg.params = params
        self.fb_w, self.fb_h, self.ss = fb_w, fb_h, max(1, ss)
        self.h = C.c_void_p(self.o.yo_create(C.byref(cfg)))
        self.scene = scene
        self.upload_scene(scene, use_host_trees, mesh_form)
        pos, yaw, pitch, fov = scene.default_camera()
        self.set_fov(fov)
        self.set_camera(pos, yaw, pitch)

    def upload_scene(self, scene, use_host_trees=True, mesh_form="soa"):
        for i in range(scene.n_meshes):
            if mesh_form == "soa":
                assert self.o.yo_mesh_upload_soa(self.h, i, scene.mesh(i)) == 0
            else:  # the oracle builds the MeshBVH itself from A,B,C
                tris = scene.mesh_triangles(i)
                m = scene.mesh(i).contents
                assert self.o.yo_mesh_upload_triangles(self.h, i, len(tris), _ptr(tris), C.byref(m.material)) == 0
        for i in range(scene.n_volumes):
            assert self.o.yo_volume_upload(self.h, i, scene.volume(i)) == 0
        for i in range(scene.n_textures):
            t = scene.texture(i)
            assert self.o.yo_texture_upload(self.h, i, t.shape[1], t.shape[0], _ptr(t)) == 0
        flat = scene.flat
        if use_host_trees:
            assert self.o.yo_scene_upload(self.h, flat) == 0
        else:  # the oracle builds the top-level BVH itself
            s2 = api.Scene()
            C.memmove(C.byref(s2), flat, C.sizeof(api.Scene))
            s2.bvh = None
            assert self.o.yo_scene_upload(self.h, C.byref(s2)) == 0

    def close(self):
        if getattr(self, "h", None):
            self.o.yo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_camera(self, pos, yaw, pitch):
        p = (C.c_float * 3)(*pos)
        self.o.yo_set_camera(self.h, p, yaw, pitch)

    def set_fov(self, fov):
        self.o.yo_set_fov(self.h, fov)

    def reset_history(self):
        self.o.yo_reset_history(self.h)

    def lights_update(self, lights):
        arr = (api.Light * len(lights))()
        for i, (pos, col, inten) in enumerate(lights):
            arr[i].pos[:] = pos
            arr[i].color[:] = col
            arr[i].intensity = inten
        assert self.o.yo_lights_update(self.h, len(lights), arr) == 0

    def globals_update(self, bg_top, bg_bottom, ambient_color, ambient_intensity):
        a, b, c = (C.c_float * 3)(*bg_top), (C.c_float * 3)(*bg_bottom), (C.c_float * 3)(*ambient_color)
        assert self.o.yo_globals_update(self.h, a, b, c, ambient_intensity) == 0

    def resize(self, fb_w, fb_h, ss):
        self.o.yo_resize(self.h, fb_w, fb_h, ss)
        self.fb_w, self.fb_h, self.ss = fb_w, fb_h, max(1, ss)

    def render_frame(self, threads=1, fast_post=False):
        out = np.empty((self.fb_h, self.fb_w), api.CELL_DTYPE)
        rc = self.o.yo_render_frame(self.h, _ptr(out), threads, 1 if fast_post else 0)
        assert rc == 0, rc
        return out

    def stats(self):
        s = api.Stats()
        self.o.yo_get_stats(self.h, C.byref(s))
        return s.as_dict()

    @property
    def hi_w(self):
        return self.fb_w * self.ss

    @property
    def hi_h(self):
        return self.fb_h * 2 * self.ss

    def debug_read(self, kind):
        if kind == api.DBG_RAYS:
            a = np.empty((self.hi_h, self.hi_w, 6), np.float32)
        elif kind == api.DBG_PRIM_ID:
            a = np.empty((self.hi_h, self.hi_w, 2), np.int32)
        elif kind == api.DBG_LOG_SAMPLES:
            step = max(2, self.ss * 2)
            a = np.empty(((self.hi_h + step - 1) // step, (self.hi_w + step - 1) // step), np.float32)
        else:
            a = np.empty((self.hi_h, self.hi_w, 4), np.float32)
        assert self.o.yo_debug_read(self.h, kind, _ptr(a), a.nbytes) == 0
        return a

    def raw_normal(self):
        a = np.empty((self.hi_h, self.hi_w, 3), np.float32)
        assert self.o.yo_debug_raw_normal(self.h, _ptr(a)) == 0
        return a

    def bvh_arrays(self, which=-1):
        n, root, nl = C.c_int(), C.c_int(), C.c_int()
        sf = C.c_uint64()
        assert self.o.yo_bvh_info(self.h, which, C.byref(n), C.byref(root), C.byref(nl), C.byref(sf)) == 0
        boxes = np.empty((n.value, 6), np.float32)
        lrsc = np.empty((n.value, 4), np.int32)
        leaf = np.empty((nl.value,), np.int32)
        assert self.o.yo_bvh_read(self.h, which, _ptr(boxes), _ptr(lrsc), _ptr(leaf)) == 0
        return dict(root=root.value, boxes=boxes, lrsc=lrsc, leaf=leaf, sort_fallbacks=sf.value)

    def scene_hit(self, rays6, tmin=0.001, tmax=3.4028234663852886e38, use_bvh=True):
        r = np.ascontiguousarray(rays6, np.float32)
        n = len(r)
        t = np.empty(n, np.float32)
        ids = np.empty((n, 2), np.int32)
        nrm = np.empty((n, 3), np.float32)
        self.o.yo_scene_hit(self.h, n, _ptr(r), tmin, tmax, 1 if use_bvh else 0, _ptr(t), _ptr(ids), _ptr(nrm))
        return t, ids, nrm
