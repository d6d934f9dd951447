"""ctypes bindings for the product libraries.

* ``libycge.so``       — the C ABI of ``include/ycge.h`` (CUDA, sm_100a).  No CPU path: every call fails loudly
                         (``YcgeError``) when no CUDA device is usable.
* ``libycge_host.so``  — the C++ mirror of the reference's C# host side (scene factories of ``BuildSceneTable()``,
                         ``MeshLoader``, ``Framebuffer``/``Chexel``, ``IConsoleRenderer``, ``ANSITerminalRenderer``).

The Python classes below mirror the reference-facing names (``RaytraceEntity.IConsoleRenderer``: SetCamera / SetFov /
TryFlipAndBlit / Resize, ConsoleGame/RaytraceEntity.cs:12-18) so that the parity tests read like calls into the
reference.  Nothing in this module imports or calls the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YCGE_LIB", os.path.join(_PKG_DIR, "libycge.so"))  # YCGE_LIB: development aid (kernel build variants)
HOST_LIB_PATH = os.path.join(_PKG_DIR, "libycge_host.so")


class YcgeError(RuntimeError):
    """Raised for any negative ycge_status (the reference throws InvalidOperationException / ArgumentException)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"ycge error {code}: {message}")
        self.code = code


# ---------------------------------------------------------------------------------------------- ABI structs (ycge.h)
class Material(C.Structure):
    _fields_ = [("albedo", C.c_float * 3), ("reflectivity", C.c_float), ("emission", C.c_float * 3), ("transparency", C.c_float),
                ("transmission", C.c_float * 3), ("ior", C.c_float), ("specular", C.c_float), ("tex_id", C.c_int32),
                ("tex_weight", C.c_float), ("uv_scale", C.c_float)]


class Object(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mat_a", C.c_int32), ("mat_b", C.c_int32), ("checker_scale", C.c_float), ("override_sr", C.c_int32),
                ("specular", C.c_float), ("reflectivity", C.c_float), ("ref_id", C.c_int32), ("p", C.c_float * 12)]


class Light(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("color", C.c_float * 3), ("intensity", C.c_float)]


class Bvh(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("root", C.c_int32), ("n_leaf_refs", C.c_int32)] + \
               [(n, C.POINTER(C.c_float)) for n in ("min_x", "min_y", "min_z", "max_x", "max_y", "max_z")] + \
               [(n, C.POINTER(C.c_int32)) for n in ("left", "right", "start", "count", "leaf_index")]


class Scene(C.Structure):
    _fields_ = [("bg_top", C.c_float * 3), ("bg_bottom", C.c_float * 3), ("ambient_color", C.c_float * 3), ("ambient_intensity", C.c_float),
                ("is_volume_scene", C.c_int32), ("n_lights", C.c_int32), ("lights", C.POINTER(Light)), ("n_materials", C.c_int32),
                ("materials", C.POINTER(Material)), ("n_objects", C.c_int32), ("objects", C.POINTER(Object)), ("bvh", C.POINTER(Bvh))]


class MeshSoa(C.Structure):
    _fields_ = [("n_tris", C.c_int32)] + \
               [(n, C.POINTER(C.c_float)) for n in ("ax", "ay", "az", "e1x", "e1y", "e1z", "e2x", "e2y", "e2z", "nx", "ny", "nz")] + \
               [("material", Material), ("bvh", C.POINTER(Bvh))]


class Volume(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("min_corner", C.c_float * 3), ("voxel_size", C.c_float * 3),
                ("mat", C.POINTER(C.c_int32)), ("meta", C.POINTER(C.c_int32)), ("wireframe", C.c_int32), ("wire_width_frac", C.c_float),
                ("wire_max_distance", C.c_float), ("palette_n_ids", C.c_int32), ("palette_meta_levels", C.c_int32),
                ("palette", C.POINTER(C.c_int32)), ("palette_default", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("diffuse_bounces", C.c_int32), ("max_mirror_bounces", C.c_int32), ("max_refractions", C.c_int32), ("atrous_iterations", C.c_int32),
                ("mirror_threshold", C.c_float), ("eps", C.c_float), ("taa_alpha", C.c_float), ("motion_trans_reset", C.c_float),
                ("motion_rot_reset", C.c_float), ("diffuse_sigma_deg", C.c_float), ("luminance_pad", C.c_float),
                ("c_phi", C.c_float), ("n_phi", C.c_float), ("z_phi", C.c_float), ("a_phi", C.c_float),
                ("tone_exposure", C.c_float), ("tone_gamma", C.c_float), ("ae_key", C.c_float), ("ae_speed", C.c_float), ("ae_min", C.c_float),
                ("ae_max", C.c_float), ("saturation", C.c_float), ("vibrance", C.c_float), ("auto_exposure", C.c_int32), ("seed_salt", C.c_uint64)]


class Config(C.Structure):
    _fields_ = [("fb_w", C.c_int32), ("fb_h", C.c_int32), ("ss", C.c_int32), ("device", C.c_int32), ("tile_row0", C.c_int32),
                ("tile_rows", C.c_int32), ("params", Params), ("n_devices", C.c_int32), ("devices", C.c_int32 * 8)]


class Halo(C.Structure):
    _fields_ = [("recv_ptr", C.c_void_p), ("recv_bytes", C.c_size_t), ("recv_row0", C.c_int32), ("recv_rows", C.c_int32),
                ("send_ptr", C.c_void_p), ("send_bytes", C.c_size_t), ("send_row0", C.c_int32), ("send_rows", C.c_int32)]


class Peer(C.Structure):
    _fields_ = [("sa", C.c_void_p), ("sb", C.c_void_p), ("flags", C.c_void_p),
                ("sa_ipc", C.c_ubyte * 64), ("sb_ipc", C.c_ubyte * 64), ("flags_ipc", C.c_ubyte * 64)]


class Stats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("rays", C.c_uint64), ("rays_total", C.c_uint64), ("top_nodes_popped", C.c_uint64), ("mesh_nodes_popped", C.c_uint64),
                ("leaf_refs", C.c_uint64), ("tris_tested", C.c_uint64), ("prims_tested", C.c_uint64), ("dda_cells", C.c_uint64),
                ("ms_trace", C.c_float), ("ms_taa", C.c_float), ("ms_atrous", C.c_float), ("ms_exposure", C.c_float), ("ms_cells", C.c_float),
                ("ms_total", C.c_float), ("ae_exposure", C.c_float), ("log_sum", C.c_float), ("log_cnt", C.c_int32), ("kernel_launches", C.c_int32),
                ("ms_atrous_chain", C.c_float), ("fast_div", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


CELL_DTYPE = np.dtype([("glyph", "<u2"), ("fg16", "u1"), ("bg16", "u1"), ("fg_ansi", "u1"), ("bg_ansi", "u1"), ("attr", "<u2"),
                       ("fg", "<f4", (3,)), ("bg", "<f4", (3,))])
assert CELL_DTYPE.itemsize == 32

DBG_RAYS, DBG_HDR, DBG_ALBEDO_SKY, DBG_NORMAL_DEPTH, DBG_TAA, DBG_DENOISED, DBG_PRIM_ID, DBG_LOG_SAMPLES = range(8)
PTR_CELLS, PTR_LOG_SAMPLES, PTR_HIST, PTR_GND, PTR_GAS, PTR_EXPOSURE = 0, 1, 2, 3, 4, 5

# every symbol include/ycge.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "ycge_default_params", "ycge_create", "ycge_destroy", "ycge_last_error", "ycge_resize", "ycge_mesh_upload_soa",
    "ycge_mesh_upload_triangles", "ycge_mesh_build_device", "ycge_mesh_debug_read", "ycge_volume_upload", "ycge_texture_upload", "ycge_scene_upload", "ycge_lights_update", "ycge_globals_update",
    "ycge_set_camera", "ycge_set_fov", "ycge_set_trace_variant", "ycge_set_inplace_variant", "ycge_reset_history", "ycge_render_frame", "ycge_render_frame_stats",
    "ycge_render_frames_async", "ycge_wait", "ycge_pipeline_config", "ycge_submit_frame", "ycge_frame_wait", "ycge_read_cells", "ycge_peer_export", "ycge_peer_attach", "ycge_stash_config", "ycge_frame_stash",
    "ycge_frame_finish_stashed", "ycge_stash_logs_ptr", "ycge_frame_begin", "ycge_frame_halo", "ycge_frame_inplace",
    "ycge_frame_finish", "ycge_frame_front", "ycge_back_config", "ycge_back_ptr", "ycge_back_denoise", "ycge_back_finish", "ycge_ansi_emit", "ycge_device_ptr",
    "ycge_set_stream", "ycge_debug_read", "ycge_get_stats", "ycge_get_frame_counter", "ycge_rng_kat",
]

_lib = None
_host = None


def library_source_digest() -> str:
    """sha256 over the sources libycge.so is built from -- the prerequisites of `libycge.so` in the Makefile, in the order listed,
    and the compiler flags (NVFLAGS).  Identifies a BUILD of the library for profiles/ncu_traffic.json: nvcc's output is not
    byte-reproducible (internal-linkage symbols carry a per-compilation id), the sources are."""
    import hashlib
    here = os.path.dirname(os.path.abspath(__file__))
    mk = open(os.path.join(here, "Makefile")).read().replace("\\\n", " ")
    deps = next(l for l in mk.splitlines() if l.startswith("libycge.so:")).split(":", 1)[1].split()
    flags = next(l for l in mk.splitlines() if l.startswith("NVFLAGS"))
    h = hashlib.sha256()
    h.update(flags.encode() + b"\0")
    for f in deps:
        h.update(os.path.basename(f).encode() + b"\0")
        h.update(open(os.path.join(here, f), "rb").read())
    return h.hexdigest()


def load_lib() -> C.CDLL:
    """Load libycge.so.  Raises (never falls back) if the extension was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback for the ray tracing path)")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        vp = C.c_void_p
        lib.ycge_last_error.restype = C.c_char_p
        lib.ycge_last_error.argtypes = [vp]
        lib.ycge_default_params.argtypes = [C.POINTER(Params)]
        lib.ycge_default_params.restype = None
        lib.ycge_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        lib.ycge_destroy.argtypes = [vp]
        lib.ycge_destroy.restype = None
        lib.ycge_resize.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
        lib.ycge_mesh_upload_soa.argtypes = [vp, C.c_int32, vp]
        lib.ycge_mesh_upload_triangles.argtypes = [vp, C.c_int32, C.c_int32, vp, vp]
        lib.ycge_mesh_build_device.argtypes = [vp, C.c_int32, C.c_int32, vp, vp]
        lib.ycge_mesh_debug_read.argtypes = [vp, C.c_int32, C.c_int32, vp, C.POINTER(C.c_size_t)]
        lib.ycge_volume_upload.argtypes = [vp, C.c_int32, vp]
        lib.ycge_scene_upload.argtypes = [vp, vp]
        lib.ycge_texture_upload.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp]
        lib.ycge_lights_update.argtypes = [vp, C.c_int32, vp]
        lib.ycge_globals_update.argtypes = [vp, vp, vp, vp, C.c_float]
        lib.ycge_set_camera.argtypes = [vp, vp, C.c_float, C.c_float]
        lib.ycge_set_fov.argtypes = [vp, C.c_float]
        lib.ycge_set_trace_variant.argtypes = [vp, C.c_int32]
        lib.ycge_set_inplace_variant.argtypes = [vp, C.c_int32]
        lib.ycge_reset_history.argtypes = [vp]
        lib.ycge_render_frame.argtypes = [vp, vp, C.c_int32]
        lib.ycge_render_frame_stats.argtypes = [vp, vp, C.c_int32]
        lib.ycge_render_frames_async.argtypes = [vp, C.c_int32]
        lib.ycge_wait.argtypes = [vp]
        lib.ycge_pipeline_config.argtypes = [vp, C.c_int32]
        lib.ycge_submit_frame.argtypes = [vp, vp, C.c_int32, C.POINTER(C.c_int64)]
        lib.ycge_frame_wait.argtypes = [vp, C.c_int64]
        lib.ycge_read_cells.argtypes = [vp, vp, C.c_int32]
        lib.ycge_frame_begin.argtypes = [vp]
        lib.ycge_frame_finish.argtypes = [vp]
        lib.ycge_frame_halo.argtypes = [vp, C.POINTER(Halo)]
        lib.ycge_stash_config.argtypes = [vp, C.c_int32]
        lib.ycge_frame_stash.argtypes = [vp, C.c_int32]
        lib.ycge_frame_finish_stashed.argtypes = [vp, C.c_int32, vp]
        lib.ycge_stash_logs_ptr.argtypes = [vp, C.c_int32, C.POINTER(vp), C.POINTER(C.c_size_t)]
        lib.ycge_peer_export.argtypes = [vp, C.POINTER(Peer)]
        lib.ycge_peer_attach.argtypes = [vp, C.POINTER(Peer), C.POINTER(Peer), C.c_int32]
        lib.ycge_frame_inplace.argtypes = [vp]
        lib.ycge_frame_front.argtypes = [vp]
        lib.ycge_back_config.argtypes = [vp, C.c_int32]
        lib.ycge_back_ptr.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(vp), C.POINTER(C.c_size_t)]
        lib.ycge_back_denoise.argtypes = [vp, C.c_int32, vp]
        lib.ycge_back_finish.argtypes = [vp, C.c_int32, vp]
        lib.ycge_device_ptr.argtypes = [vp, C.c_int32, C.POINTER(vp), C.POINTER(C.c_size_t)]
        lib.ycge_ansi_emit.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        lib.ycge_set_stream.argtypes = [vp, vp]
        lib.ycge_debug_read.argtypes = [vp, C.c_int32, vp, C.c_size_t]
        lib.ycge_get_stats.argtypes = [vp, C.POINTER(Stats)]
        lib.ycge_get_frame_counter.argtypes = [vp, C.POINTER(C.c_int64)]
        lib.ycge_rng_kat.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, C.c_int32, vp, vp]
        _lib = lib
    return _lib


def load_host() -> C.CDLL:
    global _host
    if _host is None:
        load_lib()
        if not os.path.exists(HOST_LIB_PATH):
            raise ImportError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
        h = C.CDLL(HOST_LIB_PATH)
        vp = C.c_void_p
        h.ycgeh_last_error.restype = C.c_char_p
        h.ycgeh_set_asset_dir.argtypes = [C.c_char_p]
        h.ycgeh_set_asset_dir.restype = None
        h.ycgeh_scene_create.argtypes = [C.c_char_p]
        h.ycgeh_scene_create.restype = vp
        h.ycgeh_scene_from_triangles.argtypes = [C.c_char_p, C.c_int, vp, C.c_int, vp]
        h.ycgeh_scene_from_triangles.restype = vp
        h.ycgeh_scene_destroy.argtypes = [vp]
        h.ycgeh_scene_destroy.restype = None
        h.ycgeh_scene_flat.argtypes = [vp]
        h.ycgeh_scene_flat.restype = C.POINTER(Scene)
        h.ycgeh_scene_n_meshes.argtypes = [vp]
        h.ycgeh_scene_mesh.argtypes = [vp, C.c_int]
        h.ycgeh_scene_mesh.restype = C.POINTER(MeshSoa)
        h.ycgeh_scene_n_volumes.argtypes = [vp]
        h.ycgeh_scene_write_snapshot.argtypes = [vp, C.c_char_p]
        h.ycgeh_scene_n_textures.argtypes = [vp]
        h.ycgeh_scene_texture.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        h.ycgeh_scene_texture.restype = C.POINTER(C.c_uint32)
        h.ycgeh_scene_set_texture.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        h.ycgeh_scene_volume.argtypes = [vp, C.c_int]
        h.ycgeh_scene_volume.restype = C.POINTER(Volume)
        h.ycgeh_scene_mesh_triangles.argtypes = [vp, C.c_int]
        h.ycgeh_scene_mesh_triangles.restype = C.POINTER(C.c_float)
        h.ycgeh_scene_camera.argtypes = [vp, vp, vp, vp, vp]
        h.ycgeh_scene_camera.restype = None
        h.ycgeh_scene_name.argtypes = [vp]
        h.ycgeh_scene_name.restype = C.c_char_p
        h.ycgeh_scene_counts.argtypes = [vp, vp, vp, vp, vp, vp]
        h.ycgeh_scene_bvh.argtypes = [vp, C.c_int, C.POINTER(C.POINTER(Bvh)), C.POINTER(C.c_uint64)]
        h.ycgeh_renderer_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        h.ycgeh_renderer_create.restype = vp
        h.ycgeh_renderer_create_multi.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        h.ycgeh_renderer_create_multi.restype = vp
        h.ycgeh_renderer_destroy.argtypes = [vp]
        h.ycgeh_renderer_destroy.restype = None
        h.ycgeh_renderer_ctx.argtypes = [vp]
        h.ycgeh_renderer_ctx.restype = vp
        h.ycgeh_renderer_set_camera.argtypes = [vp, vp, C.c_float, C.c_float]
        h.ycgeh_renderer_set_fov.argtypes = [vp, C.c_float]
        h.ycgeh_renderer_sync_lights.argtypes = [vp, vp]
        h.ycgeh_scene_update.argtypes = [vp, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        h.ycgeh_renderer_sync_geometry.argtypes = [vp, vp]
        h.ycgeh_scene_rebuild_bvh_ms.argtypes = [vp]
        h.ycgeh_scene_rebuild_bvh_ms.restype = C.c_double
        h.ycgeh_renderer_resize.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        h.ycgeh_renderer_render_cells.argtypes = [vp, vp]
        h.ycgeh_renderer_blit_ansi.argtypes = [vp, C.POINTER(C.POINTER(C.c_uint8))]
        h.ycgeh_renderer_blit_ansi.restype = C.c_int64
        h.ycgeh_renderer_charinfo.argtypes = [vp, vp]
        h.ycgeh_ansi_from_cells.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int64]
        h.ycgeh_ansi_from_cells.restype = C.c_int64
        h.ycgeh_synthetic_height.argtypes = [C.c_int, C.c_int, C.c_int]
        h.ycgeh_synthetic_height.restype = C.c_float
        h.ycgeh_write_synthetic_world.argtypes = [C.c_char_p, C.c_int, C.c_int]
        h.ycgeh_write_island_world.argtypes = [C.c_char_p, C.c_int, C.c_int]
        h.ycgeh_island_height.argtypes = [C.c_int] * 5
        h.ycgeh_gradient_noise2d.argtypes = [C.c_float, C.c_float, C.c_int]
        h.ycgeh_gradient_noise2d.restype = C.c_float
        _host = h
    return _host


def default_params() -> Params:
    p = Params()
    load_lib().ycge_default_params(C.byref(p))
    return p


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


# ---------------------------------------------------------------------------------------------- host-side scene
BENCH_POSE = ((0.0, 0.9, -1.4), float(np.float32(np.pi)), -0.15)  # default mesh poses look away from the mesh


class HostScene:
    """A scene built by the host mirror (the reference's BuildSceneTable() factories), flattened for the C ABI."""

    def __init__(self, name: str, asset_dir: Optional[str] = None):
        self._h = load_host()
        if asset_dir is None:
            # the engine's assets/ directory (ConsoleGame/assets); default: the committed binary twins of the reference meshes
            asset_dir = os.environ.get("YCGE_ASSETS", os.path.join(os.path.dirname(_PKG_DIR), "tests", "golden", "meshes"))
        self._h.ycgeh_set_asset_dir(asset_dir.encode())
        self.handle = self._h.ycgeh_scene_create(name.encode())
        if not self.handle:
            raise ValueError(self._h.ycgeh_last_error().decode())
        self.name = self._h.ycgeh_scene_name(self.handle).decode()

    @classmethod
    def from_triangles(cls, name: str, verts: np.ndarray, faces: np.ndarray) -> "HostScene":
        self = cls.__new__(cls)
        self._h = load_host()
        v = np.ascontiguousarray(verts, np.float32)
        f = np.ascontiguousarray(faces, np.int32)
        self.handle = self._h.ycgeh_scene_from_triangles(name.encode(), len(v), _ptr(v), len(f), _ptr(f))
        if not self.handle:
            raise ValueError(self._h.ycgeh_last_error().decode())
        self.name = name
        return self

    def close(self):
        if getattr(self, "handle", None):
            self._h.ycgeh_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def flat(self):
        return self._h.ycgeh_scene_flat(self.handle)

    @property
    def n_meshes(self) -> int:
        return self._h.ycgeh_scene_n_meshes(self.handle)

    def write_snapshot(self, path: str) -> int:
        """SceneSyncProtocol.WriteSnapshot (Scenes/SyncScene.cs:282-393): the engine's 'SCNE' v1 snapshot of this scene; load it
        back with HostScene("snapshot:<path>").  Returns the number of bytes written."""
        n = self._h.ycgeh_scene_write_snapshot(self.handle, path.encode())
        if n < 0:
            raise ValueError(self._h.ycgeh_last_error().decode())
        return n

    @property
    def n_textures(self) -> int:
        return self._h.ycgeh_scene_n_textures(self.handle)

    def texture(self, i: int) -> np.ndarray:
        """Texture i as an (h, w) uint32 array in the reference's int[] pixel layout (byte 0 = R; Texture.cs:81-90)."""
        w, hh = C.c_int(), C.c_int()
        p = self._h.ycgeh_scene_texture(self.handle, i, C.byref(w), C.byref(hh))
        return np.ctypeslib.as_array(p, shape=(hh.value, w.value)).copy()

    def set_texture(self, i: int, rgba: np.ndarray):
        """Replace texture i (before a renderer is created), e.g. with an image decoded by the caller."""
        a = np.ascontiguousarray(rgba, np.uint32)
        if self._h.ycgeh_scene_set_texture(self.handle, i, a.shape[1], a.shape[0], _ptr(a)) != 0:
            raise ValueError(self._h.ycgeh_last_error().decode())

    @property
    def n_volumes(self) -> int:
        return self._h.ycgeh_scene_n_volumes(self.handle)

    def mesh(self, i):
        return self._h.ycgeh_scene_mesh(self.handle, i)

    def mesh_triangles(self, i) -> np.ndarray:
        n = self.mesh(i).contents.n_tris
        p = self._h.ycgeh_scene_mesh_triangles(self.handle, i)
        return np.ctypeslib.as_array(p, shape=(n, 9)).copy()

    def volume(self, i):
        return self._h.ycgeh_scene_volume(self.handle, i)

    def update(self, delta_time_ms: float):
        """Scene.Update(deltaTimeMS) (Scenes/Scene.cs:100-127): the scene's entities run with dt = ms / 1000 (DayNightEntity rewrites
        the lights and the sky gradient, the animated entities move spheres and lights), then the tree is rebuilt if geometry
        changed; the flat view follows.  Returns (lights_version, geometry_version): push what changed with
        CudaRaytraceRenderer.SyncLights / SyncGeometry.  When geometry_version moved, the scene was flattened again: pointers
        taken earlier from .flat, .mesh(i), .volume(i) are stale -- fetch them again."""
        lv, gv = C.c_int(), C.c_int()
        if self._h.ycgeh_scene_update(self.handle, delta_time_ms, C.byref(lv), C.byref(gv)) != 0:
            raise YcgeError(-1, self._h.ycgeh_last_error().decode())
        return lv.value, gv.value

    def rebuild_bvh_ms(self) -> float:
        """Scene.RebuildBVH() (Scenes/Scene.cs:66-69) again, timed: the host-side cost of a geometry change, before SyncGeometry."""
        ms = self._h.ycgeh_scene_rebuild_bvh_ms(self.handle)
        if ms < 0:
            raise YcgeError(-1, self._h.ycgeh_last_error().decode())
        return ms

    def lights(self):
        f = self.flat.contents
        return [(tuple(f.lights[i].pos), tuple(f.lights[i].color), f.lights[i].intensity) for i in range(f.n_lights)]

    def background(self):
        f = self.flat.contents
        return tuple(f.bg_top), tuple(f.bg_bottom)

    def default_camera(self):
        pos = (C.c_float * 3)()
        yaw, pitch, fov = C.c_float(), C.c_float(), C.c_float()
        self._h.ycgeh_scene_camera(self.handle, pos, C.byref(yaw), C.byref(pitch), C.byref(fov))
        return (pos[0], pos[1], pos[2]), yaw.value, pitch.value, fov.value

    def counts(self):
        no, nl, nm = C.c_int(), C.c_int(), C.c_int()
        nt, nv = C.c_int64(), C.c_int64()
        self._h.ycgeh_scene_counts(self.handle, C.byref(no), C.byref(nl), C.byref(nm), C.byref(nt), C.byref(nv))
        return dict(objects=no.value, lights=nl.value, materials=nm.value, triangles=nt.value, voxels=nv.value)

    def bvh_arrays(self, which: int = -1):
        """The host-built tree (top level: which=-1, else mesh index) as numpy arrays."""
        pb = C.POINTER(Bvh)()
        sf = C.c_uint64()
        if self._h.ycgeh_scene_bvh(self.handle, which, C.byref(pb), C.byref(sf)) != 0:
            raise IndexError(which)
        b = pb.contents
        n = b.n_nodes
        boxes = np.stack([np.ctypeslib.as_array(getattr(b, k), shape=(n,)) for k in ("min_x", "min_y", "min_z", "max_x", "max_y", "max_z")], 1).copy() if n else np.zeros((0, 6), np.float32)
        lrsc = np.stack([np.ctypeslib.as_array(getattr(b, k), shape=(n,)) for k in ("left", "right", "start", "count")], 1).copy() if n else np.zeros((0, 4), np.int32)
        leaf = np.ctypeslib.as_array(b.leaf_index, shape=(b.n_leaf_refs,)).copy() if b.n_leaf_refs else np.zeros((0,), np.int32)
        return dict(root=b.root, boxes=boxes, lrsc=lrsc, leaf=leaf, sort_fallbacks=sf.value)


# ---------------------------------------------------------------------------------------------- the drop-in renderer
class CudaRaytraceRenderer:
    """Mirror of RaytraceRenderer's public surface (RaytraceRenderer.cs:74,110,140,150,157) behind IConsoleRenderer,
    producing frames through libycge.so.  ``tile_row0/tile_rows`` select a row tile for one-process-per-GPU sharding."""

    def __init__(self, scene: HostScene, fb_w: int, fb_h: int, super_sample: int = 1, device: int = 0, tile_row0: int = 0, tile_rows: int = 0, devices=None):
        """``devices``: two or more CUDA ordinals of one box -> ONE renderer over all of them (ycge_config.n_devices): frames in parallel
        inside the library, bit-identical to a one-GPU renderer; TryFlipAndBlit, submit_frame / frame_wait, wait, stats."""
        self._h = load_host()
        self._lib = load_lib()
        self.scene = scene
        self.fb_w, self.fb_h, self.ss = fb_w, fb_h, max(1, super_sample)
        self.tile_row0 = tile_row0
        self.tile_rows = tile_rows if tile_rows > 0 else fb_h - tile_row0
        if devices is not None and len(devices) >= 2:
            arr = (C.c_int * len(devices))(*devices)
            self.handle = self._h.ycgeh_renderer_create_multi(scene.handle, fb_w, fb_h, self.ss, len(devices), arr)
        else:
            self.handle = self._h.ycgeh_renderer_create(scene.handle, fb_w, fb_h, self.ss, device, tile_row0, tile_rows)
        if not self.handle:
            raise YcgeError(-2, self._h.ycgeh_last_error().decode())
        self.ctx = C.c_void_p(self._h.ycgeh_renderer_ctx(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self._h.ycgeh_renderer_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- IConsoleRenderer --
    def SetCamera(self, pos, yaw: float, pitch: float):
        p = (C.c_float * 3)(*pos)
        if self._h.ycgeh_renderer_set_camera(self.handle, p, yaw, pitch) != 0:
            raise YcgeError(-1, self._h.ycgeh_last_error().decode())

    def SyncLights(self, scene: "HostScene"):
        """After scene.update(dt): ycge_lights_update + ycge_globals_update with what the entities changed (no history reset,
        like the reference, whose renderer simply reads scene.Lights on the next frame)."""
        if self._h.ycgeh_renderer_sync_lights(self.handle, scene.handle) != 0:
            raise YcgeError(-1, self._h.ycgeh_last_error().decode())

    def SyncGeometry(self, scene: "HostScene"):
        """After scene.update(ms) moved objects: the object list and the rebuilt top-level tree again (ycge_scene_upload); meshes,
        volumes, textures and the TAA history stay (the reference rebuilds its BVH and renders on, Scene.cs:121-126)."""
        if self._h.ycgeh_renderer_sync_geometry(self.handle, scene.handle) != 0:
            raise YcgeError(-1, self._h.ycgeh_last_error().decode())

    def SetFov(self, fov_deg: float):
        if self._h.ycgeh_renderer_set_fov(self.handle, fov_deg) != 0:
            raise YcgeError(-1, self._h.ycgeh_last_error().decode())

    def Resize(self, fb_w: int, fb_h: int, super_sample: int):
        if self._h.ycgeh_renderer_resize(self.handle, fb_w, fb_h, super_sample) != 0:
            raise YcgeError(-1, self._h.ycgeh_last_error().decode())
        whole = self.tile_row0 == 0 and self.tile_rows == self.fb_h
        self.fb_w, self.fb_h, self.ss = fb_w, fb_h, max(1, super_sample)
        if whole:  # ycge_resize keeps the row tile of a sharded context (and refuses one that no longer fits)
            self.tile_row0, self.tile_rows = 0, fb_h

    def TryFlipAndBlit(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Synchronous frame: returns the (tile_rows, fb_w) cell array (the Framebuffer's Chexels)."""
        if out is None:
            out = np.empty((self.tile_rows, self.fb_w), CELL_DTYPE)
        if self._h.ycgeh_renderer_render_cells(self.handle, _ptr(out)) != 0:
            raise YcgeError(-2, self._h.ycgeh_last_error().decode())
        return out

    def blit_ansi(self) -> bytes:
        """TryFlipAndBlit into the host Framebuffer, then ANSITerminalRenderer.Render's byte stream."""
        p = C.POINTER(C.c_uint8)()
        n = self._h.ycgeh_renderer_blit_ansi(self.handle, C.byref(p))
        if n < 0:
            raise YcgeError(-2, self._h.ycgeh_last_error().decode())
        return C.string_at(p, n)

    # -- C-ABI passthroughs used by tests / bench --
    def _ck(self, rc: int):
        if rc != 0:
            raise YcgeError(rc, (self._lib.ycge_last_error(self.ctx) or b"").decode())

    def reset_history(self):
        self._ck(self._lib.ycge_reset_history(self.ctx))

    def set_trace_variant(self, variant: int):
        """0: one thread per pixel path; 1 (default): ray stream with lane refill.  Bit-identical results."""
        self._ck(self._lib.ycge_set_trace_variant(self.ctx, variant))

    def set_inplace_variant(self, variant: int):
        """0: one warp per chain of the in-place a-trous iteration (post.cuh); 1 (default): systolic bands (wavefront.cuh).  Bit-identical."""
        self._ck(self._lib.ycge_set_inplace_variant(self.ctx, variant))

    def rebuild_meshes_on_device(self, also_time_host_builder: bool = False) -> dict:
        """builds every mesh tree of the scene again ON THE DEVICE from the raw triangles (ycge_mesh_build_device:
        MeshBVH.BuildRecursive node for node) and re-syncs the object table; frames are bit-identical to those of the
        host-built trees.  Returns wall times in ms: 'device' = the call AND the wait for the finished tree (the call itself
        returns with the build in flight; the GPU part is 7 ms for 280 k triangles, the rest is staging the triangle list and
        allocating the mesh's arrays, which varies from call to call: with `also_time_host_builder` the best of three builds)
        and, on request, 'host' (the library's host builder on the same triangles, ycge_mesh_upload_triangles) -- what a scene
        switch pays for its mesh trees either way."""
        import time
        out = {"device": 0.0, "triangles": 0}
        if also_time_host_builder:
            out["host"] = 0.0
        for i in range(self.scene.n_meshes):
            tris = np.ascontiguousarray(self.scene.mesh_triangles(i), np.float32)
            mat = self.scene.mesh(i).contents.material
            out["triangles"] += len(tris)
            if also_time_host_builder:
                t0 = time.perf_counter()
                self._ck(self._lib.ycge_mesh_upload_triangles(self.ctx, i, len(tris), tris.ctypes.data, C.byref(mat)))
                out["host"] += 1e3 * (time.perf_counter() - t0)
            best = None
            for _ in range(3 if also_time_host_builder else 1):
                self._ck(self._lib.ycge_wait(self.ctx))  # whatever an earlier upload left on the stream is not this build's time
                t0 = time.perf_counter()
                self._ck(self._lib.ycge_mesh_build_device(self.ctx, i, len(tris), tris.ctypes.data, C.byref(mat)))
                n = C.c_size_t(0)
                self._lib.ycge_mesh_debug_read(self.ctx, i, 3, None, C.byref(n))  # asking for the root record waits for the build (fails, harmlessly, on a multi-GPU context)
                dt = 1e3 * (time.perf_counter() - t0)
                best = dt if best is None else min(best, dt)
            out["device"] += best
        self._ck(self._lib.ycge_scene_upload(self.ctx, self.scene.flat))
        return out

    def lights_update(self, lights):
        """Per-frame light changes without re-uploading geometry (DayNightCycle.cs:80-83): [(pos3, color3, intensity), ...]"""
        arr = (Light * len(lights))()
        for i, (pos, col, inten) in enumerate(lights):
            arr[i].pos[:] = pos
            arr[i].color[:] = col
            arr[i].intensity = inten
        self._ck(self._lib.ycge_lights_update(self.ctx, len(lights), arr))

    def globals_update(self, bg_top, bg_bottom, ambient_color, ambient_intensity):
        a, b, c = (C.c_float * 3)(*bg_top), (C.c_float * 3)(*bg_bottom), (C.c_float * 3)(*ambient_color)
        self._ck(self._lib.ycge_globals_update(self.ctx, a, b, c, ambient_intensity))

    def render_frame_stats(self) -> np.ndarray:
        out = np.empty((self.tile_rows, self.fb_w), CELL_DTYPE)
        self._ck(self._lib.ycge_render_frame_stats(self.ctx, _ptr(out), 0))
        return out

    def render_frames_async(self, n: int):
        self._ck(self._lib.ycge_render_frames_async(self.ctx, n))

    def wait(self):
        self._ck(self._lib.ycge_wait(self.ctx))

    def pipeline_config(self, n_slots: int):
        """Frames in flight on this GPU for render_frames_async / submit_frame (ycge_pipeline_config); 1 = strictly serial."""
        self._ck(self._lib.ycge_pipeline_config(self.ctx, n_slots))

    def submit_frame(self, out: np.ndarray) -> int:
        """TryFlipAndBlit without the wait: the frame's cells land in ``out`` (keep it alive; pinned memory makes the copy
        asynchronous) once ``frame_wait(id)`` returns."""
        fid = C.c_int64()
        self._ck(self._lib.ycge_submit_frame(self.ctx, _ptr(out), 0, C.byref(fid)))
        return fid.value

    def frame_wait(self, frame_id: int):
        self._ck(self._lib.ycge_frame_wait(self.ctx, frame_id))

    def read_cells(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.tile_rows, self.fb_w), CELL_DTYPE)
        self._ck(self._lib.ycge_read_cells(self.ctx, _ptr(out), 0))
        return out

    def frame_begin(self):
        self._ck(self._lib.ycge_frame_begin(self.ctx))

    def frame_finish(self):
        self._ck(self._lib.ycge_frame_finish(self.ctx))

    def frame_halo(self) -> Optional[Halo]:
        """The boundary rows of a pending in-place pass (None when no pass is pending)."""
        h = Halo()
        rc = self._lib.ycge_frame_halo(self.ctx, C.byref(h))
        if rc < 0:
            self._ck(rc)
        return h if rc == 1 else None

    def frame_inplace(self):
        self._ck(self._lib.ycge_frame_inplace(self.ctx))

    # -- frame-parallel sharding (ycge.h: FRONT on row tiles, BACK of whole frames round-robin over the ranks)
    def frame_front(self):
        self._ck(self._lib.ycge_frame_front(self.ctx))

    def back_config(self, n_slots: int):
        self._ck(self._lib.ycge_back_config(self.ctx, n_slots))

    def back_ptr(self, slot: int, kind: int):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.ycge_back_ptr(self.ctx, slot, kind, C.byref(p), C.byref(n)))
        return p.value, n.value

    def back_denoise(self, slot: int, cuda_stream: int):
        self._ck(self._lib.ycge_back_denoise(self.ctx, slot, C.c_void_p(cuda_stream)))

    def back_finish(self, slot: int, cuda_stream: int):
        self._ck(self._lib.ycge_back_finish(self.ctx, slot, C.c_void_p(cuda_stream)))

    def ansi_stream(self) -> bytes:
        """ANSITerminalRenderer.Render's byte stream for the last frame, produced on the device (ycge_ansi_emit)."""
        cap = 64 + self.tile_rows * 16 + self.tile_rows * self.fb_w * 24
        buf = np.empty(cap, np.uint8)
        n = C.c_size_t()
        self._ck(self._lib.ycge_ansi_emit(self.ctx, _ptr(buf), cap, C.byref(n)))
        return buf[:n.value].tobytes()

    def stash_config(self, n_slots: int):
        self._ck(self._lib.ycge_stash_config(self.ctx, n_slots))

    def frame_stash(self, slot: int):
        self._ck(self._lib.ycge_frame_stash(self.ctx, slot))

    def frame_finish_stashed(self, slot: int, cuda_stream: int):
        self._ck(self._lib.ycge_frame_finish_stashed(self.ctx, slot, C.c_void_p(cuda_stream)))

    def stash_logs_ptr(self, slot: int):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.ycge_stash_logs_ptr(self.ctx, slot, C.byref(p), C.byref(n)))
        return p.value, n.value

    def peer_export(self) -> Peer:
        p = Peer()
        self._ck(self._lib.ycge_peer_export(self.ctx, C.byref(p)))
        return p

    def peer_attach(self, above: Optional[Peer], below: Optional[Peer], via_ipc: bool):
        self._ck(self._lib.ycge_peer_attach(self.ctx, C.byref(above) if above is not None else None,
                                            C.byref(below) if below is not None else None, 1 if via_ipc else 0))

    def set_stream(self, cuda_stream: int):
        self._ck(self._lib.ycge_set_stream(self.ctx, C.c_void_p(cuda_stream)))

    def device_ptr(self, kind: int):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.ycge_device_ptr(self.ctx, kind, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stats(self) -> dict:
        s = Stats()
        self._ck(self._lib.ycge_get_stats(self.ctx, C.byref(s)))
        return s.as_dict()

    @property
    def hi_w(self):
        return self.fb_w * self.ss

    @property
    def hi_h(self):
        return self.fb_h * 2 * self.ss

    def debug_read(self, kind: int) -> np.ndarray:
        n = self.hi_w * self.hi_h
        if kind == DBG_RAYS:
            a = np.empty((self.hi_h, self.hi_w, 6), np.float32)
        elif kind == DBG_PRIM_ID:
            a = np.empty((self.hi_h, self.hi_w, 2), np.int32)
        elif kind == DBG_LOG_SAMPLES:
            step = max(2, self.ss * 2)
            a = np.empty(((self.hi_h + step - 1) // step, (self.hi_w + step - 1) // step), np.float32)
        else:
            a = np.empty((self.hi_h, self.hi_w, 4), np.float32)
        rc = self._lib.ycge_debug_read(self.ctx, kind, _ptr(a), a.nbytes)
        if rc != 0 and kind == DBG_RAYS:  # first call only arms the tap
            return None
        self._ck(rc)
        return a

    def rng_kat(self, which: int, xs, ys, frames, n_draws: int):
        x = np.ascontiguousarray(xs, np.int32)
        y = np.ascontiguousarray(ys, np.int32)
        f = np.ascontiguousarray(frames, np.int64)
        bits = np.empty((len(x), n_draws), np.uint32)
        seeds = np.empty(len(x), np.uint64)
        self._ck(self._lib.ycge_rng_kat(self.ctx, which, len(x), _ptr(x), _ptr(y), _ptr(f), n_draws, _ptr(bits), _ptr(seeds)))
        return bits, seeds


def ansi_from_cells(cells: np.ndarray) -> bytes:
    """ANSITerminalRenderer.Render's byte stream (ANSITerminalRenderer.cs:86-153) for a cell array (host-side C++)."""
    h = load_host()
    cells = np.ascontiguousarray(cells)
    fb_h, fb_w = cells.shape
    cap = 64 + fb_w * fb_h * 32
    buf = np.empty(cap, np.uint8)
    n = h.ycgeh_ansi_from_cells(_ptr(cells), fb_w, fb_h, _ptr(buf), cap)
    if n < 0:
        raise YcgeError(-1, "ANSI buffer too small")
    return buf[:n].tobytes()
