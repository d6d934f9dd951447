"""the device builder (ycge_mesh_build_device, csrc/bvh_device.cuh) against the host restatement of
MeshBVH.BuildRecursive (ycge_mesh_upload_triangles -> csrc/bvh_build.hpp, itself pinned against the transpiled reference and
a numpy transcription in tests/test_bvh_builder_literal.py, tests/test_reference_transpiled.py): every pair node, every
leaf-ordered triangle, every leaf slot and the root record must be the same bytes."""
import ctypes as C
import os
import sys
import time

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yetanotherconsolegameengine_b200 import api  # noqa: E402

pytestmark = pytest.mark.gpu


def _ctx(lib):
    cfg = api.Config()
    cfg.fb_w, cfg.fb_h, cfg.ss = 8, 4, 1
    lib.ycge_default_params(C.byref(cfg.params))
    ctx = C.c_void_p()
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == 0
    return ctx


def _arrays(lib, ctx, mesh_id):
    out = []
    for what in range(4):
        n = C.c_size_t(0)
        assert lib.ycge_mesh_debug_read(ctx, mesh_id, what, None, C.byref(n)) == 0
        buf = np.zeros(max(1, n.value), np.uint8)
        assert lib.ycge_mesh_debug_read(ctx, mesh_id, what, buf.ctypes.data, C.byref(n)) == 0
        out.append(buf[: n.value].copy())
    return out


def _compare(lib, ctx, tris, material, label):
    tris = np.ascontiguousarray(tris, np.float32)
    t0 = time.perf_counter()
    assert lib.ycge_mesh_upload_triangles(ctx, 1, len(tris), tris.ctypes.data, C.byref(material)) == 0
    t1 = time.perf_counter()
    assert lib.ycge_mesh_build_device(ctx, 2, len(tris), tris.ctypes.data, C.byref(material)) == 0, lib.ycge_last_error(ctx)
    t2 = time.perf_counter()
    host, dev = _arrays(lib, ctx, 1), _arrays(lib, ctx, 2)
    for name, h, d in zip(("pair nodes", "triangles", "leaf slots", "root"), host, dev):
        assert h.size == d.size, f"{label}: {name}: {h.size} vs {d.size} bytes"
        if not np.array_equal(h, d):
            w = 64 if name == "pair nodes" else (48 if name == "triangles" else 4)
            bad = np.nonzero(h != d)[0]
            raise AssertionError(f"{label}: {name} differ, first at record {bad[0] // w} of {h.size // w} ({len(bad)} bytes in all)")
    return t1 - t0, t2 - t1


@pytest.mark.parametrize("scene", ["knot:40x12", "teapot", "cow", "bunny", "dragon"])
def test_device_built_tree_equals_the_host_built_tree(scene):
    lib = api.load_lib()
    s = api.HostScene(scene)
    ctx = _ctx(lib)
    tris = s.mesh_triangles(0)
    th, td = _compare(lib, ctx, tris, s.mesh(0).contents.material, scene)
    print(f"{scene}: {len(tris)} triangles, host build + upload {th * 1e3:.1f} ms, device build {td * 1e3:.1f} ms")
    lib.ycge_destroy(ctx)
    s.close()


def test_small_degenerate_and_duplicate_inputs():
    """Leaf-only trees, one split, many identical triangles (all centroids equal: the Array.Sort fallback), a line of
    triangles with equal centroids on two axes, random soup."""
    lib = api.load_lib()
    s = api.HostScene("teapot")
    mat = s.mesh(0).contents.material
    ctx = _ctx(lib)
    rnd = np.random.default_rng(3)
    base = rnd.random((1, 9), dtype=np.float32)
    cases = {
        "1 triangle": rnd.random((1, 9), dtype=np.float32),
        "8 triangles": rnd.random((8, 9), dtype=np.float32),
        "9 triangles": rnd.random((9, 9), dtype=np.float32),
        "17 triangles": rnd.random((17, 9), dtype=np.float32),
        "100 copies of one triangle": np.repeat(base, 100, axis=0),
        "1000 copies of one triangle": np.repeat(base, 1000, axis=0),
        "three stacks of duplicates": np.concatenate([np.repeat(rnd.random((1, 9), dtype=np.float32) + k, 40, axis=0) for k in range(3)]),
        "a line along x": np.stack([np.array([k, 0, 0, k + 1, 0, 0, k, 1, 0], np.float32) for k in range(300)]),
        "random soup 5000": (rnd.random((5000, 9), dtype=np.float32) * 10).astype(np.float32),
        "clustered soup 20000": (rnd.normal(size=(20000, 1, 3)).astype(np.float32) * 5 + rnd.random((20000, 3, 3), dtype=np.float32) * 0.2).reshape(20000, 9).astype(np.float32),
    }
    for label, tris in cases.items():
        _compare(lib, ctx, tris, mat, label)
    lib.ycge_destroy(ctx)
    s.close()


def test_frame_rendered_from_a_device_built_mesh_equals_the_default_path():
    """The bunny scene with its mesh tree built on the device: same cells as with the host-uploaded tree."""
    lib = api.load_lib()
    s = api.HostScene("bunny")
    r1 = api.CudaRaytraceRenderer(s, 48, 14, 2)
    r1.SetCamera(*api.BENCH_POSE)
    a = r1.TryFlipAndBlit().copy()
    cfg = api.Config()
    cfg.fb_w, cfg.fb_h, cfg.ss = 48, 14, 2
    lib.ycge_default_params(C.byref(cfg.params))
    ctx = C.c_void_p()
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == 0
    tris = s.mesh_triangles(0)
    assert lib.ycge_mesh_build_device(ctx, 0, len(tris), tris.ctypes.data, C.byref(s.mesh(0).contents.material)) == 0
    assert lib.ycge_scene_upload(ctx, s.flat) == 0
    pos = (C.c_float * 3)(*api.BENCH_POSE[0])
    lib.ycge_set_camera(ctx, pos, api.BENCH_POSE[1], api.BENCH_POSE[2])
    b = np.empty((14, 48), api.CELL_DTYPE)
    assert lib.ycge_render_frame(ctx, b.ctypes.data, 0) == 0
    for k in ("glyph", "fg16", "bg16", "fg_ansi", "bg_ansi"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["fg"].view(np.uint32), b["fg"].view(np.uint32)) and np.array_equal(a["bg"].view(np.uint32), b["bg"].view(np.uint32))
    lib.ycge_destroy(ctx)
    r1.close()
    s.close()
