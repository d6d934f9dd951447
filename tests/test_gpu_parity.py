"""GPU parity tests (run on the B200 with -m gpu): the CUDA path, called through the C ABI (libycge.so) behind the
IConsoleRenderer mirror, against the CPU oracle and the committed golden vectors.

Bar (BASELINE.json north_star): cells (glyph, ANSI-256, 16-colour index) and primary-hit ids bit-exact except documented
float ties <= 0.1 % of cells; linear RGB within 1e-3 absolute.  Both sides evaluate transcendentals through
include/ycge_detmath.h and never contract a*b+c, so the tests demand MORE: every intermediate plane bit-identical.
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, require_cuda
from yetanotherconsolegameengine_b200 import api
from oracle_binding import Oracle

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu

RGB_TOL = 1e-3          # north_star: linear RGB within 1e-3 absolute per channel
TIE_FRACTION = 0.001    # north_star: <= 0.1 % of cells may differ at documented float ties
CELL_KEYS = ("glyph", "fg16", "bg16", "fg_ansi", "bg_ansi", "attr")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits_differ(a, b):
    return int((np.ascontiguousarray(a).view(np.uint32) != np.ascontiguousarray(b).view(np.uint32)).sum())


def assert_cells_equal(g, c, what=""):
    for k in CELL_KEYS:
        assert np.array_equal(g[k], c[k]), f"{what}: cell field {k} differs in {int((g[k] != c[k]).sum())} cells"
    assert bits_differ(g["fg"], c["fg"]) == 0 and bits_differ(g["bg"], c["bg"]) == 0, f"{what}: SDR colour bits differ"


def assert_frame_parity(r, o, what, check_rays=False):
    """All taps of one rendered frame, GPU vs oracle.  First the north-star gates, then the stricter bit-identity."""
    gp, cp = r.debug_read(api.DBG_PRIM_ID), o.debug_read(api.DBG_PRIM_ID)
    assert (gp != cp).any(-1).mean() <= TIE_FRACTION, f"{what}: primary ids"
    gd, cd = r.debug_read(api.DBG_DENOISED), o.debug_read(api.DBG_DENOISED)
    finite = np.isfinite(cd) & np.isfinite(gd)
    assert np.array_equal(np.isfinite(cd), np.isfinite(gd))
    assert np.abs(gd[finite] - cd[finite]).max(initial=0.0) <= RGB_TOL, f"{what}: linear RGB"
    assert np.array_equal(gp, cp), f"{what}: primary objId/subId not bit-exact"
    kinds = [(api.DBG_HDR, "hdr"), (api.DBG_ALBEDO_SKY, "albedo/sky"), (api.DBG_NORMAL_DEPTH, "normal/depth"), (api.DBG_TAA, "taa"),
             (api.DBG_DENOISED, "denoised"), (api.DBG_LOG_SAMPLES, "log samples")]
    if check_rays:
        kinds.insert(0, (api.DBG_RAYS, "rays"))
    for kind, nm in kinds:
        assert bits_differ(r.debug_read(kind), o.debug_read(kind)) == 0, f"{what}: {nm} not bit-identical"
    gs, cs = r.stats(), o.stats()
    assert gs["rays"] == cs["rays"], f"{what}: ray count"
    assert np.float32(gs["ae_exposure"]).view(np.uint32) == np.float32(cs["ae_exposure"]).view(np.uint32), f"{what}: aeExposure"
    assert np.float32(gs["log_sum"]).view(np.uint32) == np.float32(cs["log_sum"]).view(np.uint32) and gs["log_cnt"] == cs["log_cnt"]


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    require_cuda()


# ------------------------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("case", make_golden.CASES, ids=[c[0] for c in make_golden.CASES])
def test_gpu_matches_golden(case):
    name, scene, fb_w, fb_h, ss, frames, pose = case
    gold = np.load(os.path.join(GOLDEN, f"oracle_{name}.npz"))
    s = api.HostScene(scene)
    r = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    if pose is not None:
        r.SetCamera(*pose)
    for f in range(1, frames + 1):
        cells = r.TryFlipAndBlit()
        assert_cells_equal(cells, gold[f"cells_{f}"], f"{name} frame {f}")
        assert np.array_equal(r.debug_read(api.DBG_PRIM_ID), gold[f"prim_{f}"])
        assert sha(r.debug_read(api.DBG_HDR)[..., :3]) == str(gold[f"sha_hdr_{f}"])
        assert sha(r.debug_read(api.DBG_TAA)[..., :3]) == str(gold[f"sha_taa_{f}"])
        assert sha(r.debug_read(api.DBG_DENOISED)[..., :3]) == str(gold[f"sha_den_{f}"])
        st = r.stats()
        assert st["rays"] == int(gold[f"rays_{f}"])
        assert np.float32(st["ae_exposure"]).view(np.uint32) == gold[f"ae_{f}"].view(np.uint32)
    r.close()
    s.close()


@pytest.mark.parametrize("case", make_golden.CASES, ids=[c[0] for c in make_golden.CASES])
def test_warp_per_chain_in_place_pass_matches_golden(case):
    """The other form of the in-place a-trous iteration (ycge_set_inplace_variant(0): one warp per chain, csrc/post.cuh; the
    default is 1: bands of 4 rows in lock step, csrc/wavefront.cuh, whose schedule is replayed on the CPU by
    tests/test_wavefront_layout.py) against the same golden vectors."""
    name, scene, fb_w, fb_h, ss, frames, pose = case
    gold = np.load(os.path.join(GOLDEN, f"oracle_{name}.npz"))
    s = api.HostScene(scene)
    r = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    r.set_inplace_variant(0)
    if pose is not None:
        r.SetCamera(*pose)
    for f in range(1, frames + 1):
        cells = r.TryFlipAndBlit()
        assert sha(r.debug_read(api.DBG_DENOISED)[..., :3]) == str(gold[f"sha_den_{f}"]), f"{name} frame {f}"
        assert_cells_equal(cells, gold[f"cells_{f}"], f"{name} frame {f}")
    r.close()
    s.close()


def test_both_forms_of_the_in_place_pass_agree_at_full_size():
    """1920x1080 (dragon stand-in, bench pose) and an odd-sized frame: both forms of the in-place pass, every denoised bit."""
    for scene, fb_w, fb_h, ss, pose in (("dragon", 480, 135, 4, api.BENCH_POSE), ("bunny", 61, 23, 3, api.BENCH_POSE)):
        s = api.HostScene(scene)
        a, b = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss), api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
        a.set_inplace_variant(1); b.set_inplace_variant(0)
        for r in (a, b):
            r.SetCamera(*pose)
        for f in range(2):
            ca, cb = a.TryFlipAndBlit(), b.TryFlipAndBlit()
            assert bits_differ(a.debug_read(api.DBG_DENOISED), b.debug_read(api.DBG_DENOISED)) == 0, f"{scene} frame {f + 1}"
            assert_cells_equal(ca, cb, f"{scene} frame {f + 1}")
        a.close(); b.close(); s.close()


# ------------------------------------------------------------------------------------------- the reference's own source, no oracle in between
REF_CASES = [("cornell", 40, 12, 2, None), ("mirror_spheres", 48, 14, 2, None), ("cylinders_disks_triangles", 32, 10, 2, None), ("teapot", 40, 12, 2, api.BENCH_POSE),
             ("knot:60x16", 36, 10, 3, api.BENCH_POSE), ("volume_grid_test", 40, 12, 2, None), ("voxel_world:64x64", 36, 10, 2, None),
             ("texture_gallery", 48, 14, 2, None)]  # SampleAlbedo -> Texture.SampleBilinear of the reference's text on rects, box faces, a triangle, a mesh, glass


@pytest.mark.parametrize("case", REF_CASES, ids=[c[0] for c in REF_CASES])
def test_gpu_equals_the_transpiled_reference(case):
    """The CUDA path against oracle/_ref -- the reference's OWN C# source text rewritten into C++ syntactically at build time
    (oracle/ref_transpile.py; the built library travels to this box) -- with no hand-written oracle in between: four frames of
    TryFlipAndBlit (trace stage + TAA + a-trous + exposure + cells), every plane and every cell field bit for bit."""
    import ref_binding
    if not ref_binding.available():
        pytest.fail("oracle/_ref/libycge_ref.so did not travel to the GPU box (it is built where /root/reference exists)")
    scene, fb_w, fb_h, ss, pose = case
    s = api.HostScene(scene)
    r = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    ref = ref_binding.RefRenderer(s, fb_w, fb_h, ss)
    if pose is not None:
        r.SetCamera(*pose)
        ref.set_camera(*pose)
    for frame in range(1, 5):
        g = r.TryFlipAndBlit()
        c = ref.render_frame()
        what = f"{scene} frame {frame}"
        asky = r.debug_read(api.DBG_ALBEDO_SKY)
        assert np.array_equal(asky[..., 3] != 0, c["sky"] != 0), what + ": sky mask"
        assert bits_differ(r.debug_read(api.DBG_HDR)[..., :3], c["hdr"]) == 0, what + ": radiance"
        assert bits_differ(asky[..., :3], c["albedo"]) == 0, what + ": albedo"
        assert bits_differ(r.debug_read(api.DBG_NORMAL_DEPTH)[..., 3], c["depth"]) == 0, what + ": depth"
        assert bits_differ(r.debug_read(api.DBG_TAA)[..., :3], c["taa"]) == 0, what + ": TAA history"
        assert bits_differ(r.debug_read(api.DBG_DENOISED)[..., :3], c["den"]) == 0, what + ": denoised"
        assert np.float32(r.stats()["ae_exposure"]).view(np.uint32) == c["expo"][:1].view(np.uint32)[0], what + ": aeExposure"
        for k in ("glyph", "fg16", "bg16", "fg_ansi", "bg_ansi"):
            assert np.array_equal(g[k], c[k]), what + ": cell field " + k
        assert bits_differ(g["fg"], c["fg"]) == 0 and bits_differ(g["bg"], c["bg"]) == 0, what + ": SDR colours"
    ref.close(); r.close(); s.close()


# ------------------------------------------------------------------------------------------- live oracle, every tap
LIVE = [  # scene, fb_w, fb_h, ss, frames, pose
    ("cornell", 60, 34, 1, 3, None),
    ("mirror_spheres", 48, 14, 4, 3, None),
    ("cylinders_disks_triangles", 60, 34, 1, 2, None),
    ("boxes", 33, 9, 3, 2, None),            # ragged: W = 99, H = 54 are not multiples of the 16x8 / 32x8 tiles
    ("test", 31, 7, 2, 2, None),
    ("volume_grid_test", 60, 34, 1, 2, None),
    ("teapot", 48, 14, 4, 2, api.BENCH_POSE),
    ("bunny", 64, 18, 2, 2, api.BENCH_POSE),
    ("knot:60x16", 48, 14, 4, 2, api.BENCH_POSE),
    ("voxel_world:64x64", 48, 14, 4, 2, None),
    ("cornell", 1, 1, 1, 2, None),           # minimum size: 1x2 pixels
    ("texture_gallery", 64, 18, 3, 2, None),  # SampleAlbedo + Texture.SampleBilinear on rects, box faces, a triangle, a mesh, glass
    ("texture_gallery", 48, 14, 2, 2, ((1.2, 1.4, -0.6), 0.5, -0.3)),
    ("texture_test", 40, 12, 2, 2, ((0.6, 0.4, 0.0), 0.25, -0.2)),   # BuildTextureTestScene (Scenes.cs:337-358), stand-in image
]


@pytest.mark.parametrize("case", LIVE, ids=[f"{c[0]}-{c[1]}x{c[2]}ss{c[3]}" for c in LIVE])
def test_gpu_matches_oracle_every_tap(case):
    scene, fb_w, fb_h, ss, frames, pose = case
    s = api.HostScene(scene)
    r = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    o = Oracle(s, fb_w, fb_h, ss)
    if pose is not None:
        r.SetCamera(*pose)
        o.set_camera(*pose)
    r.debug_read(api.DBG_RAYS)  # arms the ray tap
    for f in range(frames):
        g = r.render_frame_stats()
        c = o.render_frame(threads=os.cpu_count() or 1, fast_post=True)
        assert_cells_equal(g, c, f"{scene} frame {f + 1}")
        assert_frame_parity(r, o, f"{scene} frame {f + 1}", check_rays=True)
        gs, cs = r.stats(), o.stats()
        for k in ("top_nodes_popped", "mesh_nodes_popped", "leaf_refs", "tris_tested", "prims_tested", "dda_cells"):
            assert gs[k] == cs[k], f"{scene}: traversal event counter {k}: {gs[k]} vs {cs[k]}"
    r.close()
    o.close()
    s.close()


def morton_index(nbx, nby, x, y, z):
    """VolumeGrid.IndexOf (VolumeGrid.cs:235-252): brick ((bz * nby) + by) * nbx + bx, inside it the bits x0 y0 z0 x1 y1 z1 x2 y2 z2."""
    lx, ly, lz = x & 7, y & 7, z & 7
    m = (lx & 1) | ((ly & 1) << 1) | ((lz & 1) << 2) | ((lx & 2) << 2) | ((ly & 2) << 3) | ((lz & 2) << 4) | ((lx & 4) << 4) | ((ly & 4) << 5) | ((lz & 4) << 6)
    return ((((z >> 3) * nby) + (y >> 3)) * nbx + (x >> 3)) * 512 + m


@pytest.mark.gpu
@pytest.mark.parametrize("seed,fill", [(1, 0.02), (2, 0.10), (3, 0.45), (4, 0.0)])
def test_sparse_random_voxels_match_the_oracle(seed, fill):
    """The DDA fetches a voxel only in octants (4^3 blocks) its occupancy byte marks solid (trace.cuh volume_hit, post.cuh
    voxel_occupancy_kernel): the small voxel test grid refilled at random -- nearly empty, sparse, half full, empty -- puts every
    mix of empty and solid octants on the rays' way.  Planes, cells and the DDA cell count equal the oracle's."""
    s = api.HostScene("volume_grid_test")
    v = s.volume(0).contents
    nbx, nby, nbz = (v.nx + 7) >> 3, (v.ny + 7) >> 3, (v.nz + 7) >> 3
    mat = np.ctypeslib.as_array(v.mat, shape=(nbx * nby * nbz * 512,))
    rng = np.random.default_rng(seed)
    mat[:] = 0
    for z in range(v.nz):
        for y in range(v.ny):
            for x in range(v.nx):
                if rng.random() < fill:
                    mat[morton_index(nbx, nby, x, y, z)] = int(rng.integers(1, v.palette_n_ids))
    r = api.CudaRaytraceRenderer(s, 44, 12, 2)
    o = Oracle(s, 44, 12, 2)
    pose = ((0.4, 5.5, 5.0), 0.05, -0.46)  # outside the grid (x -4..4, y 0..4, z -6..2), looking down and across it
    r.SetCamera(*pose)
    o.set_camera(*pose)
    for f in range(2):
        g = r.render_frame_stats()
        c = o.render_frame(threads=os.cpu_count() or 1, fast_post=True)
        assert_cells_equal(g, c, f"seed {seed} frame {f + 1}")
        assert_frame_parity(r, o, f"seed {seed} frame {f + 1}")
        assert r.stats()["dda_cells"] == o.stats()["dda_cells"] and (fill == 0.0 or r.stats()["dda_cells"] > 0)
    r.close()
    o.close()
    s.close()


@pytest.mark.parametrize("scene,fb_w,fb_h,ss,pose", [("mirror_spheres", 57, 19, 3, None), ("knot:60x16", 48, 27, 4, api.BENCH_POSE),
                                                     ("voxel_world:64x64", 50, 13, 2, None), ("texture_gallery", 41, 11, 2, None)])
def test_trace_kernel_forms_are_bit_identical(scene, fb_w, fb_h, ss, pose):
    """ycge_set_trace_variant: the ray-stream kernel (default; lanes refilled per ray) and the thread-per-path kernel perform
    the same per-pixel arithmetic: every plane and every traversal event counter must agree bit for bit."""
    s = api.HostScene(scene)
    a = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    b = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    a.set_trace_variant(0)
    b.set_trace_variant(1)
    for r in (a, b):
        if pose is not None:
            r.SetCamera(*pose)
        r.debug_read(api.DBG_RAYS)
    for f in range(3):
        ca, cb = a.render_frame_stats(), b.render_frame_stats()
        assert_cells_equal(ca, cb, f"{scene} frame {f + 1}")
        for kind in (api.DBG_RAYS, api.DBG_HDR, api.DBG_ALBEDO_SKY, api.DBG_NORMAL_DEPTH, api.DBG_TAA, api.DBG_DENOISED, api.DBG_PRIM_ID):
            assert bits_differ(a.debug_read(kind), b.debug_read(kind)) == 0, (scene, f, kind)
        sa, sb = a.stats(), b.stats()
        for k in ("rays", "top_nodes_popped", "mesh_nodes_popped", "leaf_refs", "tris_tested", "prims_tested", "dda_cells"):
            assert sa[k] == sb[k], (scene, k, sa[k], sb[k])
    a.close()
    b.close()
    s.close()


@pytest.mark.parametrize("scene,fb_w,fb_h,ss", [("test", 8, 3, 2), ("knot:12x5", 8, 3, 2), ("volume_grid_test", 8, 3, 2)])
def test_gpu_trace_matches_the_literal_python_transcription(scene, fb_w, fb_h, ss):
    """The CUDA trace stage against the numpy transcription of the C# source (tests/test_oracle_trace_literal.py) directly, without
    the C++ oracle in between: radiance, G-buffer, sky flag, primary ids and the ray count of every pixel, bit for bit."""
    from test_oracle_trace_literal import F, LiteralTracer, Rng, frac, per_frame_seed, v3
    from oracle_binding import load_oracle
    lib = load_oracle()  # only ycge_detmath.h's sin / cos / tan / pow behind yo_math
    lib.yo_set_math_mode(0)
    s = api.HostScene(scene)
    r = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    lt = LiteralTracer(s, lib)
    pos, yaw, pitch, fov = s.default_camera()
    if scene.startswith("knot"):
        pos, yaw, pitch = api.BENCH_POSE
        r.SetCamera(pos, yaw, pitch)
    cam, yaw, pitch, fov = v3(*pos), F(yaw), F(pitch), F(fov)
    w, h = fb_w * ss, fb_h * 2 * ss
    aspect = F(F(w) / F(h))
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for frame in (1, 2):
            r.TryFlipAndBlit()
            hdr, als, nd, prim = r.debug_read(api.DBG_HDR), r.debug_read(api.DBG_ALBEDO_SKY), r.debug_read(api.DBG_NORMAL_DEPTH), r.debug_read(api.DBG_PRIM_ID)
            frame_idx = frame & 0x7FFFFFFF
            jrx, jry = frac(F(F(frame_idx + 1) * F(0.61803398875))), frac(F(F(frame_idx + 1) * F(0.38196601125)))
            lt.rays = 0
            for py in range(h):
                for px in range(w):
                    ro, rd = lt.make_ray(cam, yaw, pitch, fov, aspect, px, py, w, h, jrx, jry, frame_idx)
                    rad, is_sky, g = lt.trace_full(ro, rd, Rng(per_frame_seed(px, py, frame)))
                    where = (scene, frame, px, py)
                    assert np.array_equal(rad.view(np.uint32), hdr[py, px, :3].view(np.uint32)), where + ("radiance",)
                    assert bool(als[py, px, 3]) == is_sky and np.array_equal(g["albedo"].view(np.uint32), als[py, px, :3].view(np.uint32)), where + ("albedo / sky",)
                    assert F(g["depth"]).view(np.uint32) == nd[py, px, 3].view(np.uint32) and (g["obj"], g["sub"]) == tuple(prim[py, px]), where + ("depth / ids",)
            assert lt.rays == r.stats()["rays"], (scene, frame, "Scene.Hit invocations")
    r.close()
    s.close()


def test_library_builds_the_same_trees_as_the_host():
    """ycge_scene_upload without a host tree / ycge_mesh_upload_triangles: the library's own builder (BVH.cs:258-459,
    MeshBVH.cs:371-576) must give the same frame as the host's uploaded trees."""
    s = api.HostScene("knot:40x12")
    lib = api.load_lib()
    r1 = api.CudaRaytraceRenderer(s, 40, 12, 2)
    r1.SetCamera(*api.BENCH_POSE)
    a = r1.TryFlipAndBlit().copy()
    # second context fed with plain triangles and no top-level tree
    cfg = api.Config()
    cfg.fb_w, cfg.fb_h, cfg.ss = 40, 12, 2
    lib.ycge_default_params(C.byref(cfg.params))
    ctx = C.c_void_p()
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == 0
    tris = s.mesh_triangles(0)
    m = s.mesh(0).contents
    assert lib.ycge_mesh_upload_triangles(ctx, 0, len(tris), tris.ctypes.data, C.byref(m.material)) == 0
    s2 = api.Scene()
    C.memmove(C.byref(s2), s.flat, C.sizeof(api.Scene))
    s2.bvh = None
    assert lib.ycge_scene_upload(ctx, C.byref(s2)) == 0
    pos = (C.c_float * 3)(*api.BENCH_POSE[0])
    lib.ycge_set_camera(ctx, pos, api.BENCH_POSE[1], api.BENCH_POSE[2])
    b = np.empty((12, 40), api.CELL_DTYPE)
    assert lib.ycge_render_frame(ctx, b.ctypes.data, 0) == 0
    assert_cells_equal(a, b, "library-built trees")
    lib.ycge_destroy(ctx)
    r1.close()


# ------------------------------------------------------------------------------------------- RNG known answers on the device
def test_rng_known_answers_on_device():
    s = api.HostScene("cornell")
    r = api.CudaRaytraceRenderer(s, 4, 2, 1)
    kat = [((0, 0, 1), 0x17EF7D0094EB2C76, 7379119), ((1, 0, 1), 0xE30B62FE1AC2EDC5, 1918614), ((0, 1, 1), 0x7EDFB4004F82140E, 13519905),
           ((959, 539, 1), 0xC3C5BBAF2004442A, 15095364), ((1919, 1079, 64), 0x5BEAD3AD13E75BBB, 9491773)]
    xs, ys, fs = zip(*[k[0] for k in kat])
    bits, seeds = r.rng_kat(0, xs, ys, fs, 4)
    for i, (_, seed, m24) in enumerate(kat):
        assert int(seeds[i]) == seed
        exp = np.float32(np.float32(m24) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
        assert int(bits[i, 0]) == int(np.float32(exp).view(np.uint32))
    # both streams against the oracle over random inputs
    from oracle_binding import load_oracle
    lib = load_oracle()
    rng = np.random.default_rng(5)
    xs, ys, fs = rng.integers(0, 4000, 64), rng.integers(0, 4000, 64), rng.integers(1, 1 << 33, 64)
    bits, seeds = r.rng_kat(0, xs, ys, fs, 8)
    for i in range(64):
        sd = lib.yo_per_frame_seed(int(xs[i]), int(ys[i]), int(fs[i]), 0, 0, 0x9E3779B97F4A7C15)
        assert int(seeds[i]) == sd
        b = np.zeros(8, np.uint32)
        lib.yo_rng_draws(sd, 8, b.ctypes.data, None)
        assert np.array_equal(b, bits[i])
    bits, seeds = r.rng_kat(1, xs, ys, fs, 8)  # ConsoleRayTracing.Rng (Rng.cs)
    for i in range(64):
        b = np.zeros(8, np.uint32)
        lib.yo_rng_cs_draws(int(seeds[i]), 8, b.ctypes.data)
        assert np.array_equal(b, bits[i])
    r.close()


# ------------------------------------------------------------------------------------------- renderer state machine
def test_camera_motion_reset_resize_lights_and_globals():
    s = api.HostScene("mirror_spheres")
    r = api.CudaRaytraceRenderer(s, 32, 10, 2)
    o = Oracle(s, 32, 10, 2)
    pos, yaw, pitch, fov = s.default_camera()

    def both(fn):
        fn(r)
        fn(o)

    def frame(tag):
        g = r.TryFlipAndBlit()
        c = o.render_frame(threads=4, fast_post=True)
        assert_cells_equal(g, c, tag)
        assert_frame_parity(r, o, tag)

    frame("frame 1")
    frame("frame 2 (history blend)")
    r.SetCamera((pos[0] + 0.5, pos[1] + 0.1, pos[2]), yaw + 0.2, pitch - 0.05); o.set_camera((pos[0] + 0.5, pos[1] + 0.1, pos[2]), yaw + 0.2, pitch - 0.05)
    frame("camera moved: history reset")
    r.SetCamera((pos[0] + 0.501, pos[1] + 0.1, pos[2]), yaw + 0.2, pitch - 0.05); o.set_camera((pos[0] + 0.501, pos[1] + 0.1, pos[2]), yaw + 0.2, pitch - 0.05)
    frame("sub-threshold motion: no reset, history is blended with a shifted frame")
    r.SetFov(60.0); o.set_fov(60.0)
    frame("fov change")
    lights = [((0.0, 3.0, 1.0), (1.0, 0.5, 0.25), 30.0), ((-2.0, 2.0, -1.0), (0.2, 0.4, 1.0), 12.0), ((2.0, 4.0, 2.0), (1.0, 1.0, 1.0), 5.0)]
    both(lambda x: x.lights_update(lights))
    frame("lights_update (3 lights)")
    both(lambda x: x.globals_update((0.1, 0.2, 0.6), (0.7, 0.8, 0.9), (1.0, 0.9, 0.8), 0.15))
    frame("globals_update (sky + ambient)")
    r.Resize(20, 7, 3); o.resize(20, 7, 3)
    frame("after Resize (frame counter and exposure survive)")
    assert r.stats()["frames"] == o.stats()["frames"] == 8
    both(lambda x: x.reset_history())
    frame("reset_history")
    r.close()
    o.close()


def test_async_path_equals_synchronous_path():
    s = api.HostScene("boxes")
    a = api.CudaRaytraceRenderer(s, 40, 12, 2)
    b = api.CudaRaytraceRenderer(s, 40, 12, 2)
    for _ in range(5):
        last = a.TryFlipAndBlit()
    b.render_frames_async(5)
    b.wait()
    assert_cells_equal(last, b.read_cells(), "render_frames_async(5)")
    assert a.stats()["rays_total"] == b.stats()["rays_total"]
    a.close()
    b.close()


@pytest.mark.parametrize("n_slots", [2, 3, 5])
@pytest.mark.parametrize("scene,fb_w,fb_h,ss,pose", [("boxes", 40, 12, 2, None), ("knot:60x16", 48, 27, 4, api.BENCH_POSE)])
def test_frames_pipelined_on_one_gpu_equal_serial_frames(scene, fb_w, fb_h, ss, pose, n_slots):
    """ycge_pipeline_config: up to n_slots frames in flight (their à-trous passes overlap the next frames' trace / TAA).
    Every frame's cells, the last frame's intermediate planes and the exposure recursion must be bit-identical to the
    strictly serial schedule, and the two schedules can be mixed and re-configured between frames."""
    s = api.HostScene(scene)
    a = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    b = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    if pose is not None:
        a.SetCamera(*pose)
        b.SetCamera(*pose)
    b.TryFlipAndBlit()                      # frame 1 on the serial schedule, before the slots exist
    b.pipeline_config(n_slots)
    serial = [a.TryFlipAndBlit().copy() for _ in range(12)]
    # streaming: every frame's cells land in host memory, n_slots frames in flight
    bufs = [np.empty((fb_h, fb_w), api.CELL_DTYPE) for _ in range(n_slots)]
    ids = []
    for f in range(2, 9):
        if len(ids) == n_slots:
            fid = ids.pop(0)
            b.frame_wait(fid)
            assert_cells_equal(bufs[fid % n_slots], serial[fid - 1], f"streamed frame {fid}")
        ids.append(b.submit_frame(bufs[f % n_slots]))
        assert ids[-1] == f
    for fid in ids:
        b.frame_wait(fid)
        assert_cells_equal(bufs[fid % n_slots], serial[fid - 1], f"streamed frame {fid}")
    with pytest.raises(api.YcgeError):
        b.frame_wait(3)                      # already waited for
    # headless: frames 9..11 enqueued at once
    b.render_frames_async(3)
    b.wait()
    assert_cells_equal(b.read_cells(), serial[10], "pipelined render_frames_async")
    # the serial renderer is at frame 12, the pipelined one at 11: re-configure, then one synchronous frame
    b.pipeline_config(1 if n_slots != 3 else 4)
    assert_cells_equal(b.TryFlipAndBlit(), serial[11], "synchronous frame after pipelined frames")
    for kind, nm in [(api.DBG_HDR, "hdr"), (api.DBG_ALBEDO_SKY, "albedo/sky"), (api.DBG_NORMAL_DEPTH, "normal/depth"), (api.DBG_TAA, "taa"),
                     (api.DBG_DENOISED, "denoised"), (api.DBG_LOG_SAMPLES, "log samples"), (api.DBG_PRIM_ID, "prim")]:
        assert bits_differ(a.debug_read(kind), b.debug_read(kind)) == 0, f"{nm} differs after pipelined frames"
    sa, sb = a.stats(), b.stats()
    assert sa["rays_total"] == sb["rays_total"] and sa["frames"] == sb["frames"] == 12
    assert np.float32(sa["ae_exposure"]).view(np.uint32) == np.float32(sb["ae_exposure"]).view(np.uint32)
    a.close()
    b.close()


def test_pipelined_frames_follow_a_moving_camera():
    """Camera motion inside a pipelined stream: the history reset decision (TemporalAA.cs:58-67) is taken per submitted
    frame from the camera latched at submission, exactly as the serial path takes it."""
    s = api.HostScene("cylinders_disks_triangles")
    a = api.CudaRaytraceRenderer(s, 36, 10, 2)
    b = api.CudaRaytraceRenderer(s, 36, 10, 2)
    b.pipeline_config(3)
    poses = [((0.0, 1.0, 0.0), 0.0, 0.0)] * 3 + [((0.0, 1.0, 0.001), 0.0, 0.0)] * 2 + [((0.2, 1.0, 0.0), 0.1, -0.05)] * 3
    bufs = [np.empty((10, 36), api.CELL_DTYPE) for _ in poses]
    ids = []
    want = []
    for k, p in enumerate(poses):
        a.SetCamera(*p)
        want.append(a.TryFlipAndBlit().copy())
        b.SetCamera(*p)
        if len(ids) == 3:
            b.frame_wait(ids.pop(0))
        ids.append(b.submit_frame(bufs[k]))
    for fid in ids:
        b.frame_wait(fid)
    for k in range(len(poses)):
        assert_cells_equal(bufs[k], want[k], f"moving camera, frame {k + 1}")
    a.close()
    b.close()


def test_ansi_stream_through_the_host_framebuffer():
    s = api.HostScene("cornell")
    a = api.CudaRaytraceRenderer(s, 30, 10, 1)
    b = api.CudaRaytraceRenderer(s, 30, 10, 1)
    stream = a.blit_ansi()          # TryFlipAndBlit(fb) -> Chexels -> ANSITerminalRenderer.Render
    cells = b.TryFlipAndBlit()
    assert stream == api.ansi_from_cells(cells)
    assert stream.count(b"\xe2\x96\x80") == 300  # U+2580 per cell
    a.close()
    b.close()


def test_texture_upload_through_the_c_abi_and_errors():
    """ycge_texture_upload by hand: a material that names a texture that was never uploaded is an error; replacing the
    pixels changes the picture; the scene's own textures give the oracle's picture."""
    lib = api.load_lib()
    s = api.HostScene("texture_test")
    cfg = api.Config()
    cfg.fb_w, cfg.fb_h, cfg.ss = 24, 8, 2
    lib.ycge_default_params(C.byref(cfg.params))
    ctx = C.c_void_p()
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == 0
    assert lib.ycge_scene_upload(ctx, s.flat) != 0
    assert b"texture" in lib.ycge_last_error(ctx)
    t = s.texture(0)
    assert lib.ycge_texture_upload(ctx, 0, t.shape[1], t.shape[0], t.ctypes.data) == 0
    assert lib.ycge_texture_upload(ctx, -1, 1, 1, t.ctypes.data) != 0
    assert lib.ycge_scene_upload(ctx, s.flat) == 0
    pose = ((0.6, 0.4, 0.0), 0.25, -0.2)
    lib.ycge_set_camera(ctx, (C.c_float * 3)(*pose[0]), pose[1], pose[2])
    a = np.empty((8, 24), api.CELL_DTYPE)
    assert lib.ycge_render_frame(ctx, a.ctypes.data, 0) == 0
    o = Oracle(s, 24, 8, 2)
    o.set_camera(*pose)
    assert_cells_equal(a, o.render_frame(threads=2), "hand-uploaded texture")
    white = np.full((3, 5), 0xFFFFFFFF, np.uint32)
    assert lib.ycge_texture_upload(ctx, 0, 5, 3, white.ctypes.data) == 0
    assert lib.ycge_scene_upload(ctx, s.flat) == 0
    lib.ycge_reset_history(ctx)
    b = np.empty((8, 24), api.CELL_DTYPE)
    assert lib.ycge_render_frame(ctx, b.ctypes.data, 0) == 0
    assert (a["fg"] != b["fg"]).any()
    lib.ycge_destroy(ctx)
    o.close()
    s.close()


def test_non_default_params_through_the_config_struct():
    """ycge_params carries the reference's compile-time constants (RaytraceRenderer.cs:31-43,:65,:222; ToneMapper.cs:8-21) as data.
    Other values must act the same on both sides: the mirror threshold lowered so that the 0.6 / 0.85 spheres really mirror, four
    mirror bounces (BASELINE config 2 as its text describes it -- not reachable with the reference's constants, hence no parity
    claim against the engine, only GPU = oracle), two diffuse bounces, two à-trous iterations, another TAA alpha and exposure."""
    lib = api.load_lib()
    s = api.HostScene("mirror_spheres")
    cfg = api.Config()
    cfg.fb_w, cfg.fb_h, cfg.ss = 40, 12, 2
    lib.ycge_default_params(C.byref(cfg.params))
    P = cfg.params
    P.mirror_threshold, P.max_mirror_bounces, P.diffuse_bounces, P.atrous_iterations = 0.5, 4, 2, 2
    P.taa_alpha, P.ae_key, P.ae_speed, P.saturation, P.tone_gamma, P.diffuse_sigma_deg = 0.2, 0.25, 0.5, 1.3, 2.0, 10.0
    ctx = C.c_void_p()
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == 0
    assert lib.ycge_scene_upload(ctx, s.flat) == 0
    o = Oracle(s, 40, 12, 2, params=P)
    d = Oracle(s, 40, 12, 2)
    for f in range(3):
        g = np.empty((12, 40), api.CELL_DTYPE)
        assert lib.ycge_render_frame(ctx, g.ctypes.data, 0) == 0
        c = o.render_frame(threads=4)
        assert_cells_equal(g, c, f"non-default params, frame {f + 1}")
        ref_default = d.render_frame(threads=4)
    st = api.Stats()
    assert lib.ycge_get_stats(ctx, C.byref(st)) == 0
    assert st.rays == o.stats()["rays"] and st.rays > d.stats()["rays"], "more bounces trace more rays"
    assert np.float32(st.ae_exposure).view(np.uint32) == np.float32(o.stats()["ae_exposure"]).view(np.uint32)
    assert (g["fg_ansi"] != ref_default["fg_ansi"]).any(), "the parameters must change the picture"
    lib.ycge_destroy(ctx)
    o.close(); d.close(); s.close()


def test_errors_are_loud():
    lib = api.load_lib()
    cfg = api.Config()
    cfg.fb_w, cfg.fb_h, cfg.ss = 8, 4, 1
    lib.ycge_default_params(C.byref(cfg.params))
    ctx = C.c_void_p()
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == 0
    out = np.empty((4, 8), api.CELL_DTYPE)
    assert lib.ycge_render_frame(ctx, out.ctypes.data, 0) == -3  # YCGE_ERR_NO_SCENE ("Scene BVH not built", Scene.cs:73)
    assert b"BVH not built" in lib.ycge_last_error(ctx)
    assert lib.ycge_resize(ctx, 0, 4, 1) == -1
    assert lib.ycge_frame_finish(ctx) == -1
    lib.ycge_destroy(ctx)
    cfg.device = 99
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == -1
    cfg.device = 0
    cfg.fb_w = 0
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == -1


def test_bad_caller_input_is_rejected_not_executed():
    """negative counts, NULL arrays behind positive counts, a palette default outside the material table and bounce
    limits beyond the device's branch stack must come back as a status, never as undefined behaviour or a C++ exception."""
    lib = api.load_lib()
    cfg = api.Config()
    cfg.fb_w, cfg.fb_h, cfg.ss = 8, 4, 1
    lib.ycge_default_params(C.byref(cfg.params))
    ctx = C.c_void_p()
    cfg.params.max_mirror_bounces = 16
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == -5 and b"max_mirror_bounces" in lib.ycge_last_error(None)
    cfg.params.max_mirror_bounces = 2
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == 0
    s = api.HostScene("volume_grid_test")
    v = s.volume(0).contents
    bad = api.Volume.from_buffer_copy(v)
    bad.palette_n_ids = -3
    assert lib.ycge_volume_upload(ctx, 0, C.byref(bad)) == -1
    ok = api.Volume.from_buffer_copy(v)
    ok.palette_default = 200  # inside [0, 254] but outside this scene's material table
    assert lib.ycge_volume_upload(ctx, 0, C.byref(ok)) == 0
    assert lib.ycge_scene_upload(ctx, s.flat) == -1 and b"default" in lib.ycge_last_error(ctx)
    assert lib.ycge_volume_upload(ctx, 0, s.volume(0)) == 0
    sc = api.Scene.from_buffer_copy(s.flat.contents)
    sc.lights = None
    assert sc.n_lights > 0 and lib.ycge_scene_upload(ctx, C.byref(sc)) == -1
    m = api.MeshSoa()
    m.n_tris = -1
    m.bvh = C.pointer(api.Bvh())
    assert lib.ycge_mesh_upload_soa(ctx, 0, C.byref(m)) == -1
    m.n_tris = 4
    assert lib.ycge_mesh_upload_soa(ctx, 0, C.byref(m)) == -1  # NULL triangle arrays
    assert lib.ycge_set_inplace_variant(ctx, 7) == -1
    assert lib.ycge_scene_upload(ctx, s.flat) == 0             # and the context is still usable
    out = np.empty((4, 8), api.CELL_DTYPE)
    assert lib.ycge_render_frame(ctx, out.ctypes.data, 0) == 0
    lib.ycge_destroy(ctx)
    s.close()


# ------------------------------------------------------------------------------------------- BASELINE sizes
def test_full_size_1080p_dragon_two_frames_vs_oracle():
    """The north-star workload at its real size (1920x1080 internal = 480x135 cells, ss = 4): two frames (the second
    blends TAA history) against the oracle on the box's host cores."""
    s = api.HostScene("dragon")
    r = api.CudaRaytraceRenderer(s, 480, 135, 4)
    o = Oracle(s, 480, 135, 4)
    r.SetCamera(*api.BENCH_POSE)
    o.set_camera(*api.BENCH_POSE)
    for f in range(2):
        g = r.TryFlipAndBlit()
        c = o.render_frame(threads=os.cpu_count() or 1, fast_post=True)
        diff = sum((g[k] != c[k]).any() for k in CELL_KEYS)
        assert_cells_equal(g, c, f"dragon 1080p frame {f + 1}")
        assert_frame_parity(r, o, f"dragon 1080p frame {f + 1}")
        assert diff == 0
    r.close()
    o.close()


def test_full_size_mirror_spheres_1080p_vs_oracle():
    """BASELINE config 2: mirror spheres on checker plane, 1920x1080 internal, one frame, reference constants."""
    s = api.HostScene("mirror_spheres")
    r = api.CudaRaytraceRenderer(s, 480, 135, 4)
    o = Oracle(s, 480, 135, 4)
    g = r.TryFlipAndBlit()
    c = o.render_frame(threads=os.cpu_count() or 1, fast_post=True)
    assert_cells_equal(g, c, "mirror spheres 1080p")
    assert_frame_parity(r, o, "mirror spheres 1080p")
    r.close()
    o.close()


def _fullsize():
    import json
    return json.load(open(os.path.join(GOLDEN, "fullsize_hashes.json")))


def _frame_digest(r, cells):
    import make_golden_fullsize
    return make_golden_fullsize.frame_digest(cells, r.debug_read(api.DBG_PRIM_ID), r.debug_read(api.DBG_HDR), r.debug_read(api.DBG_TAA),
                                            r.debug_read(api.DBG_DENOISED), r.stats())


FULLSIZE = ["c3_cylinders_disks_triangles", "c3_boxes", "c3_bunny", "c3_teapot", "c4_voxel_world_1440p", "c4_voxel_island_1440p", "c5_dragon_2160p", "c5_dragon_1080p"]


@pytest.mark.parametrize("name", FULLSIZE)
def test_full_size_vs_oracle_hashes(name):
    """BASELINE.json's configurations at their FULL sizes against the CPU oracle: the oracle rendered them once on the CPU
    (tools/make_golden_fullsize.py -> tests/golden/fullsize_hashes.json: SHA-256 of the cells, the primary ids, the HDR / TAA /
    denoised planes, plus ray count and exposure scalars); the GPU must reproduce every digest.  C3: showcase scenes and meshes
    at 1920x1080 with TAA accumulated over 64 frames; C4: both voxel worlds at 2560x1440; C5: the dragon at 3840x2160."""
    import make_golden_fullsize
    gold = _fullsize().get(name)
    if gold is None:
        pytest.fail(f"{name} is missing from tests/golden/fullsize_hashes.json: run tools/make_golden_fullsize.py {name}")
    scene, fb_w, fb_h, ss, frames, pose = make_golden_fullsize.CASES[name]
    s = api.HostScene(scene)
    assert s.name == gold["scene"] and s.counts()["triangles"] == gold["triangles"]
    r = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    if pose is not None:
        r.SetCamera(*pose)
    for f in range(1, max(frames) + 1):
        if f in frames:
            cells = r.TryFlipAndBlit()
            got, want = _frame_digest(r, cells), gold["frames"][str(f)]
            bad = [k for k in want if got[k] != want[k]]
            assert not bad, f"{name} frame {f}: {bad} differ from the oracle"
        else:
            r.TryFlipAndBlit()
    r.close()
    s.close()


def test_full_size_properties_voxel_world_1440p():
    """Size-independent properties at config 4's size (the digests above pin the values): deterministic across contexts, frame 1
    history == current HDR, sky mask consistent with primary ids, all cells inside the ANSI cube, exposure inside its clamp."""
    s = api.HostScene("voxel_world")
    a = api.CudaRaytraceRenderer(s, 320, 90, 8)
    b = api.CudaRaytraceRenderer(s, 320, 90, 8)
    ca, cb = a.TryFlipAndBlit(), b.TryFlipAndBlit()
    assert_cells_equal(ca, cb, "determinism")
    hdr, taa = a.debug_read(api.DBG_HDR), a.debug_read(api.DBG_TAA)
    assert bits_differ(hdr, taa) == 0
    prim, sky = a.debug_read(api.DBG_PRIM_ID), a.debug_read(api.DBG_ALBEDO_SKY)[..., 3]
    assert np.array_equal(prim[..., 0] < 0, sky != 0)
    assert 0.05 < (sky != 0).mean() < 0.95
    assert ca["fg_ansi"].min() >= 16 and ca["fg_ansi"].max() <= 231 and np.all(ca["glyph"] == 0x2580)
    assert 0.10 <= a.stats()["ae_exposure"] <= 1.50
    a.close()
    b.close()


# ------------------------------------------------------------------------------------------- several GPUs behind one context
@pytest.mark.parametrize("scene,fb_w,fb_h,ss,n,pose", [("boxes", 40, 24, 2, 3, None), ("knot:60x16", 48, 27, 4, 4, api.BENCH_POSE), ("mirror_spheres", 33, 9, 1, 2, None),
                                                       ("voxel_world:64x64", 40, 12, 2, 5, None), ("entities_demo", 40, 16, 2, 3, None)])
def test_multi_device_context_equals_one_device_context(scene, fb_w, fb_h, ss, n, pose):
    """ycge_config.n_devices >= 2 (include/ycge.h): ONE context, the library renders frames in parallel over the listed devices --
    FRONT on balanced row tiles, peer copies into a back slot of the frame's root device, BACK and FINISH there, exposure state
    handed on in frame order.  Here the list names device 0 n times (this box may have one GPU; tests/test_multigpu.py runs it
    over distinct GPUs): frames submitted several at a time, a moving camera, moving lights; every cell equals the one-device frame."""
    s = api.HostScene(scene)
    one = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    many = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss, devices=[0] * n)
    if pose is None:
        pose = s.default_camera()[:3]
    bufs = [np.empty((fb_h, fb_w), api.CELL_DTYPE) for _ in range(2 * n)]
    frame = 0
    for batch in (1, 2 * n, 3, n + 1):
        ids, want = [], []
        for k in range(batch):
            p = ((pose[0][0] + 0.0005 * frame, pose[0][1], pose[0][2] + (0.02 if frame == 5 else 0.0)), pose[1], pose[2])
            if scene == "entities_demo":
                s.update(16.0)
                for r in (one, many):
                    r.SyncLights(s)
                    r.SyncGeometry(s)
            one.SetCamera(*p)
            many.SetCamera(*p)
            want.append(one.TryFlipAndBlit())
            ids.append(many.submit_frame(bufs[k]))
            frame += 1
        for k, fid in enumerate(ids):
            many.frame_wait(fid)
            assert_cells_equal(bufs[k], want[k], f"{scene} x{n}: frame {frame - batch + k + 1}")
    assert_cells_equal(many.TryFlipAndBlit(), one.TryFlipAndBlit(), "the synchronous call on the multi-device context")
    st1, stn = one.stats(), many.stats()
    assert stn["frames"] == st1["frames"] == frame + 1
    assert np.float32(stn["ae_exposure"]).view(np.uint32) == np.float32(st1["ae_exposure"]).view(np.uint32)
    many.close(); one.close(); s.close()


def test_multi_device_context_refuses_what_it_cannot_do():
    lib = api.load_lib()
    cfg = api.Config()
    cfg.fb_w, cfg.fb_h, cfg.ss = 16, 8, 1
    lib.ycge_default_params(C.byref(cfg.params))
    cfg.n_devices = 2
    cfg.devices[0], cfg.devices[1] = 0, 99
    ctx = C.c_void_p()
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == -1
    cfg.devices[1] = 0
    cfg.tile_rows = 4
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == -1
    cfg.tile_rows = 0
    assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == 0
    out = np.empty((8, 16), api.CELL_DTYPE)
    assert lib.ycge_render_frame(ctx, out.ctypes.data, 0) == -3  # no scene yet
    assert lib.ycge_frame_begin(ctx) == -1 and b"multi-GPU" in lib.ycge_last_error(ctx)
    assert lib.ycge_debug_read(ctx, api.DBG_HDR, out.ctypes.data, out.nbytes) == -1
    lib.ycge_destroy(ctx)


# ------------------------------------------------------------------------------------------- row tiles (sharding kernels)
@pytest.mark.parametrize("peers", [False, True], ids=["copy-handoff", "peer-handoff"])
@pytest.mark.parametrize("scene,fb_w,fb_h,ss,n_tiles,pose", [("boxes", 40, 24, 2, 3, None), ("knot:60x16", 48, 27, 4, 4, api.BENCH_POSE),
                                                             ("mirror_spheres", 33, 8, 1, 8, None), ("voxel_world:64x64", 40, 12, 2, 2, None)])
def test_row_tiles_equal_the_unsharded_frame(scene, fb_w, fb_h, ss, n_tiles, pose, peers):
    """N row-tile contexts on ONE GPU, driven through the phase API with a loop-back exchange (device copies instead of
    NCCL): boundary rows of the in-place pass handed tile -> tile, exposure samples summed, cells concatenated.  Must equal
    the unsharded frame bit for bit, over several frames (TAA history and exposure state live per tile)."""
    import torch
    from yetanotherconsolegameengine_b200 import sharding
    s = api.HostScene(scene)
    full = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    tiles = []
    with torch.cuda.device(0):
        for r in range(n_tiles):
            row0, rows = sharding.tile_rows(r, n_tiles, fb_h)
            tiles.append(sharding.CudaTileBackend(s, fb_w, fb_h, ss, row0, rows, 0))
        if peers and min(t.rows for t in tiles) * 2 * ss < 4:
            # a tile shorter than the in-place pass reaches cannot use the peer hand-off: the library says so
            ex = [t.peer_export() for t in tiles]
            with pytest.raises(api.YcgeError):
                tiles[0].peer_attach(None, ex[1], via_ipc=False)
            for t in tiles:
                t.close()
            full.close()
            return
        if peers:  # the wavefront kernels store the boundary rows straight into the tile below (raw pointers: one process)
            ex = [t.peer_export() for t in tiles]
            for i, t in enumerate(tiles):
                t.peer_attach(ex[i - 1] if i > 0 else None, ex[i + 1] if i + 1 < n_tiles else None, via_ipc=False)
        if pose is not None:
            full.SetCamera(*pose)
            for t in tiles:
                t.set_camera(*pose)
        for frame in range(3):
            ref = full.TryFlipAndBlit()
            for t in tiles:
                t.begin()
            while True:
                halos = [t.halo() for t in tiles]
                if halos[0] is None:
                    assert all(h is None for h in halos)
                    break
                prev_send = None
                for t, (recv, send) in zip(tiles, halos):
                    assert not (peers and (recv is not None or send is not None))
                    if recv is not None:
                        assert prev_send is not None and prev_send.numel() == recv.numel()
                        recv.copy_(prev_send)
                    t.inplace()
                    prev_send = send
            total = torch.stack([t.logs for t in tiles]).sum(0)  # each slot is owned by one tile, the others hold 0
            for t in tiles:
                t.logs.copy_(total)
                t.finish()
            torch.cuda.synchronize()
            got = np.concatenate([t.cells.cpu().numpy().view(api.CELL_DTYPE).reshape(t.rows, fb_w) for t in tiles])
            assert_cells_equal(got, ref, f"{scene} frame {frame + 1}, {n_tiles} tiles")
    for t in tiles:
        t.close()
    full.close()


@pytest.mark.parametrize("scene,fb_w,fb_h,ss,n_tiles,pose", [("boxes", 40, 24, 2, 3, None), ("knot:60x16", 48, 27, 4, 4, api.BENCH_POSE), ("voxel_world:64x64", 32, 18, 1, 5, None)])
def test_front_tiles_plus_whole_frame_back_equal_the_unsharded_frame(scene, fb_w, fb_h, ss, n_tiles, pose):
    """Frame-parallel sharding on one GPU: FRONT (trace + TAA) on row-tile contexts with a single halo row, the tiles' rows
    of history + guides copied into a back slot of a whole-frame context, BACK (à-trous passes, exposure samples) and
    FINISH there -- what sharding.FrameParallelRenderer does with NCCL between ranks.  Cells of every frame and the
    exposure recursion must equal the unsharded renderer's bit for bit, with two back slots used alternately."""
    import torch
    from yetanotherconsolegameengine_b200 import sharding
    s = api.HostScene(scene)
    ref = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    back = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    back.back_config(2)
    tiles = [sharding.tile_rows(r, n_tiles, fb_h) for r in range(n_tiles)]
    fronts = [api.CudaRaytraceRenderer(s, fb_w, fb_h, ss, tile_row0=t[0], tile_rows=t[1]) for t in tiles]
    for r in [ref] + fronts:
        if pose is not None:
            r.SetCamera(*pose)
    W, H = fb_w * ss, fb_h * 2 * ss
    nbytes, row_bytes = W * H * 16, W * 16
    kinds = (api.PTR_HIST, api.PTR_GND, api.PTR_GAS)
    st = torch.cuda.Stream()
    for f in range(5):
        slot = f % 2
        dst = [sharding.device_bytes(back.back_ptr(slot, k)[0], nbytes, 0) for k in kinds]
        for fr, (row0, rows) in zip(fronts, tiles):
            fr.frame_front()
            fr.wait()
            a, b = row0 * 2 * ss * row_bytes, (row0 + rows) * 2 * ss * row_bytes
            for k, kind in enumerate(kinds):
                src = sharding.device_bytes(fr.device_ptr(kind)[0], nbytes, 0)
                dst[k][a:b].copy_(src[a:b])
        torch.cuda.synchronize()
        back.back_denoise(slot, st.cuda_stream)
        back.back_finish(slot, st.cuda_stream)
        st.synchronize()
        got = sharding.device_bytes(*back.back_ptr(slot, api.PTR_CELLS), 0).cpu().numpy().view(api.CELL_DTYPE).reshape(fb_h, fb_w)
        assert_cells_equal(got, ref.TryFlipAndBlit(), f"{scene}: frame {f + 1}, {n_tiles} front tiles + whole-frame back")
    e_ref = ref.stats()["ae_exposure"]
    e_got = sharding.device_bytes(*back.device_ptr(api.PTR_EXPOSURE), 0).cpu().numpy().view(np.float32)[0]
    assert np.float32(e_ref).view(np.uint32) == np.float32(e_got).view(np.uint32)
    for r in [ref, back] + fronts:
        r.close()
    s.close()


def test_frame_parallel_renderer_on_one_rank():
    """sharding.FrameParallelRenderer with world = 1: the orchestration (back slots, streams, FINISH chain, batches) without
    the collectives; tools/multigpu_check.py runs it over real ranks."""
    import torch
    from yetanotherconsolegameengine_b200 import sharding
    s = api.HostScene("boxes")
    ref = api.CudaRaytraceRenderer(s, 40, 24, 2)
    with torch.cuda.device(0):
        fp = sharding.FrameParallelRenderer(s, 0, 1, 40, 24, 2, 0, back_slots=2)
        got = fp.render(5, collect=True) + fp.render(3, collect=True)
        torch.cuda.synchronize()
        assert len(got) == 8
        for f, g in enumerate(got):
            assert_cells_equal(fp.cells_host(g), ref.TryFlipAndBlit(), f"frame-parallel frame {f + 1}")
        fp.close()
    ref.close()
    s.close()


def test_stashed_finish_equals_the_synchronous_frame():
    """The frame-pipelining path on one GPU (world = 1): frames are rendered back to back on one stream and finished from
    their stash slots on a side stream (ordered exposure sum + cells), out of step with the rendering.  Same cells."""
    import torch
    from yetanotherconsolegameengine_b200 import sharding
    s = api.HostScene("boxes")
    ref = api.CudaRaytraceRenderer(s, 40, 24, 2)
    with torch.cuda.device(0):
        b = sharding.CudaTileBackend(s, 40, 24, 2, 0, 24, 0)
        sr = sharding.ShardedRenderer(b, 0, 1, 40, 24)
        got = sr.render_pipelined(7, collect=True)
        torch.cuda.synchronize()
        assert len(got) == 7
        for f, g in enumerate(got):
            assert_cells_equal(sr.assemble(g), ref.TryFlipAndBlit(), f"pipelined frame {f + 1}")
        # and the two paths can be mixed: a synchronous frame after pipelined ones continues the same sequence
        assert_cells_equal(sr.TryFlipAndBlit(), ref.TryFlipAndBlit(), "synchronous frame after pipelined frames")
    b.close()
    ref.close()


def test_taa_accumulation_over_64_frames_vs_oracle():
    """BASELINE config 3: a mesh scene with a static camera over 64 frames — the TAA history (alpha = 0.01) and the exposure
    recursion accumulate frame after frame; the 64th frame must still be bit-identical to the oracle's 64th frame."""
    s = api.HostScene("teapot")
    r = api.CudaRaytraceRenderer(s, 40, 12, 2)
    o = Oracle(s, 40, 12, 2)
    r.SetCamera(*api.BENCH_POSE)
    o.set_camera(*api.BENCH_POSE)
    for f in range(64):
        g = r.TryFlipAndBlit()
        c = o.render_frame(threads=os.cpu_count() or 1, fast_post=True)
        if f in (0, 1, 15, 63):
            assert_cells_equal(g, c, f"teapot frame {f + 1}")
            assert_frame_parity(r, o, f"teapot frame {f + 1}")
    assert r.stats()["frames"] == o.stats()["frames"] == 64
    r.close()
    o.close()


def test_4k_internal_resolution_multi_launch_wavefront():
    """BASELINE config 5 resolution (3840x2160 internal = 480x135 cells, ss = 8).  Too slow for the CPU oracle inside a
    test; at this size the wavefront kernel does not fit the GPU in one co-resident launch, so the rows go out in several
    launches.  Properties: two row tiles (loop-back hand-off, each tile fits one launch) assemble to exactly the unsharded
    frame (computed by the multi-launch path), two frames (the second one blends history), deterministic across contexts."""
    import torch
    from yetanotherconsolegameengine_b200 import sharding
    s = api.HostScene("knot:200x40")
    fb_w, fb_h, ss = 480, 135, 8
    full = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    full.SetCamera(*api.BENCH_POSE)
    with torch.cuda.device(0):
        tiles = []
        for k in range(2):
            row0, rows = sharding.tile_rows(k, 2, fb_h)
            t = sharding.CudaTileBackend(s, fb_w, fb_h, ss, row0, rows, 0)
            t.set_camera(*api.BENCH_POSE)
            tiles.append(t)
        for frame in range(2):
            ref = full.TryFlipAndBlit()
            assert ref["fg_ansi"].min() >= 16 and np.all(ref["glyph"] == 0x2580)
            for t in tiles:
                t.begin()
            prev = None
            halos = [t.halo() for t in tiles]
            for t, (recv, send) in zip(tiles, halos):
                if recv is not None:
                    recv.copy_(prev)
                t.inplace()
                prev = send
            assert all(t.halo() is None for t in tiles)
            total = torch.stack([t.logs for t in tiles]).sum(0)
            for t in tiles:
                t.logs.copy_(total)
                t.finish()
            torch.cuda.synchronize()
            got = np.concatenate([t.cells.cpu().numpy().view(api.CELL_DTYPE).reshape(t.rows, fb_w) for t in tiles])
            assert_cells_equal(got, ref, f"4K frame {frame + 1}")
    for t in tiles:
        t.close()
    full.close()


@pytest.mark.parametrize("scene,fb_w,fb_h,ss", [("cornell", 61, 17, 1), ("mirror_spheres", 240, 67, 2), ("voxel_world:64x64", 33, 40, 1), ("boxes", 1, 1, 1)])
def test_ansi_stream_emitted_on_the_device(scene, fb_w, fb_h, ss):
    """ANSITerminalRenderer.Render's byte stream produced by a kernel (row prefix, colour-change escapes,
    UTF-8 glyph, final reset) equals the host mirror's Render over the same cells, byte for byte, frame after frame."""
    s = api.HostScene(scene)
    r = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    for _ in range(3):
        cells = r.TryFlipAndBlit()
        dev = r.ansi_stream()
        host = api.ansi_from_cells(cells)
        assert dev == host, (len(dev), len(host))
    assert dev.startswith(b"\x1b[1;1H") and dev.endswith(b"\x1b[0m")
    r.close()
