"""The oracle against THE REFERENCE'S OWN SOURCE TEXT.

oracle/ref_transpile.py rewrites the C# files under /root/reference into C++ syntactically (no arithmetic expression is touched)
and oracle/Makefile compiles the result into oracle/_ref/libycge_ref.so (git-ignored; nothing generated is committed).  What
runs here is therefore the reference author's code -- TemporalBlendWithClamp, ApplyAtrousDenoise and the verbatim tail of
TryFlipAndBlit with its buffer juggling (RaytraceRenderer.cs:218-264, :274-398, :622-722), ToneMapper.cs, Chexel.cs, the
ANSI-256 quantiser and the byte-stream Render of ANSITerminalRenderer.cs, RaytraceSampler.cs, Vec3.cs -- with one documented substitution: MathF.Exp /
Log / Pow / Sin / Cos forward to include/ycge_detmath.h, as they do in the oracle and the product.  The hand-written
oracle (oracle/ycge_oracle.cpp) must agree with it bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from yetanotherconsolegameengine_b200 import api
from oracle_binding import Oracle

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libycge_ref.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libycge_ref.so is not built (needs /root/reference at build time)")
    lib = C.CDLL(REF_SO)
    vp, f32, u64, i64 = C.c_void_p, C.c_float, C.c_uint64, C.c_int64
    lib.ref_renderer_create.restype = vp
    lib.ref_renderer_create.argtypes = [C.c_int] * 4
    lib.ref_renderer_destroy.argtypes = [vp]
    lib.ref_post_frame.argtypes = [vp] * 6 + [C.c_int] + [vp] * 10
    lib.ref_per_frame_seed.restype = u64
    lib.ref_per_frame_seed.argtypes = [C.c_int, C.c_int, i64, C.c_int, C.c_int, u64]
    lib.ref_splitmix64.restype = u64
    lib.ref_splitmix64.argtypes = [u64]
    lib.ref_rng_draws.argtypes = [u64, C.c_int, vp]
    lib.ref_blue_noise.restype = f32
    lib.ref_blue_noise.argtypes = [C.c_int] * 4
    lib.ref_cosine_sample.argtypes = [f32, f32, f32, u64, vp]
    lib.ref_ansi256.argtypes = [f32, f32, f32]
    lib.ref_nearest16.argtypes = [f32, f32, f32]
    lib.ref_linear_to_srgb8.argtypes = [C.c_double]
    lib.ref_ansi_render.argtypes = [C.c_int, C.c_int, vp, vp, vp] + [C.c_int] * 6 + [vp, C.c_int]
    return lib


def P(a):
    return C.c_void_p(a.ctypes.data)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


CASES = [("cornell", 24, 10, 2, None), ("mirror_spheres", 30, 9, 2, None), ("knot:40x10", 24, 9, 3, api.BENCH_POSE), ("voxel_world:64x64", 20, 8, 2, None),
         ("boxes", 17, 5, 1, None), ("cylinders_disks_triangles", 16, 6, 4, None)]


@pytest.mark.parametrize("scene,fb_w,fb_h,ss,pose", CASES, ids=[c[0] for c in CASES])
def test_image_passes_equal_the_reference_source(ref, scene, fb_w, fb_h, ss, pose):
    """Five frames (history reset on frame 1 and, forced, on frame 4): the oracle's trace planes go into the TRANSPILED tail of
    TryFlipAndBlit; its TAA history, denoised image (three a-trous iterations, the second one in place -- nobody wrote that down,
    it follows from the reference's own `dst = (tmp == scratchA) ? scratchB : scratchA`), exposure and every cell field must
    equal the oracle's."""
    s = api.HostScene(scene)
    o = Oracle(s, fb_w, fb_h, ss)
    if pose is not None:
        o.set_camera(*pose)
    W, H = fb_w * ss, fb_h * 2 * ss
    h = ref.ref_renderer_create(fb_w, fb_h, ss, 3)
    taa, den = np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.float32)
    expo = np.empty(2, np.float32)
    glyph = np.empty((fb_h, fb_w), np.uint16)
    fg16, bg16, fga, bga = (np.empty((fb_h, fb_w), np.uint8) for _ in range(4))
    fg, bg = np.empty((fb_h, fb_w, 3), np.float32), np.empty((fb_h, fb_w, 3), np.float32)
    for frame in range(1, 6):
        if frame == 4:
            o.reset_history()
        cells = o.render_frame(threads=2)
        hdr = np.ascontiguousarray(o.debug_read(api.DBG_HDR)[..., :3])
        asky = o.debug_read(api.DBG_ALBEDO_SKY)
        alb, sky = np.ascontiguousarray(asky[..., :3]), np.ascontiguousarray((asky[..., 3] != 0).astype(np.uint8))
        nrm = o.raw_normal()
        dep = np.ascontiguousarray(o.debug_read(api.DBG_NORMAL_DEPTH)[..., 3])
        rc = ref.ref_post_frame(h, P(hdr), P(alb), P(nrm), P(dep), P(sky), 1 if frame in (1, 4) else 0, P(taa), P(den), P(expo), P(glyph), P(fg16), P(bg16), P(fga), P(bga), P(fg), P(bg))
        assert rc == 0
        what = f"{scene} frame {frame}"
        assert np.array_equal(bits(taa), bits(o.debug_read(api.DBG_TAA)[..., :3])), what + ": TAA history"
        assert np.array_equal(bits(den), bits(o.debug_read(api.DBG_DENOISED)[..., :3])), what + ": denoised"
        assert bits(expo[:1])[0] == bits(np.float32(o.stats()["ae_exposure"]))[()], what + ": aeExposure"
        assert np.array_equal(glyph, cells["glyph"]) and np.array_equal(fg16, cells["fg16"]) and np.array_equal(bg16, cells["bg16"]), what + ": cells"
        assert np.array_equal(fga, cells["fg_ansi"]) and np.array_equal(bga, cells["bg_ansi"]), what + ": ANSI-256"
        attr = np.array([[ref.ref_map_attributes(int(a), int(b)) for a, b in zip(ra, rb)] for ra, rb in zip(fg16, bg16)], np.uint16)
        assert np.array_equal(attr, cells["attr"]), what + ": Win32 attribute word (Win32TerminalRenderer.MapAttributes)"
        assert np.array_equal(bits(fg), bits(cells["fg"])) and np.array_equal(bits(bg), bits(cells["bg"])), what + ": SDR colours"
    ref.ref_renderer_destroy(h)
    o.close()
    s.close()


def test_sampler_equals_the_reference_source(ref, oracle_lib):
    rng = np.random.default_rng(5)
    for x, y, f in zip(rng.integers(0, 4000, 300), rng.integers(0, 2200, 300), rng.integers(1, 1 << 40, 300)):
        a = ref.ref_per_frame_seed(int(x), int(y), int(f), 0, 0, 0x9E3779B97F4A7C15)
        assert a == oracle_lib.yo_per_frame_seed(int(x), int(y), int(f), 0, 0, 0x9E3779B97F4A7C15)
        assert ref.ref_splitmix64(a) == oracle_lib.yo_splitmix64(a)
        d = np.empty(16, np.float32)
        ref.ref_rng_draws(a, 16, P(d))
        b, m = np.empty(16, np.uint32), np.empty(16, np.uint32)
        oracle_lib.yo_rng_draws(C.c_uint64(a), 16, P(b), P(m))
        assert np.array_equal(d.view(np.uint32), b)
    assert ref.ref_per_frame_seed(0, 0, 1, 0, 0, 0x9E3779B97F4A7C15) == 0x17EF7D0094EB2C76
    assert ref.ref_per_frame_seed(1919, 1079, 64, 0, 0, 0x9E3779B97F4A7C15) == 0x5BEAD3AD13E75BBB
    # Rng.cs (ConsoleRayTracing.Rng; SURVEY 8 a3'): the reference's text against the oracle's RngCs (what the device's rng_kat_kernel equals)
    ref.ref_rng_cs_draws.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
    for seed in (0, 1, 12345, 0xDEADBEEFCAFEBABE, 0xFFFFFFFFFFFFFFFF):
        a, b = np.zeros(64, np.uint32), np.zeros(64, np.uint32)
        ref.ref_rng_cs_draws(seed, 64, P(a))
        oracle_lib.yo_rng_cs_draws(seed, 64, P(b))
        assert np.array_equal(a, b), seed
    oracle_lib.yo_blue_noise.restype = C.c_float
    for x in range(0, 40, 3):
        for y in range(0, 17, 2):
            for fi in (0, 1, 7, 123456):
                for ch in (0, 1):
                    assert np.float32(ref.ref_blue_noise(x, y, fi, ch)).view(np.uint32) == np.float32(oracle_lib.yo_blue_noise(x, y, fi, ch)).view(np.uint32)
    oracle_lib.yo_cosine_sample.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_void_p]
    for k in range(200):
        n = rng.normal(size=3)
        n = (n / np.linalg.norm(n)).astype(np.float32)
        if k == 0:
            n = np.array([0.0, 0.0, -1.0], np.float32)  # the wz < -0.999999 branch
        a, b = np.empty(3, np.float32), np.empty(3, np.float32)
        seed = int(rng.integers(1, 1 << 62))
        ref.ref_cosine_sample(float(n[0]), float(n[1]), float(n[2]), seed, P(a))
        oracle_lib.yo_cosine_sample(float(n[0]), float(n[1]), float(n[2]), seed, P(b))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_quantisers_equal_the_reference_source(ref, oracle_lib):
    oracle_lib.yo_ansi256.argtypes = [C.c_float] * 3
    oracle_lib.yo_nearest16.argtypes = [C.c_float] * 3
    oracle_lib.yo_linear_to_srgb8.argtypes = [C.c_double]
    rng = np.random.default_rng(11)
    cols = np.concatenate([rng.random((4000, 3)), rng.random((500, 1)).repeat(3, 1), np.array([[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1.5, -0.2, 0.3]])]).astype(np.float32)
    for r, g, b in cols:
        assert ref.ref_ansi256(float(r), float(g), float(b)) == oracle_lib.yo_ansi256(float(r), float(g), float(b))
        assert ref.ref_nearest16(float(r), float(g), float(b)) == oracle_lib.yo_nearest16(float(r), float(g), float(b))
    for c in np.linspace(-0.1, 1.1, 5001):
        assert ref.ref_linear_to_srgb8(float(c)) == oracle_lib.yo_linear_to_srgb8(float(c))


def ref_ansi_stream(ref, cells, vx=0, vy=0, console=None, known=None):
    """ANSITerminalRenderer.Render of the transpiled reference over one Framebuffer holding `cells` (glyph + float colours)."""
    fb_h, fb_w = cells.shape
    console = console or (fb_w, fb_h)
    known = known or console
    glyph = np.ascontiguousarray(cells["glyph"], np.uint16)
    fg = np.ascontiguousarray(cells["fg"], np.float32)
    bg = np.ascontiguousarray(cells["bg"], np.float32)
    cap = 64 + console[0] * console[1] * 32
    buf = np.empty(cap, np.uint8)
    n = ref.ref_ansi_render(fb_w, fb_h, P(glyph), P(fg), P(bg), vx, vy, console[0], console[1], known[0], known[1], P(buf), cap)
    assert n >= 0
    return buf[:n].tobytes()


def test_ansi_byte_stream_equals_the_reference_source(ref):
    """the bytes the terminal receives.  ANSITerminalRenderer.Render (:86-153) with GetChexelForPoint, AppendInt,
    AppendCharUtf8 ... runs as the reference's own text; the host mirror's Render (what the device's ycge_ansi_emit is compared
    with on the GPU) must produce the same stream from the same cells: cursor addressing per row, colour runs (both / fg only /
    bg only -- the first `if` needs BOTH to differ, :120), 1-, 2- and 3-byte glyphs, the final reset; a console whose size
    changed gets the clear-screen prologue in front."""
    rng = np.random.default_rng(5)
    for fb_w, fb_h in [(31, 9), (1, 1), (120, 33)]:
        cells = np.zeros((fb_h, fb_w), api.CELL_DTYPE)
        cells["glyph"] = 0x2580
        cells["fg"] = rng.random((fb_h, fb_w, 3)).astype(np.float32)
        cells["bg"] = rng.random((fb_h, fb_w, 3)).astype(np.float32)
        if fb_h > 2:
            cells["fg"][2, :] = (0.1, 0.7, 0.3)   # runs: no escape while nothing changes, then bg-only changes
            cells["bg"][2, 5:] = (0.9, 0.2, 0.2)
            cells["bg"][3, :] = (0.2, 0.2, 0.9)   # fg-only changes
            cells["glyph"][0, 0] = ord("A")
            cells["glyph"][0, 1] = 0x00E9
        for y in range(fb_h):
            for x in range(fb_w):
                cells["fg_ansi"][y, x] = ref.ref_ansi256(*map(float, cells["fg"][y, x]))
                cells["bg_ansi"][y, x] = ref.ref_ansi256(*map(float, cells["bg"][y, x]))
        want = ref_ansi_stream(ref, cells)
        assert want.startswith(b"\x1b[1;1H") and want.endswith(b"\x1b[0m")
        assert api.ansi_from_cells(cells) == want, (fb_w, fb_h)
        # the renderer remembers another console size: onResize, then ESC[2J ESC[H in front of the same stream (:88-103)
        assert ref_ansi_stream(ref, cells, known=(fb_w + 1, fb_h)) == b"\x1b[2J\x1b[H" + want
    # a framebuffer smaller than the console, placed at (2, 1): what it does not cover is a space in the default colours
    # (ConsoleColor.Black on Black here), and so is a cell whose glyph is a space (GetChexelForPoint :67-84)
    cells = np.zeros((2, 3), api.CELL_DTYPE)
    cells["glyph"] = 0x2580
    cells["glyph"][1, 1] = ord(" ")
    cells["fg"][:] = (1.0, 1.0, 1.0)
    cells["bg"][:] = (1.0, 0.0, 0.0)
    k = ref.ref_ansi256(0.0, 0.0, 0.0)
    f, b = ref.ref_ansi256(1.0, 1.0, 1.0), ref.ref_ansi256(1.0, 0.0, 0.0)
    got = ref_ansi_stream(ref, cells, vx=2, vy=1, console=(6, 4))
    blk = "\u2580".encode()
    want = (b"\x1b[1;1H\x1b[38;5;%d;48;5;%dm      " % (k, k) + b"\x1b[2;1H  \x1b[38;5;%d;48;5;%dm" % (f, b) + blk * 3 + b"\x1b[38;5;%d;48;5;%dm " % (k, k)
            + b"\x1b[3;1H  \x1b[38;5;%d;48;5;%dm" % (f, b) + blk + b"\x1b[38;5;%d;48;5;%dm " % (k, k) + b"\x1b[38;5;%d;48;5;%dm" % (f, b) + blk
            + b"\x1b[38;5;%d;48;5;%dm " % (k, k) + b"\x1b[4;1H      \x1b[0m")
    assert got == want


@pytest.mark.parametrize("scene,pose", [("cornell", None), ("texture_gallery", None), ("volume_grid_test", None), ("teapot", api.BENCH_POSE)])
def test_whole_frames_of_the_reference_source_equal_the_oracle(ref, scene, pose):
    """RefRenderer -- the harness the GPU tests, smoke() and `bench.py --impl reference` drive: the verbatim head and tail of
    TryFlipAndBlit with the reference's own TemporalAA decision -- against the oracle, three frames, every cell field."""
    import ref_binding
    s = api.HostScene(scene)
    rr = ref_binding.RefRenderer(s, 40, 12, 2)
    o = Oracle(s, 40, 12, 2)
    if pose is not None:
        rr.set_camera(*pose)
        o.set_camera(*pose)
    for f in range(3):
        a, c = rr.render_frame(), o.render_frame(threads=2)
        for k in ("glyph", "fg16", "bg16", "fg_ansi", "bg_ansi"):
            assert np.array_equal(a[k], c[k]), f"{scene} frame {f + 1}: {k}"
        assert np.array_equal(bits(a["fg"]), bits(c["fg"])) and np.array_equal(bits(a["bg"]), bits(c["bg"])), f"{scene} frame {f + 1}: SDR colours"
    rr.close()
    o.close()
    s.close()


ISLAND_SIZES = [(64, 128), (96, 128), (256, 256)] + ([(1024, 256)] if os.environ.get("YCGE_FULL_ISLAND") else [])


@pytest.mark.parametrize("size,height", ISLAND_SIZES, ids=[f"{a}x{b}" for a, b in ISLAND_SIZES])
def test_island_generator_equals_the_reference_source(ref, size, height, tmp_path):
    """BASELINE config 4's world.  The reference's island generator as its authors wrote it -- GenMath (gradient noise, FBM, ridged FBM, the
    FNV hash), TerrainNoise (domain warp, island mask, heights, inland water), RiverNetworkGlobal (D8 descent, Array.Sort by height,
    accumulation, carving), BiomeMap, Layering, StrataMap, FloraPlacer (trees, desert props) and the three passes of
    WorldManager.GenerateAndSaveWorld, with BuildMinecraftLike's WorldConfig -- against the host mirror's generator (the one whose
    chunk grids are uploaded to the GPU): every (block id, meta) of the world, through the mirror's own VG01 writer.
    256 x 256 x 256 holds every block kind the full world has (sand, wood, leaves, the three rock metas).  The full 1024 x 256 x 1024
    world (268 M cells, a minute of CPU and 9 GB) was compared once, equal cell for cell; YCGE_FULL_ISLAND=1 runs it again."""
    ref.ref_generate_island.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p]
    n = size * height * size
    ids, metas = np.empty(n, np.int32), np.empty(n, np.int32)
    ref_path = str(tmp_path / "island_ref.vg") if size <= 256 else None   # ... and the reference's own VG01 writer (:607-631): the same file, byte for byte
    assert ref.ref_generate_island(size, height, P(ids), P(metas), ref_path.encode() if ref_path else None) == 0
    path = str(tmp_path / "island.vg")
    assert api.load_host().ycgeh_write_island_world(path.encode(), size, height) == 0
    assert open(path, "rb").read(4) == b"VG01" and tuple(np.fromfile(path, np.int32, 3, offset=4)) == (size, height, size)   # WorldManager.cs:612-616
    cells = np.fromfile(path, np.int32, offset=16).reshape(-1, 2)                                                           # x, y, z order, (id, meta) pairs (:617-629)
    assert len(cells) == n
    assert np.array_equal(cells[:, 0], ids), f"{int((cells[:, 0] != ids).sum())} block ids differ"
    assert np.array_equal(cells[:, 1], metas), f"{int((cells[:, 1] != metas).sum())} metas differ"
    if size >= 256:
        assert set(np.unique(ids)) >= {0, 1, 2, 3, 4, 5, 6, 7} and set(np.unique(metas)) == {0, 1, 2}
    if ref_path:
        assert open(ref_path, "rb").read() == open(path, "rb").read(), "VG01 file of the mirror's writer differs from the reference writer's"


def test_texture_sampler_equals_the_reference_source(ref, oracle_lib):
    """Texture.SampleBilinear (static image: wrap by frac, (w - 1) scaling, the % wrap of the +1 texel, RGBA32.toVec3, two Lerps, Saturate)
    and RGBA32's int constructor as the reference wrote them, against the oracle's sampler (which the device's equals on the GPU)."""
    ref.ref_texture_sample.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    oracle_lib.yo_texture_sample.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    rng = np.random.default_rng(21)
    for w, h in [(12, 6), (1, 1), (2, 5), (64, 64)]:
        px = rng.integers(0, 1 << 32, size=(h, w), dtype=np.uint64).astype(np.uint32)
        uv = [(0.0, 0.0), (1.0, 1.0), (0.5, 0.5), (-0.25, 2.75), (0.99999994, 0.99999994), (1e-8, -1e-8)] + [tuple(x) for x in rng.uniform(-3.0, 3.0, size=(300, 2))]
        a, b = np.empty(3, np.float32), np.empty(3, np.float32)
        for u, v in uv:
            ref.ref_texture_sample(w, h, P(px), u, v, P(a))
            oracle_lib.yo_texture_sample(w, h, P(px), u, v, P(b))
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (w, h, u, v)


def test_history_reset_follows_the_reference_camera_thresholds(ref):
    """TryFlipAndBlit :171 / :266 with the reference's own TemporalAA.cs (ShouldResetHistory :58-67, CommitCamera): whole frames of the
    transpiled reference along a camera path that stays, creeps below the 0.0025 thresholds, jumps above them, turns by just less and
    just more than 0.0025 rad in yaw and in pitch.  The oracle (its own restatement of the rule) must take the same decision every
    frame: the TAA history -- a 1 % blend against a plain copy -- and everything after it are compared bit for bit."""
    import ref_binding
    s = api.HostScene("cornell")
    fb_w, fb_h, ss = 16, 6, 2
    rr = ref_binding.RefRenderer(s, fb_w, fb_h, ss)
    o = Oracle(s, fb_w, fb_h, ss)
    (x, y, z), yaw, pitch = s.default_camera()[:3]
    f32 = lambda v: float(np.float32(v))
    path = [((x, y, z), yaw, pitch, False), ((x, y, z), yaw, pitch, False),
            ((f32(x + 0.001), y, z), yaw, pitch, False),                          # 0.001 < 0.0025: history kept
            ((f32(x + 0.001), f32(y + 0.002), z), yaw, pitch, False),             # 0.002 from the LAST pose: kept (no drift accumulates)
            ((f32(x + 0.001), f32(y + 0.002), f32(z + 0.01)), yaw, pitch, True),  # 0.01: reset
            ((f32(x + 0.001), f32(y + 0.002), f32(z + 0.01)), f32(yaw + 0.002), pitch, False),
            ((f32(x + 0.001), f32(y + 0.002), f32(z + 0.01)), f32(yaw + 0.002 + 0.003), pitch, True),
            ((f32(x + 0.001), f32(y + 0.002), f32(z + 0.01)), f32(yaw + 0.005), f32(pitch - 0.002), False),
            ((f32(x + 0.001), f32(y + 0.002), f32(z + 0.01)), f32(yaw + 0.005), f32(pitch - 0.002 - 0.0026), True)]
    for k, (pos, yw, pt, expect_reset) in enumerate(path):
        rr.set_camera(pos, yw, pt)
        o.set_camera(pos, yw, pt)
        a = rr.render_frame()
        c = o.render_frame(threads=1, fast_post=False)
        assert rr.last_reset == expect_reset, f"frame {k + 1}: the reference's own decision"
        assert np.array_equal(bits(a["taa"]), bits(o.debug_read(api.DBG_TAA)[..., :3])), f"frame {k + 1}: TAA history"
        assert np.array_equal(bits(a["den"]), bits(o.debug_read(api.DBG_DENOISED)[..., :3])), f"frame {k + 1}: denoised"
        assert np.array_equal(a["fg_ansi"], c["fg_ansi"]) and np.array_equal(a["bg_ansi"], c["bg_ansi"]), f"frame {k + 1}: cells"
    rr.close()
    o.close()
    s.close()


def test_resize_follows_the_reference_source(ref):
    """RaytraceRenderer.Resize (:110-138) as the reference wrote it, in the middle of a run: two frames at 16x6 cells ss = 2, Resize to
    20x5 cells ss = 3, two more frames, Resize back.  What the text implies -- history invalid (the first frame after it is a plain
    copy although the camera did not move), camera memory forgotten (taa.Resize), auto-exposure state and frame counter (hence the
    RNG seeds and jitter) carried over -- must come out of the oracle's resize the same way: rays, TAA history, denoised image,
    exposure and cells bit for bit on every frame."""
    import ref_binding
    s = api.HostScene("cornell")
    rr = ref_binding.RefRenderer(s, 16, 6, 2)
    o = Oracle(s, 16, 6, 2)
    frame = 0
    for fb_w, fb_h, ss in [(16, 6, 2), (20, 5, 3), (16, 6, 2)]:
        if frame:
            rr.resize(fb_w, fb_h, ss)
            o.resize(fb_w, fb_h, ss)
        for _ in range(2):
            frame += 1
            a = rr.render_frame()
            c = o.render_frame(threads=1, fast_post=False)
            what = f"frame {frame} at {fb_w}x{fb_h} ss={ss}"
            assert not rr.last_reset, what + ": the camera never moves"
            assert np.array_equal(bits(a["rays"]), bits(o.debug_read(api.DBG_RAYS))), what + ": rays (frame counter carried over)"
            assert np.array_equal(bits(a["taa"]), bits(o.debug_read(api.DBG_TAA)[..., :3])), what + ": TAA history"
            assert np.array_equal(bits(a["den"]), bits(o.debug_read(api.DBG_DENOISED)[..., :3])), what + ": denoised"
            assert bits(a["expo"][:1])[0] == bits(np.float32(o.stats()["ae_exposure"]))[()], what + ": aeExposure (carried over)"
            assert np.array_equal(a["fg_ansi"], c["fg_ansi"]) and np.array_equal(a["bg_ansi"], c["bg_ansi"]) and np.array_equal(a["fg16"], c["fg16"]), what + ": cells"
    rr.close()
    o.close()
    s.close()


PRIM_SCENES = ["cornell", "mirror_spheres", "boxes", "cylinders_disks_triangles", "test", "texture_gallery"]


@pytest.mark.parametrize("scene", PRIM_SCENES)
def test_primitive_hits_equal_the_reference_source(ref, scene):
    """Sphere, Plane, Disk, the three rects, Box, CylinderY and Triangle (scalar path): the objects of the scene factories are rebuilt
    by the reference's OWN constructors (transpiled Surfaces.cs, BoundedObjects.cs, Triangle.cs) from the flat description the product
    consumes, and thousands of rays go through the reference's own Hit methods, objects in order with a shrinking tMax; the oracle's
    linear walk over the same objects must return the same object, the same t, the same normal -- bit for bit."""
    ref.ref_objects_hit.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p, C.c_float, C.c_float] + [C.c_void_p] * 5
    s = api.HostScene(scene)
    f = s.flat.contents
    keep = [k for k in range(f.n_objects) if f.objects[k].kind <= 8]
    if len(keep) != f.n_objects:
        pytest.skip("scene holds meshes / volumes")
    n_obj = f.n_objects
    kind = np.array([f.objects[k].kind for k in range(n_obj)], np.int32)
    p12 = np.array([list(f.objects[k].p) for k in range(n_obj)], np.float32)

    import ref_binding
    textures = ref_binding.set_textures(ref, s)  # noqa: F841  (textured materials point at the harness's copies)

    def mat(i):
        return ref_binding.mat13(f.materials[i])
    ma = np.array([mat(f.objects[k].mat_a) for k in range(n_obj)], np.float32)
    mb = np.array([mat(f.objects[k].mat_b) for k in range(n_obj)], np.float32)
    cs = np.array([f.objects[k].checker_scale for k in range(n_obj)], np.float32)
    sp = np.array([f.objects[k].specular for k in range(n_obj)], np.float32)
    rf = np.array([f.objects[k].reflectivity for k in range(n_obj)], np.float32)
    # normals of planes / disks were normalised by the ctor already; the transpiled ctor normalises again: only idempotent ones qualify
    for k in range(n_obj):
        if kind[k] in (1, 2):
            n = p12[k, 3:6]
            l2 = np.float32(np.float32(n[0] * n[0] + n[1] * n[1]) + n[2] * n[2])
            inv = np.float32(1.0) / np.sqrt(l2, dtype=np.float32)
            assert np.array_equal((n * inv).astype(np.float32).view(np.uint32), n.view(np.uint32)), "a plane / disk normal is not a fixed point of Normalized(): extend the harness"
    o = Oracle(s, 16, 8, 1)
    rng = np.random.default_rng(3)
    cam = np.array(s.default_camera()[0], np.float32)
    n = 6000
    org = np.where(rng.random((n, 1)) < 0.5, cam[None, :], rng.uniform(-3, 3, (n, 3))).astype(np.float32)
    tgt = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    tgt[:, 1] = np.abs(tgt[:, 1])
    d = tgt - org
    d /= np.maximum(1e-6, np.linalg.norm(d, axis=1, keepdims=True))
    d[:200] = np.round(d[:200])  # axis-parallel rays: the degenerate branches (zero components, rays inside slabs)
    d[np.all(d == 0, axis=1)] = (0, -1, 0)
    rays = np.ascontiguousarray(np.concatenate([org, d.astype(np.float32)], 1), np.float32)
    t_o, ids_o, n_o = o.scene_hit(rays, use_bvh=False)
    ids = np.empty(n, np.int32)
    t, nn, pp, mm = np.empty(n, np.float32), np.empty((n, 3), np.float32), np.empty((n, 3), np.float32), np.empty((n, 5), np.float32)
    rc = ref.ref_objects_hit(n_obj, P(kind), P(p12), P(ma), P(mb), P(cs), P(sp), P(rf), n, P(rays), np.float32(0.001), np.float32(3.4028234663852886e38), P(ids), P(t), P(nn), P(pp), P(mm))
    assert rc == 0
    assert np.array_equal(ids, ids_o[:, 0]), f"{scene}: {int((ids != ids_o[:, 0]).sum())} rays hit a different object"
    hit = ids >= 0
    assert hit.sum() > 500
    assert np.array_equal(t[hit].view(np.uint32), t_o[hit].view(np.uint32)), f"{scene}: t differs"
    assert np.array_equal(nn[hit].view(np.uint32), n_o[hit].view(np.uint32)), f"{scene}: normals differ"
    o.close()
    s.close()


@pytest.mark.parametrize("scene", ["teapot", "knot:40x10", "cow", "bunny"])
def test_mesh_tree_and_hits_equal_the_reference_source(ref, scene):
    """MeshBVH.cs whole, transpiled: its constructor builds the tree (binned SAH, the in-place partition, the Array.Sort fallback) from
    the triangles MeshLoader produced; node for node and leaf slot for leaf slot it must be the tree the host mirror built (the one the
    GPU traverses), and its Hit -- BoxHitFast, TriHit, the near-child-first stack walk -- must return the oracle's t and normal."""
    vp = C.c_void_p
    ref.ref_mesh_build.restype = vp
    ref.ref_mesh_build.argtypes = [C.c_int, vp, vp]
    ref.ref_mesh_destroy.argtypes = [vp]
    ref.ref_mesh_info.argtypes = [vp] * 4
    ref.ref_mesh_tree.argtypes = [vp] * 4
    ref.ref_mesh_hit.argtypes = [vp, C.c_int, vp, C.c_float, C.c_float, vp, vp, vp]
    s = api.HostScene(scene)
    tris = np.ascontiguousarray(s.mesh_triangles(0), np.float32)
    m = s.mesh(0).contents.material
    mat = np.array(list(m.albedo) + [m.specular, m.reflectivity] + list(m.emission) + [m.transparency, m.ior] + list(m.transmission), np.float32)
    h = ref.ref_mesh_build(len(tris), P(tris), P(mat))
    assert h
    nn, root, nl = C.c_int(), C.c_int(), C.c_int()
    ref.ref_mesh_info(h, C.byref(nn), C.byref(root), C.byref(nl))
    host = s.bvh_arrays(0)
    assert nn.value == len(host["boxes"]) and root.value == host["root"] and nl.value == len(host["leaf"])
    boxes, lrsc, leaf = np.empty((nn.value, 6), np.float32), np.empty((nn.value, 4), np.int32), np.empty(nl.value, np.int32)
    ref.ref_mesh_tree(h, P(boxes), P(lrsc), P(leaf))
    assert np.array_equal(boxes.view(np.uint32), host["boxes"].view(np.uint32)), "node boxes"
    assert np.array_equal(lrsc, host["lrsc"]) and np.array_equal(leaf, host["leaf"]), "node links / leaf order"
    o = Oracle(s, 16, 8, 1)
    rng = np.random.default_rng(9)
    n = 4000
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    org = (rng.uniform(-1.5, 1.5, (n, 3)) + (lo + hi) / 2).astype(np.float32)
    tgt = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.ascontiguousarray(np.concatenate([org, d.astype(np.float32)], 1), np.float32)
    t_o, ids_o, n_o = o.scene_hit(rays)
    hit, t, nrm = np.empty(n, np.uint8), np.empty(n, np.float32), np.empty((n, 3), np.float32)
    ref.ref_mesh_hit(h, n, P(rays), np.float32(0.001), np.float32(3.4028234663852886e38), P(hit), P(t), P(nrm))
    f = s.flat.contents
    mesh_obj = [k for k in range(f.n_objects) if f.objects[k].kind == 9][0]
    on_mesh = ids_o[:, 0] == mesh_obj
    assert on_mesh.sum() > 1000
    assert hit[on_mesh].all()
    assert np.array_equal(t[on_mesh].view(np.uint32), t_o[on_mesh].view(np.uint32)), "t"
    assert np.array_equal(nrm[on_mesh].view(np.uint32), n_o[on_mesh].view(np.uint32)), "normal"
    closer = (hit == 1) & ~on_mesh & (ids_o[:, 0] >= 0)
    assert (t[closer] >= t_o[closer]).all(), "the reference's mesh walk found a hit the oracle's scene walk missed"
    ref.ref_mesh_destroy(h)
    o.close()
    s.close()


TRACE_CASES = [("cornell", 20, 8, 2, None), ("mirror_spheres", 24, 8, 2, None), ("boxes", 20, 8, 1, None), ("cylinders_disks_triangles", 20, 8, 2, None), ("test", 24, 9, 2, None),
               ("teapot", 20, 8, 2, api.BENCH_POSE), ("knot:40x10", 18, 7, 2, api.BENCH_POSE), ("cow", 16, 6, 2, api.BENCH_POSE),
               ("volume_grid_test", 24, 9, 2, None), ("voxel_world:64x64", 20, 8, 2, None), ("voxel_island:64x64", 20, 8, 2, None),
               # textured materials: SampleAlbedo (:724-735) -> Texture.SampleBilinear (Texture.cs:108-163) of the reference's text, with the U, V its Hit methods report
               ("texture_gallery", 32, 9, 2, None), ("texture_gallery", 24, 8, 2, ((1.2, 1.4, -0.6), 0.5, -0.3)), ("texture_test", 20, 8, 2, ((0.6, 0.4, 0.0), 0.25, -0.2))]


@pytest.mark.parametrize("scene,fb_w,fb_h,ss,pose", TRACE_CASES, ids=[f"{c[0]}-{k}" for k, c in enumerate(TRACE_CASES)])
def test_trace_stage_equals_the_reference_source(ref, scene, fb_w, fb_h, ss, pose):
    """The whole trace stage as the reference wrote it: the scene's objects rebuilt by the reference's constructors (primitives, Mesh ->
    MeshBVH), Scene.RebuildBVH -> BVH.cs, then the verbatim head of TryFlipAndBlit -- frame counter, jitter rotations, MakeJitteredRay,
    PerFrameSeed, TraceFull with its work stack, ComputeTransmittanceToLight, OrenNayar, the cosine-weighted bounce -- for three frames.
    Rays, radiance, albedo, raw normal, depth and the sky mask of every pixel must equal the oracle's, bit for bit."""
    vp = C.c_void_p
    ref.ref_trace_create.restype = vp
    ref.ref_trace_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int] + [vp] * 10 + [C.c_int] + [vp] * 4 + [C.c_int, vp, C.c_int, vp, C.c_int]
    ref.ref_trace_destroy.argtypes = [vp]
    ref.ref_trace_frame.argtypes = [vp, vp, C.c_float, C.c_float] + [vp] * 6
    s = api.HostScene(scene)
    f = s.flat.contents
    n_obj = f.n_objects
    assert all(f.objects[k].kind <= 10 for k in range(n_obj))
    vols = (api.Volume * max(1, s.n_volumes))(*[s.volume(i).contents for i in range(s.n_volumes)])
    kind = np.array([f.objects[k].kind for k in range(n_obj)], np.int32)
    p12 = np.array([list(f.objects[k].p) for k in range(n_obj)], np.float32)

    import ref_binding
    ref_binding.set_textures(ref, s)
    mat = ref_binding.mat13
    ma = np.array([mat(f.materials[max(0, f.objects[k].mat_a)]) for k in range(n_obj)], np.float32)
    mb = np.array([mat(f.materials[max(0, f.objects[k].mat_b)]) for k in range(n_obj)], np.float32)
    cs = np.array([f.objects[k].checker_scale for k in range(n_obj)], np.float32)
    sp = np.array([f.objects[k].specular for k in range(n_obj)], np.float32)
    rf = np.array([f.objects[k].reflectivity for k in range(n_obj)], np.float32)
    meshes = [np.ascontiguousarray(s.mesh_triangles(i), np.float32) for i in range(s.n_meshes)]
    mesh_n = np.array([len(m) for m in meshes] + [0], np.int32)
    mesh_ptrs = (C.c_void_p * max(1, len(meshes)))(*[m.ctypes.data for m in meshes])
    mesh_mat = np.array([mat(s.mesh(i).contents.material) for i in range(s.n_meshes)] + ref_binding.MAT_PAD, np.float32)
    lights = np.array([list(f.lights[i].pos) + list(f.lights[i].color) + [f.lights[i].intensity] for i in range(f.n_lights)] + [[0] * 7], np.float32)
    top, bot = np.array(list(f.bg_top), np.float32), np.array(list(f.bg_bottom), np.float32)
    amb = np.array(list(f.ambient_color) + [f.ambient_intensity], np.float32)
    cam = pose if pose is not None else s.default_camera()[:3]
    fov = s.default_camera()[3]
    all_mats = np.array([mat(f.materials[i]) for i in range(f.n_materials)] + ref_binding.MAT_PAD, np.float32)
    h = ref.ref_trace_create(fb_w, fb_h, ss, fov, n_obj, P(kind), P(p12), P(ma), P(mb), P(cs), P(sp), P(rf), P(mesh_n), mesh_ptrs, P(mesh_mat), f.n_lights, P(lights), P(top), P(bot), P(amb),
                             s.n_volumes, C.cast(vols, C.c_void_p), f.n_materials, P(all_mats), f.is_volume_scene)
    assert h
    o = Oracle(s, fb_w, fb_h, ss)
    o.set_camera(*cam)
    o.debug_read(api.DBG_RAYS) if False else None
    W, H = fb_w * ss, fb_h * 2 * ss
    rays, hdr, alb, nrm = np.empty((H, W, 6), np.float32), np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.float32)
    dep, sky = np.empty((H, W), np.float32), np.empty((H, W), np.uint8)
    c3 = np.array(cam[0], np.float32)
    for frame in range(1, 4):
        o.render_frame(threads=2)
        assert ref.ref_trace_frame(h, P(c3), np.float32(cam[1]), np.float32(cam[2]), P(rays), P(hdr), P(alb), P(nrm), P(dep), P(sky)) == 0
        what = f"{scene} frame {frame}"
        asky = o.debug_read(api.DBG_ALBEDO_SKY)
        assert np.array_equal(sky != 0, asky[..., 3] != 0), what + ": sky mask"
        assert np.array_equal(bits(hdr), bits(o.debug_read(api.DBG_HDR)[..., :3])), what + f": radiance differs in {int((bits(hdr) != bits(o.debug_read(api.DBG_HDR)[..., :3])).any(-1).sum())} pixels"
        assert np.array_equal(bits(alb), bits(asky[..., :3])), what + ": albedo"
        assert np.array_equal(bits(nrm), bits(o.raw_normal())), what + ": normal"
        assert np.array_equal(bits(dep), bits(o.debug_read(api.DBG_NORMAL_DEPTH)[..., 3])), what + ": depth"
    ref.ref_trace_destroy(h)
    o.close()
    s.close()
