"""Known-answer tests that pin the CPU oracle's bit-defined pieces.

The reference has no tests or golden vectors ("parity unpinned"); these vectors are derived from the arithmetic written
in the reference source with arbitrary-precision Python integers, independent of the oracle's C++.
Citations are relative to /root/reference/ConsoleGame/.
"""
import ctypes as C
import struct

import numpy as np
import pytest

M64 = (1 << 64) - 1


def sm64(z):  # RaytraceSampler.cs:71-80
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def per_frame_seed(x, y, frame, jx=0, jy=0, salt=0x9E3779B97F4A7C15):  # RaytraceSampler.cs:56-68
    h = 1469598103934665603
    for v, k in ((x, 0x9E3779B97F4A7C15), (y, 0xC2B2AE3D27D4EB4F), (frame, 0x165667B19E3779F9)):
        h ^= ((v & M64) * k) & M64
        h = sm64(h)
    h ^= (((jx & M64) << 32) ^ (jy & 0xFFFFFFFF)) & M64 if (jx or jy) else 0
    h = sm64(h)
    h ^= salt
    return sm64(h)


def f32(x):
    return struct.unpack("<f", struct.pack("<f", x))[0]


def bits(f):
    return struct.unpack("<I", struct.pack("<f", f))[0]


def test_splitmix64_canonical(oracle_lib):
    # canonical SplitMix64: first output from state 0
    assert sm64(0) == 0xE220A8397B1DCDAF
    assert oracle_lib.yo_splitmix64(0) == 0xE220A8397B1DCDAF
    rng = np.random.default_rng(1)
    for z in rng.integers(0, 1 << 63, 200, dtype=np.uint64):
        assert oracle_lib.yo_splitmix64(int(z)) == sm64(int(z))


SEED_KATS = [  # (x, y, frame) -> seed, first m24
    ((0, 0, 1), 0x17EF7D0094EB2C76, 7379119),
    ((1, 0, 1), 0xE30B62FE1AC2EDC5, 1918614),
    ((0, 1, 1), 0x7EDFB4004F82140E, 13519905),
    ((959, 539, 1), 0xC3C5BBAF2004442A, 15095364),
    ((1919, 1079, 64), 0x5BEAD3AD13E75BBB, 9491773),
]


@pytest.mark.parametrize("xyf,seed,m24", SEED_KATS)
def test_per_frame_seed_kat(oracle_lib, xyf, seed, m24):
    x, y, f = xyf
    assert per_frame_seed(x, y, f) == seed
    assert oracle_lib.yo_per_frame_seed(x, y, f, 0, 0, 0x9E3779B97F4A7C15) == seed
    b = np.zeros(1, np.uint32)
    m = np.zeros(1, np.uint32)
    oracle_lib.yo_rng_draws(seed, 1, b.ctypes.data, m.ctypes.data)
    assert int(m[0]) == m24 == sm64(seed) >> 40
    # NextUnit = ((state>>40) + 0.5f) * (1/16777216f): the float add rounds to even for m24 >= 2^23 (RaytraceSampler.cs:47-52)
    expect = np.float32(np.float32(m24) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    assert int(b[0]) == bits(float(expect))


def test_per_frame_seed_random(oracle_lib):
    rng = np.random.default_rng(2)
    for _ in range(300):
        x, y = int(rng.integers(0, 4096)), int(rng.integers(0, 4096))
        f = int(rng.integers(1, 1 << 40))
        assert oracle_lib.yo_per_frame_seed(x, y, f, 0, 0, 0x9E3779B97F4A7C15) == per_frame_seed(x, y, f)


def test_rng_stream_iterates_the_output_function(oracle_lib):
    # Rng.NextUnit: state = SplitMix64(state) (not the canonical counter mode)  RaytraceSampler.cs:47-52
    seed = 0x17EF7D0094EB2C76
    n = 64
    b = np.zeros(n, np.uint32)
    m = np.zeros(n, np.uint32)
    oracle_lib.yo_rng_draws(seed, n, b.ctypes.data, m.ctypes.data)
    s = seed
    for i in range(n):
        s = sm64(s)
        assert int(m[i]) == s >> 40
        v = np.float32(np.float32(s >> 40) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
        assert int(b[i]) == bits(float(v))
        assert 0.0 < float(v) < 1.0 or float(v) == 1.0  # (2^24-1+0.5f) rounds up to 2^24 -> exactly 1.0f is reachable
    # Rng(0) substitutes the golden constant (:41-44)
    b0 = np.zeros(1, np.uint32)
    oracle_lib.yo_rng_draws(0, 1, b0.ctypes.data, None)
    m24 = sm64(0x9E3779B97F4A7C15) >> 40
    assert int(b0[0]) == bits(float(np.float32(np.float32(m24) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)))


def test_rng_cs_stream(oracle_lib):
    # ConsoleRayTracing.Rng (Rng.cs:3-29): state = Scramble(seed + g); next: state += g; z = Scramble(state); (float)((z>>11) * 2^-53)
    g = 0x9E3779B97F4A7C15

    def scramble(x):
        x ^= x >> 30
        x = (x * 0xBF58476D1CE4E5B9) & M64
        x ^= x >> 27
        x = (x * 0x94D049BB133111EB) & M64
        x ^= x >> 31
        return x

    for seed in (0, 1, 12345, 0xDEADBEEFCAFEBABE):
        n = 32
        b = np.zeros(n, np.uint32)
        oracle_lib.yo_rng_cs_draws(seed, n, b.ctypes.data)
        state = scramble((seed + g) & M64)
        for i in range(n):
            state = (state + g) & M64
            z = scramble(state)
            v = np.float32((z >> 11) * (1.0 / 9007199254740992.0))
            assert int(b[i]) == bits(float(v)), (seed, i)


BLUE_NOISE = [  # RaytraceSampler.cs:9-19 (an 8x8 ordered-dither permutation of 0..63)
    [0, 32, 8, 40, 2, 34, 10, 42], [48, 16, 56, 24, 50, 18, 58, 26], [12, 44, 4, 36, 14, 46, 6, 38], [60, 28, 52, 20, 62, 30, 54, 22],
    [3, 35, 11, 43, 1, 33, 9, 41], [51, 19, 59, 27, 49, 17, 57, 25], [15, 47, 7, 39, 13, 45, 5, 37], [63, 31, 55, 23, 61, 29, 53, 21]]


def test_blue_noise_table_and_sample(oracle_lib):
    flat = [oracle_lib.yo_blue_noise_table(iy, ix) for iy in range(8) for ix in range(8)]
    assert flat == [v for row in BLUE_NOISE for v in row]
    assert sorted(flat) == list(range(64))
    oracle_lib.yo_blue_noise.restype = C.c_float
    for (x, y, fi, ch) in ((0, 0, 1, 0), (5, 3, 7, 1), (1919, 1079, 64, 0), (13, 250, 100000, 1)):
        base = np.float32(np.float32(BLUE_NOISE[y & 7][x & 7]) + np.float32(0.5)) * np.float32(1.0 / 64.0)
        k = np.float32(0.7548776662466927) if ch == 0 else np.float32(0.5698402909980532)
        rot = np.float32(np.float32(fi + 1) * k)
        rot = np.float32(rot - np.floor(rot))
        v = np.float32(base + rot)
        v = np.float32(v - np.floor(v))
        assert bits(oracle_lib.yo_blue_noise(x, y, fi, ch)) == bits(float(v))


def test_morton_and_index_of(oracle_lib):
    # VolumeGrid.cs:235-252: local bits [x2 y2 z2 x1 y1 z1 x0 y0 z0] with bit0 = x0, bit1 = y0, bit2 = z0
    seen = set()
    for z in range(8):
        for y in range(8):
            for x in range(8):
                m = 0
                for b in range(3):
                    m |= ((x >> b) & 1) << (3 * b) | ((y >> b) & 1) << (3 * b + 1) | ((z >> b) & 1) << (3 * b + 2)
                assert oracle_lib.yo_morton3(x, y, z) == m
                seen.add(m)
    assert seen == set(range(512))
    nx, ny, nz = 20, 9, 33
    nbx, nby = (nx + 7) >> 3, (ny + 7) >> 3
    idx = set()
    for iz in range(nz):
        for iy in range(ny):
            for ix in range(nx):
                brick = ((iz >> 3) * nby + (iy >> 3)) * nbx + (ix >> 3)
                exp = brick * 512 + oracle_lib.yo_morton3(ix & 7, iy & 7, iz & 7)
                got = oracle_lib.yo_volume_index_of(nx, ny, nz, ix, iy, iz)
                assert got == exp
                idx.add(got)
    assert len(idx) == nx * ny * nz  # injective


def srgb8(c):  # ANSITerminalRenderer.cs:298-307, Math.Round = banker's rounding = Python round()
    c = min(1.0, max(0.0, c))
    s = 12.92 * c if c <= 0.0031308 else 1.055 * (c ** (1.0 / 2.4)) - 0.055
    v = int(round(s * 255.0))
    return min(255, max(0, v))


def cube(v):  # :288-296
    return 0 if v < 48 else 1 if v < 114 else 2 if v < 154 else 3 if v < 194 else 4 if v < 234 else 5


def test_ansi256_quantisation(oracle_lib):
    for v in range(256):
        assert oracle_lib.yo_cube_level(v) == cube(v)
    rng = np.random.default_rng(3)
    for c in list(rng.random(2000)) + [0.0, 1.0, 0.0031308, 0.5, 0.18]:
        assert oracle_lib.yo_linear_to_srgb8(float(c)) == srgb8(float(c))
    for r, g, b in rng.random((3000, 3)).astype(np.float32):
        idx = oracle_lib.yo_ansi256(float(r), float(g), float(b))
        exp = 16 + 36 * cube(srgb8(float(r))) + 6 * cube(srgb8(float(g))) + cube(srgb8(float(b)))
        assert idx == exp  # the gray-ramp branch is dead: s_graySrgb is never filled (:26)
        assert 16 <= idx <= 231


def test_gray_ramp_never_wins_exhaustive():
    """ChexelToAnsi256 compares the cube candidate with gray candidate index 232 + k whose table value is always 0
    (s_graySrgb is allocated but never filled): dGray = r8^2+g8^2+b8^2 must never be < dCube.  Exhaustive over 2^24."""
    lv = np.array([0, 95, 135, 175, 215, 255], np.int64)  # xterm cube levels :29
    v = np.arange(256, dtype=np.int64)
    lvl = np.array([cube(int(x)) for x in v])
    d1 = (v - lv[lvl]) ** 2  # per-channel cube distance
    g1 = v ** 2              # per-channel distance to gray value 0
    # dGray < dCube  <=>  sum(g1) < sum(d1); since g1 >= d1 channel-wise this can never hold
    assert np.all(g1 >= d1)


PALETTE16 = [(0, 0, 0), (0, 0, .5), (0, .5, 0), (0, .5, .5), (.5, 0, 0), (.5, 0, .5), (.5, .5, 0), (.75, .75, .75),
             (.5, .5, .5), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]  # Chexel.cs:11-29


def test_nearest_console_colour(oracle_lib):
    for i, (r, g, b) in enumerate(PALETTE16):
        assert oracle_lib.yo_nearest16(r, g, b) == i
    rng = np.random.default_rng(4)
    pal = np.array(PALETTE16, np.float32)
    for c in rng.random((2000, 3)).astype(np.float32):
        d = ((c[None, :] - pal) ** 2)
        d = (d[:, 0] + d[:, 1]) + d[:, 2]  # float32, (x+y)+z
        assert oracle_lib.yo_nearest16(float(c[0]), float(c[1]), float(c[2])) == int(np.argmin(d))  # first minimum wins (strict <)
    # out-of-range inputs are clamped first (Chexel.cs:37-41)
    assert oracle_lib.yo_nearest16(2.0, -1.0, -1.0) == 12


def test_detmath_accuracy(oracle_lib):
    """ycge_detmath.h stands in for MathF.* (platform libm in .NET, not bit-reproducible): within 1 ulp of numpy's
    float64 result rounded to binary32 (it is the correctly rounded value except ~1e-6 of inputs)."""
    oracle_lib.yo_set_math_mode(0)
    rng = np.random.default_rng(5)

    def ulps(a, b):
        ia = np.array([a], np.float32).view(np.int32)[0]
        ib = np.array([b], np.float32).view(np.int32)[0]
        return abs(int(ia) - int(ib))

    for x in np.concatenate([-rng.random(300) * 30.0, rng.random(50) * 5.0]).astype(np.float32):
        assert ulps(oracle_lib.yo_math(0, float(x), 0.0), np.float32(np.exp(np.float64(x)))) <= 1
    for x in (rng.random(300) * 100.0 + 1e-6).astype(np.float32):
        assert ulps(oracle_lib.yo_math(1, float(x), 0.0), np.float32(np.log(np.float64(x)))) <= 1
    for x in rng.random(300).astype(np.float32):
        assert ulps(oracle_lib.yo_math(2, float(x), float(np.float32(1.0 / 2.2))), np.float32(np.float64(x) ** np.float64(np.float32(1.0 / 2.2)))) <= 1
        assert ulps(oracle_lib.yo_math(2, float(x), 5.0), np.float32(np.float64(x) ** 5.0)) <= 1
    for x in (rng.random(300) * 6.2831855).astype(np.float32):
        assert abs(oracle_lib.yo_math(3, float(x), 0.0) - np.sin(np.float64(x))) < 1e-7
        assert abs(oracle_lib.yo_math(4, float(x), 0.0) - np.cos(np.float64(x))) < 1e-7
    assert oracle_lib.yo_math(0, 0.0, 0.0) == 1.0
    assert oracle_lib.yo_math(2, 0.0, 0.45454547) == 0.0


def test_dotnet_sort_restatement(oracle_lib):
    """Array.Sort (introsort) restated for the BVH builders' fallback: result must be sorted and a permutation."""
    rng = np.random.default_rng(6)
    for n in (0, 1, 2, 3, 15, 16, 17, 33, 100, 1000):
        keys = rng.integers(0, max(1, n // 3 + 1), n).astype(np.float32)  # many duplicates: the unstable case
        pay = np.arange(n, dtype=np.int32)
        k2, p2 = keys.copy(), pay.copy()
        oracle_lib.yo_dotnet_sort_floats(k2.ctypes.data, p2.ctypes.data, n)
        assert np.all(np.diff(k2) >= 0)
        assert sorted(p2.tolist()) == list(range(n))
        assert np.array_equal(keys[p2], k2)
    # n <= 16 is a plain insertion sort in .NET (stable): payload order of equal keys is preserved
    keys = np.array([1, 0, 1, 0, 1, 0, 1, 0], np.float32)
    pay = np.arange(8, dtype=np.int32)
    oracle_lib.yo_dotnet_sort_floats(keys.ctypes.data, pay.ctypes.data, 8)
    assert pay.tolist() == [1, 3, 5, 7, 0, 2, 4, 6]


def test_texture_sample_bilinear_known_answers():
    """Texture.SampleBilinear (Renderer/Texture.cs:143-162): fraction wrap, (size - 1) scaling, modulo neighbour, byte/255f
    texels, two lerps, saturate -- transcribed independently with numpy binary32 scalars."""
    from oracle_binding import load_oracle
    o = load_oracle()
    F = np.float32
    rng = np.random.default_rng(7)
    w, h = 13, 7
    px = rng.integers(0, 1 << 32, size=(h, w), dtype=np.uint64).astype(np.uint32)

    def texel(x, y):
        v = int(px[y, x])
        return [F(v & 255) / F(255.0), F((v >> 8) & 255) / F(255.0), F((v >> 16) & 255) / F(255.0)]

    def sample(u, v):
        u, v = F(u), F(v)
        u = F(u - np.floor(u)); v = F(v - np.floor(v))
        fx, fy = F(u * F(w - 1)), F(v * F(h - 1))
        x0, y0 = int(np.floor(fx)), int(np.floor(fy))
        x1, y1 = (x0 + 1) % w, (y0 + 1) % h
        tx, ty = F(fx - F(x0)), F(fy - F(y0))
        lerp = lambda a, b, t: [F(F(a[k] * F(F(1.0) - t)) + F(b[k] * t)) for k in range(3)]
        c = lerp(lerp(texel(x0, y0), texel(x1, y0), tx), lerp(texel(x0, y1), texel(x1, y1), tx), ty)
        return [min(max(x, F(0.0)), F(1.0)) for x in c]

    uv = [(0.0, 0.0), (1.0, 1.0), (0.5, 0.5), (-0.25, 2.75), (0.99999994, 0.99999994), (1e-8, -1e-8), (12.0 / 12.0, 3.0 / 6.0), (1.0 / 12.0, 5.0 / 6.0)]
    uv += [tuple(x) for x in rng.uniform(-3.0, 3.0, size=(200, 2))]
    out = (C.c_float * 3)()
    for u, v in uv:
        o.yo_texture_sample(w, h, px.ctypes.data, C.c_float(u), C.c_float(v), out)
        want = sample(u, v)
        assert [bits(float(x)) for x in out] == [bits(float(x)) for x in want], (u, v)


def test_tables_match_the_reference_source(oracle_lib):
    """BlueNoise8x8, the 16-colour console palette and the colour-cube thresholds as the oracle holds them, against the tables
    extracted from RaytraceSampler.cs:9-19, Renderer/Chexel.cs:11-29 and ANSITerminalRenderer.cs:288-296 by
    tools/extract_scene_literals.py (tests/golden/scene_literals.json)."""
    import json
    import os
    t = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scene_literals.json")))["tables"]
    assert [[oracle_lib.yo_blue_noise_table(iy, ix) for ix in range(8)] for iy in range(8)] == t["blue_noise"]
    for i, (r, g, b) in enumerate(t["palette16"]):  # every palette entry is its own nearest colour; 7 (0.75 grey) and 8 (0.5 grey) are distinct
        assert oracle_lib.yo_nearest16(r, g, b) == i
    level = lambda v: next((lv for bound, lv in t["cube_thresholds"] if v < bound), 5)
    assert [oracle_lib.yo_cube_level(v) for v in range(256)] == [level(v) for v in range(256)]
