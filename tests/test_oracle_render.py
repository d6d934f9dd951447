"""CPU oracle, frame level: reproduces the committed golden vectors bit for bit on this machine; independent of thread
count; BVH == brute force; the in-place a-trous iteration (RaytraceRenderer.cs:718) checked against a literal Python
restatement of ApplyAtrousDenoise; TAA/exposure state machine."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from yetanotherconsolegameengine_b200 import api
from oracle_binding import Oracle, load_oracle

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden  # noqa: E402


def cells_equal(a, b):
    for k in ("glyph", "fg16", "bg16", "fg_ansi", "bg_ansi", "attr"):
        if not np.array_equal(a[k], b[k]):
            return False
    return np.array_equal(a["fg"].view(np.uint32), b["fg"].view(np.uint32)) and np.array_equal(a["bg"].view(np.uint32), b["bg"].view(np.uint32))


@pytest.mark.parametrize("case", make_golden.CASES, ids=[c[0] for c in make_golden.CASES])
def test_oracle_reproduces_golden(case):
    name, scene, fb_w, fb_h, ss, frames, pose = case
    gold = np.load(os.path.join(GOLDEN, f"oracle_{name}.npz"))
    out = make_golden.render_case(scene, fb_w, fb_h, ss, frames, pose, threads=2)
    for k in gold.files:
        if k.startswith("cells_"):
            assert cells_equal(out[k], gold[k]), k
        else:
            assert np.array_equal(out[k], gold[k]), k


def test_c1_golden_is_a_plausible_cornell_box():
    """Sanity of BASELINE config 1 (the reference's CPU-runnable case): every primary ray hits the closed box, the cells
    are U+2580, colours stay inside the 6x6x6 cube, left wall red / right wall green as Scenes.cs:269-309 builds them."""
    g = np.load(os.path.join(GOLDEN, "oracle_c1_cornell_240x135.npz"))
    cells, prim = g["cells_1"], g["prim_1"]
    assert cells.shape == (135, 240) and prim.shape == (270, 240, 2)
    assert np.all(cells["glyph"] == 0x2580)
    assert cells["fg_ansi"].min() >= 16 and cells["fg_ansi"].max() <= 231
    assert np.all(cells["attr"] == (cells["fg16"] & 15) | ((cells["bg16"].astype(np.uint16) & 15) << 4))
    assert (prim[..., 0] >= 0).mean() > 0.99
    left, right = cells["fg"][60:80, 5], cells["fg"][60:80, 234]
    assert left[:, 0].mean() > left[:, 1].mean() and right[:, 1].mean() > right[:, 0].mean()


def test_thread_count_and_fast_post_do_not_change_results():
    s = api.HostScene("mirror_spheres")
    outs = []
    for threads, fast in ((1, False), (3, False), (4, True)):
        o = Oracle(s, 24, 8, 2)
        frames = [o.render_frame(threads=threads, fast_post=fast).copy() for _ in range(2)]
        outs.append((frames, o.debug_read(api.DBG_DENOISED).copy(), o.stats()["rays"]))
        o.close()
    for frames, den, rays in outs[1:]:
        assert cells_equal(frames[0], outs[0][0][0]) and cells_equal(frames[1], outs[0][0][1])
        assert np.array_equal(den.view(np.uint32), outs[0][1].view(np.uint32)) and rays == outs[0][2]


@pytest.mark.parametrize("name", ["cornell", "cylinders_disks_triangles", "boxes", "knot:24x8", "volume_grid_test"])
def test_bvh_agrees_with_brute_force(name):
    s = api.HostScene(name)
    o = Oracle(s, 8, 4, 1)
    rng = np.random.default_rng(11)
    n = 3000
    org = rng.uniform(-2.5, 2.5, (n, 3)).astype(np.float32) + np.array([0, 1.0, 0], np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    rays = np.concatenate([org, d], 1)
    t1, id1, n1 = o.scene_hit(rays, use_bvh=True)
    t0, id0, n0 = o.scene_hit(rays, use_bvh=False)
    # nearest-hit distance must agree exactly; ids may differ only at exact ties (last visited wins, BVH.cs:169-181)
    assert np.array_equal(t1.view(np.uint32), t0.view(np.uint32))
    differ = (id1 != id0).any(1)
    assert differ.mean() < 0.01
    assert (t1 >= 0).mean() > 0.2
    o.close()


def atrous_literal(src, albedo, normal, depth, sky, exp_f, iterations=3, phis=(3.0, 0.35, 2.0, 0.20)):
    """ApplyAtrousDenoise (RaytraceRenderer.cs:622-722) transcribed with the reference's own buffer juggling: `cur` and
    `dst` are Python references, so the aliasing after the first swap (:718) happens here exactly as it does in C#."""
    f = np.float32
    h, w = sky.shape
    k = [f(1 / 16), f(1 / 4), f(3 / 8), f(1 / 4), f(1 / 16)]
    scratchA, scratchB = np.zeros_like(src), np.zeros_like(src)
    cur, dst = src, scratchA
    cphi, nphi, zphi, aphi = (max(f(1e-6), f(p)) for p in phis)

    def normalized(v):
        l2 = f(f(f(v[0] * v[0]) + f(v[1] * v[1])) + f(v[2] * v[2]))
        if l2 <= 0:
            return v
        inv = f(f(1.0) / np.sqrt(l2, dtype=f))
        return np.array([v[0] * inv, v[1] * inv, v[2] * inv], f)

    def lum(c):
        return f(f(f(f(0.2126) * c[0]) + f(f(0.7152) * c[1])) + f(f(0.0722) * c[2]))

    for it in range(max(1, iterations)):
        step = 1 << it
        for y in range(h):
            for x in range(w):
                if sky[y, x]:
                    dst[y, x] = cur[y, x]
                    continue
                c0 = cur[y, x].copy()
                a0, n0, z0 = albedo[y, x], normalized(normal[y, x]), depth[y, x]
                wsum = f(0)
                acc = np.zeros(3, f)
                for ky in range(-2, 3):
                    sy = min(max(y + ky * step, 0), h - 1)
                    for kx in range(-2, 3):
                        sx = min(max(x + kx * step, 0), w - 1)
                        if sky[sy, sx] != sky[y, x]:
                            continue
                        wbase = f(k[kx + 2] * k[ky + 2])
                        c, a, n, z = cur[sy, sx], albedo[sy, sx], normalized(normal[sy, sx]), depth[sy, sx]
                        dl = abs(f(lum(c) - lum(c0)))
                        dn = max(f(0), f(f(1) - f(f(f(n0[0] * n[0]) + f(n0[1] * n[1])) + f(n0[2] * n[2]))))
                        dz = abs(f(z - z0))
                        da = f(f(abs(f(a[0] - a0[0])) + abs(f(a[1] - a0[1]))) + abs(f(a[2] - a0[2])))
                        wc, wn, wz, wa = exp_f(f(-dl / cphi)), exp_f(f(-dn / nphi)), exp_f(f(-dz / zphi)), exp_f(f(-da / aphi))
                        wg = f(f(f(f(wbase * wc) * wn) * wz) * wa)
                        acc = np.array([acc[0] + f(c[0] * wg), acc[1] + f(c[1] * wg), acc[2] + f(c[2] * wg)], f)
                        wsum = f(wsum + wg)
                if wsum > f(1e-8):
                    inv = f(f(1) / wsum)
                    dst[y, x] = acc * inv
                else:
                    dst[y, x] = c0
        tmp = cur
        cur = dst
        dst = scratchB if tmp is scratchA else scratchA  # :718 — tmp is `src` after pass 0, so dst stays scratchA: in place
    return cur


def test_atrous_in_place_iteration_matches_literal_restatement():
    lib = load_oracle()
    lib.yo_set_math_mode(0)
    exp_f = lambda x: np.float32(lib.yo_math(0, float(x), 0.0))
    s = api.HostScene("boxes")
    o = Oracle(s, 14, 5, 1)  # 14x10 pixels: sky + geometry, so the sky-mismatch skip is exercised
    o.render_frame(threads=1)
    taa = o.debug_read(api.DBG_TAA)[..., :3].copy()
    alb_sky = o.debug_read(api.DBG_ALBEDO_SKY)
    nd = o.debug_read(api.DBG_NORMAL_DEPTH)
    den = o.debug_read(api.DBG_DENOISED)[..., :3]
    sky = alb_sky[..., 3] != 0
    assert 0 < sky.sum() < sky.size
    with np.errstate(over="ignore", invalid="ignore"):
        ref = atrous_literal(taa, alb_sky[..., :3].copy(), nd[..., :3].copy(), nd[..., 3].copy(), sky, exp_f)
    assert np.array_equal(ref.view(np.uint32), den.view(np.uint32))
    o.close()


def test_history_reset_and_exposure_state_machine():
    s = api.HostScene("cornell")
    o = Oracle(s, 16, 6, 1)
    o.render_frame()
    hdr1, taa1 = o.debug_read(api.DBG_HDR), o.debug_read(api.DBG_TAA)
    assert np.array_equal(hdr1.view(np.uint32), taa1.view(np.uint32))  # first frame: history <- current (:285-303)
    ae1 = o.stats()["ae_exposure"]
    o.render_frame()
    hdr2, taa2 = o.debug_read(api.DBG_HDR), o.debug_read(api.DBG_TAA)
    assert not np.array_equal(hdr2, taa2)  # alpha = 0.01 blend
    # camera moved by more than 0.0025 -> reset (TemporalAA.cs:58-67)
    pos, yaw, pitch, _ = s.default_camera()
    o.set_camera((pos[0] + 0.01, pos[1], pos[2]), yaw, pitch)
    o.render_frame()
    assert np.array_equal(o.debug_read(api.DBG_HDR).view(np.uint32), o.debug_read(api.DBG_TAA).view(np.uint32))
    # resize keeps frame counter and exposure (RaytraceRenderer.cs:110-138)
    ae3 = o.stats()["ae_exposure"]
    o.resize(12, 5, 2)
    assert o.stats()["frames"] == 3 and o.stats()["ae_exposure"] == ae3
    o.render_frame()
    assert o.stats()["frames"] == 4
    assert 0.10 <= o.stats()["ae_exposure"] <= 1.50 and ae1 != 1.0
    o.close()
