"""Generates tests/golden/oracle_*.npz: the CPU oracle's output for small cases of every scene kind plus BASELINE
config 1 at its full size (Cornell box, 240x135 cells, frame 1).  The reference itself cannot run here (C#/.NET, no
toolchain) and ships no golden vectors, so these are OUR vectors: they pin the oracle against regressions and across
machines/compilers (the GPU box's CPU must reproduce them bit for bit), and the GPU path is compared against them too.
    python tools/make_golden.py            (re)generate everything
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from yetanotherconsolegameengine_b200 import api  # noqa: E402
from oracle_binding import Oracle  # noqa: E402

# name, scene, fb_w, fb_h, ss, frames, pose (None = the scene's default camera)
CASES = [
    ("c1_cornell_240x135", "cornell", 240, 135, 1, 1, None),
    ("cornell_small", "cornell", 40, 12, 2, 3, None),
    ("mirror_spheres", "mirror_spheres", 60, 17, 4, 2, None),
    ("cylinders_disks_triangles", "cylinders_disks_triangles", 48, 14, 2, 2, None),
    ("boxes", "boxes", 48, 14, 2, 2, None),
    ("test_scene", "test", 48, 14, 2, 2, None),
    ("volume_grid_test", "volume_grid_test", 48, 14, 2, 2, None),
    ("teapot", "teapot", 48, 14, 2, 2, api.BENCH_POSE),
    ("knot", "knot:60x16", 40, 12, 4, 2, api.BENCH_POSE),
    ("voxel_world", "voxel_world:64x64", 40, 12, 2, 2, None),
    ("texture_gallery", "texture_gallery", 48, 14, 2, 2, None),
    ("texture_test", "texture_test", 40, 12, 2, 2, ((0.6, 0.4, 0.0), 0.25, -0.2)),
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def render_case(scene_name, fb_w, fb_h, ss, frames, pose, threads):
    s = api.HostScene(scene_name)
    o = Oracle(s, fb_w, fb_h, ss)
    if pose is not None:
        o.set_camera(*pose)
    out = {}
    for f in range(frames):
        cells = o.render_frame(threads=threads, fast_post=True)
        st = o.stats()
        out[f"cells_{f + 1}"] = cells
        out[f"prim_{f + 1}"] = o.debug_read(api.DBG_PRIM_ID)
        out[f"sha_hdr_{f + 1}"] = np.array(sha(o.debug_read(api.DBG_HDR)[..., :3]))
        out[f"sha_taa_{f + 1}"] = np.array(sha(o.debug_read(api.DBG_TAA)[..., :3]))
        out[f"sha_den_{f + 1}"] = np.array(sha(o.debug_read(api.DBG_DENOISED)[..., :3]))
        out[f"rays_{f + 1}"] = np.array(st["rays"], np.int64)
        out[f"ae_{f + 1}"] = np.array(st["ae_exposure"], np.float32)
        out[f"logsum_{f + 1}"] = np.array(st["log_sum"], np.float32)
    o.close()
    s.close()
    return out


if __name__ == "__main__":
    dst = os.path.join(ROOT, "tests", "golden")
    only = set(sys.argv[1:])
    for name, scene, fb_w, fb_h, ss, frames, pose in CASES:
        if only and name not in only:
            continue
        out = render_case(scene, fb_w, fb_h, ss, frames, pose, threads=os.cpu_count() or 1)
        path = os.path.join(dst, f"oracle_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes; rays frame 1:", int(out["rays_1"]))
