"""Generates tests/golden/fullsize_hashes.json: the CPU oracle's output at BASELINE.json's FULL sizes, as SHA-256 digests.

The small cases of tools/make_golden.py store whole cell arrays; at 1920x1080 .. 3840x2160 the arrays would be tens of MB and
the oracle needs seconds to minutes per frame, so these cases are rendered ONCE here (CPU, no GPU involved) and the GPU
suite compares digests of its own planes with them (tests/test_gpu_parity.py::test_full_size_*_vs_oracle_hashes):
bit-exactness is all-or-nothing anyway.

  C3  showcase x2 + bunny + teapot, 1920x1080 (480x135 cells, ss 4), frames 1, 2 and 64 (TAA accumulated, alpha 0.01)
  C4  voxel world (synthetic 1024x256x1024) and the reference's island world (seed 0), 2560x1440 (320x90, ss 8), frames 1, 2
  C5  dragon (stand-in) 3840x2160 (480x135, ss 8), frames 1, 2
      python tools/make_golden_fullsize.py [case ...]      (re)generate the named cases (default: all), merging into the file
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from yetanotherconsolegameengine_b200 import api  # noqa: E402
from oracle_binding import Oracle  # noqa: E402

# name: (scene, fb_w, fb_h, ss, frames to record, pose)
CASES = {
    "c3_cylinders_disks_triangles": ("cylinders_disks_triangles", 480, 135, 4, (1, 2, 64), None),
    "c3_boxes": ("boxes", 480, 135, 4, (1, 2, 64), None),
    "c3_bunny": ("bunny", 480, 135, 4, (1, 2, 64), api.BENCH_POSE),
    "c3_teapot": ("teapot", 480, 135, 4, (1, 2, 64), api.BENCH_POSE),
    "c4_voxel_world_1440p": ("voxel_world", 320, 90, 8, (1, 2), None),
    "c4_voxel_island_1440p": ("voxel_island", 320, 90, 8, (1, 2), None),
    "c5_dragon_2160p": ("dragon", 480, 135, 8, (1, 2), api.BENCH_POSE),
    "c5_dragon_1080p": ("dragon", 480, 135, 4, (1, 2), api.BENCH_POSE),
}
CELL_KEYS = ("glyph", "fg16", "bg16", "fg_ansi", "bg_ansi", "attr", "fg", "bg")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def frame_digest(cells, prim, hdr, taa, den, stats):
    """The digests both sides compute (the GPU test imports this function)."""
    d = {"cells_" + k: sha(cells[k]) for k in CELL_KEYS}
    d["prim"] = sha(prim)
    d["hdr"], d["taa"], d["den"] = sha(hdr[..., :3]), sha(taa[..., :3]), sha(den[..., :3])
    d["rays"] = int(stats["rays"])
    d["ae_exposure_bits"] = int(np.float32(stats["ae_exposure"]).view(np.uint32))
    d["log_sum_bits"] = int(np.float32(stats["log_sum"]).view(np.uint32))
    return d


def render_case(scene_name, fb_w, fb_h, ss, frames, pose, threads):
    s = api.HostScene(scene_name)
    o = Oracle(s, fb_w, fb_h, ss)
    if pose is not None:
        o.set_camera(*pose)
    out = {"scene": s.name, "fb_w": fb_w, "fb_h": fb_h, "ss": ss, "triangles": s.counts()["triangles"], "frames": {}}
    for f in range(1, max(frames) + 1):
        cells = o.render_frame(threads=threads, fast_post=True)
        if f in frames:
            out["frames"][str(f)] = frame_digest(cells, o.debug_read(api.DBG_PRIM_ID), o.debug_read(api.DBG_HDR), o.debug_read(api.DBG_TAA),
                                                 o.debug_read(api.DBG_DENOISED), o.stats())
    o.close()
    s.close()
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "fullsize_hashes.json")
    doc = json.load(open(path)) if os.path.exists(path) else {}
    for name in (sys.argv[1:] or list(CASES)):
        scene, fb_w, fb_h, ss, frames, pose = CASES[name]
        t0 = time.time()
        doc[name] = render_case(scene, fb_w, fb_h, ss, frames, pose, threads=os.cpu_count() or 1)
        doc[name]["oracle_seconds"] = round(time.time() - t0, 1)
        json.dump(doc, open(path, "w"), indent=1, sort_keys=True)
        print(name, doc[name]["oracle_seconds"], "s", flush=True)
