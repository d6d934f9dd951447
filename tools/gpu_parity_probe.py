"""Ad-hoc GPU-vs-oracle comparison across scenes; prints per-stage mismatch counts. (Development aid; the real gates are tests/.)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import yetanotherconsolegameengine_b200 as pkg
from yetanotherconsolegameengine_b200 import api
from oracle_binding import Oracle

def cmp_bits(a, b):
    return int((a.view(np.uint32) != b.view(np.uint32)).sum())

def run(name, fb_w, fb_h, ss, frames=2, pose=None, threads=None):
    threads = threads or os.cpu_count()
    s = pkg.HostScene(name)
    r = pkg.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    o = Oracle(s, fb_w, fb_h, ss)
    if pose:
        r.SetCamera(*pose); o.set_camera(*pose)
    r.debug_read(api.DBG_RAYS)
    for f in range(frames):
        t0 = time.time(); g = r.render_frame_stats(); t1 = time.time(); c = o.render_frame(threads=threads); t2 = time.time()
        gs, cs = r.stats(), o.stats()
        line = [f"{name} {fb_w}x{fb_h} ss{ss} f{f+1}: gpu {1e3*(t1-t0):.1f}ms (dev {gs['ms_total']:.2f}) cpu {1e3*(t2-t1):.0f}ms"]
        for kind, nm in ((api.DBG_RAYS, "rays"), (api.DBG_PRIM_ID, "prim"), (api.DBG_HDR, "hdr"), (api.DBG_ALBEDO_SKY, "alb"), (api.DBG_NORMAL_DEPTH, "nd"),
                         (api.DBG_TAA, "taa"), (api.DBG_DENOISED, "den"), (api.DBG_LOG_SAMPLES, "logs")):
            a, b = r.debug_read(kind), o.debug_read(kind)
            if kind == api.DBG_PRIM_ID:
                d = int((a != b).any(-1).sum())
            else:
                d = cmp_bits(a, b)
            line.append(f"{nm}:{d}")
        cells = sum(int((g[k] != c[k]).sum()) for k in ("glyph", "fg16", "bg16", "fg_ansi", "bg_ansi", "attr")) + cmp_bits(g["fg"], c["fg"]) + cmp_bits(g["bg"], c["bg"])
        line.append(f"cells:{cells}")
        keys = ("rays", "top_nodes_popped", "mesh_nodes_popped", "leaf_refs", "tris_tested", "prims_tested", "dda_cells")
        line.append("counters:" + ("OK" if all(gs[k] == cs[k] for k in keys) else str([(k, gs[k], cs[k]) for k in keys if gs[k] != cs[k]])))
        line.append(f"ae {gs['ae_exposure']:.6f}/{cs['ae_exposure']:.6f} logsum {gs['log_sum']:.4f}/{cs['log_sum']:.4f}")
        line.append(f"[trace {gs['ms_trace']:.2f} taa {gs['ms_taa']:.2f} atrous {gs['ms_atrous']:.2f} expo {gs['ms_exposure']:.2f} cells {gs['ms_cells']:.2f}]")
        print(" ".join(line), flush=True)
    r.close(); o.close(); s.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["cornell", "mirror_spheres", "cylinders_disks_triangles", "boxes", "test", "volume_grid_test", "knot:60x16", "voxel_world:64x64"]
    for nm in which:
        pose = api.BENCH_POSE if nm.startswith("knot") else None
        try:
            run(nm, 60, 34, 1, frames=3, pose=pose)
            run(nm, 48, 14, 4, frames=2, pose=pose)
        except Exception as e:
            print(nm, "FAILED", repr(e), flush=True)
