"""Development aid: how the à-trous time scales with rows (at W = 1920) and with width (at 4 rows): separates the
per-step time of a chain from the row-to-row hand-off lag of the in-place wavefront pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yetanotherconsolegameengine_b200 as pkg
s = pkg.HostScene(sys.argv[1] if len(sys.argv) > 1 else "cornell")
def run(fb_w, fb_h, ss=2):
    r = pkg.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
    for _ in range(3):
        r.TryFlipAndBlit()
    st = r.stats()
    print(f"W={fb_w*ss:5d} H={fb_h*2*ss:5d}: atrous {st['ms_atrous']:.3f} ms  trace {st['ms_trace']:.3f}  taa {st['ms_taa']:.3f} total {st['ms_total']:.3f}", flush=True)
    r.close()
for fb_w in (60, 120, 240, 480, 960):
    run(fb_w, 1)
for fb_h in (2, 9, 33, 135, 270):
    run(960, fb_h)
