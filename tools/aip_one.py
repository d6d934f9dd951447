"""Development aid: render a few frames of a W=1920, few-row image (the in-place pass runs with no row hand-offs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yetanotherconsolegameengine_b200 as pkg
fb_h = int(sys.argv[1]) if len(sys.argv) > 1 else 1
s = pkg.HostScene("cornell")
r = pkg.CudaRaytraceRenderer(s, 960, fb_h, 2)
for _ in range(3):
    r.TryFlipAndBlit()
print(r.stats())
