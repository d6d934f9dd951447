"""Development aid: time one frame's stages with alternative builds of libycge (YCGE_LIB=...)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, %r)
import yetanotherconsolegameengine_b200 as pkg
s = pkg.HostScene("dragon")
r = pkg.CudaRaytraceRenderer(s, 480, 135, 4)
r.SetCamera(*pkg.BENCH_POSE)
for _ in range(4):
    r.TryFlipAndBlit()
st = r.stats()
print(os.path.basename(os.environ.get("YCGE_LIB", "libycge.so")), "trace %%.3f taa %%.3f atrous %%.3f (chain %%.3f) total %%.3f" %% (st["ms_trace"], st["ms_taa"], st["ms_atrous"], st["ms_atrous_chain"], st["ms_total"]), flush=True)
''' % ROOT
for lib in sys.argv[1:]:
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, YCGE_LIB=os.path.join(ROOT, "yetanotherconsolegameengine_b200", lib)))
