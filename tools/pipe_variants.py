"""Development aid: serial stage times and pipelined frames/s with alternative builds of libycge (YCGE_LIB=...)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, time
sys.path.insert(0, %r)
import yetanotherconsolegameengine_b200 as pkg
s = pkg.HostScene("dragon")
r = pkg.CudaRaytraceRenderer(s, 480, 135, 4)
r.SetCamera(*pkg.BENCH_POSE)
for _ in range(4):
    r.TryFlipAndBlit()
st = r.stats()
out = "%%s serial: trace %%.3f atrous %%.3f (chain %%.3f) total %%.3f |" %% (os.path.basename(os.environ.get("YCGE_LIB", "libycge.so")), st["ms_trace"], st["ms_atrous"], st["ms_atrous_chain"], st["ms_total"])
for slots in [int(x) for x in os.environ.get("SLOTS", "1,2,3,4").split(",")]:
    r.pipeline_config(slots)
    r.render_frames_async(8); r.wait()
    t0 = time.perf_counter()
    r.render_frames_async(64); r.wait()
    out += " S%%d %%.1f fps" %% (slots, 64 / (time.perf_counter() - t0))
    st = r.stats()
    out += " [tr %%.2f taa %%.2f at %%.2f ch %%.2f]" %% (st["ms_trace"], st["ms_taa"], st["ms_atrous"], st["ms_atrous_chain"])
print(out, flush=True)
''' % ROOT
for lib in sys.argv[1:]:
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, YCGE_LIB=os.path.join(ROOT, "yetanotherconsolegameengine_b200", lib)))
