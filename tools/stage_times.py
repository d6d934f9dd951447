"""Development aid: serial stage times (best of 5 frames) and frames/s with 3 frames in flight.
python tools/stage_times.py SCENE:FBW:FBH:SS [...]      e.g. voxel_world:320:90:8 museum:480:135:4 dragon:480:135:4"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yetanotherconsolegameengine_b200 as pkg

for spec in sys.argv[1:]:
    scene, fbw, fbh, ss = spec.split(":")
    s = pkg.HostScene(scene)
    r = pkg.CudaRaytraceRenderer(s, int(fbw), int(fbh), int(ss))
    if scene in ("cow", "bunny", "teapot", "dragon"):
        r.SetCamera(*pkg.BENCH_POSE)
    for _ in range(4):
        r.TryFlipAndBlit()
    best = None
    for _ in range(5):
        r.TryFlipAndBlit()
        st = r.stats()
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    st = best
    out = "%s serial: trace %.3f taa %.3f atrous %.3f (chain %.3f) exposure %.3f total %.3f |" % (
        spec, st["ms_trace"], st["ms_taa"], st["ms_atrous"], st["ms_atrous_chain"], st["ms_exposure"], st["ms_total"])
    r.pipeline_config(3)
    r.render_frames_async(8); r.wait()
    t0 = time.perf_counter()
    r.render_frames_async(64); r.wait()
    out += " 3 in flight %.1f fps" % (64 / (time.perf_counter() - t0))
    print(out, flush=True)
    del r, s
