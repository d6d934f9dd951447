"""Development aid: per-band timestamps of the systolic wavefront kernel (YCGE_CHAIN_TRACE=<file>): duration of a band and the
lag between consecutive bands of equal row parity, in microseconds and in steps."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = "/tmp/wave_trace.bin"
os.environ["YCGE_CHAIN_TRACE"] = path
import yetanotherconsolegameengine_b200 as pkg
scene = sys.argv[1] if len(sys.argv) > 1 else "dragon"
fb_w, fb_h, ss = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (480, 135, 4)
s = pkg.HostScene(scene)
r = pkg.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
r.SetCamera(*pkg.BENCH_POSE)
for _ in range(3):
    r.TryFlipAndBlit()
st = r.stats()
full = np.fromfile(path, dtype=np.uint64).reshape(-1, 32).astype(np.int64)
np.save(os.path.join(ROOT, "gpurun_out", "wave_trace_full.npy"), full)
t = full[:, :2]
np.save(os.path.join(ROOT, "gpurun_out", "wave_trace.npy"), t)
t = t[(t[:, 0] > 0) & (t[:, 1] > 0)]
t0 = t[:, 0].min()
W = fb_w * ss
nt = (W + 1) // 2 + 11 + 8
dur = (t[:, 1] - t[:, 0]) / 1e3
print("bands", len(t), "kernel ms", st["ms_atrous_chain"], "span us", (t[:, 1].max() - t0) / 1e3)
print("band duration us: min %.1f median %.1f max %.1f -> step %.3f us" % (dur.min(), np.median(dur), dur.max(), np.median(dur) / nt))
print("per-band duration us (every 10th band of cy 0):", dur[0::20].round(0))
for cy in (0, 1):
    st_ = t[cy::2, 0]
    lag = np.diff(st_) / 1e3
    print("cy", cy, "start lag between bands us: median %.2f min %.2f max %.2f, first starts %s" % (np.median(lag), lag.min(), lag.max(), ((st_[:6] - t0) / 1e3).round(1)))
