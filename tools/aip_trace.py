"""Development aid: timestamps of the wavefront kernel's chains (YCGE_CHAIN_TRACE) -> per-row start/step/lag statistics."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["YCGE_CHAIN_TRACE"] = "/tmp/chain_trace.bin"
import yetanotherconsolegameengine_b200 as pkg
fb_h = int(sys.argv[1]) if len(sys.argv) > 1 else 270
s = pkg.HostScene("cornell")
r = pkg.CudaRaytraceRenderer(s, 960, fb_h, 2)
for _ in range(3):
    r.TryFlipAndBlit()
st = r.stats()
H = fb_h * 4
t = np.fromfile("/tmp/chain_trace.bin", np.uint64).reshape(H, 2, 32).astype(np.float64)
t0 = t[t > 0].min()
t = np.where(t > 0, t - t0, np.nan) / 1e3  # us
print("chain kernel %.3f ms" % st["ms_atrous_chain"])
for y in list(range(0, min(H, 12))) + list(range(100, min(H, 104))) + list(range(max(0, H - 4), H)):
    row = t[y, 0]
    print("row %4d c0: start %8.2f us  step64 deltas(us): %s" % (y, row[0], " ".join("%.1f" % d for d in np.diff(row[:15]))))
start = t[:, 0, 0]
d = start[2:] - start[:-2]
print("start lag row y vs y-2: mean %.2f us  median %.2f  (per 64 steps a chain takes median %.2f us)" % (np.nanmean(d), np.nanmedian(d), np.nanmedian(np.diff(t[:, 0, :15], axis=1))))
mid = t[:, 0, 7]
d = mid[2:] - mid[:-2]
print("lag at step 448: mean %.2f median %.2f" % (np.nanmean(d), np.nanmedian(d)))
