"""Development aid: throughput of a multi-device context (ycge_config.n_devices) and host time per submit."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import yetanotherconsolegameengine_b200 as pkg
from yetanotherconsolegameengine_b200 import api
devs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,0").split(",")]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
s = pkg.HostScene("dragon")
r = pkg.CudaRaytraceRenderer(s, 480, 135, 4, devices=devs)
r.SetCamera(*pkg.BENCH_POSE)
depth = 2 * len(devs)
ring = [torch.empty((135 * 480 * 32,), dtype=torch.uint8, pin_memory=True) for _ in range(depth)]
ring_np = [t.numpy().view(api.CELL_DTYPE).reshape(135, 480) for t in ring]
def run(k):
    ids, host = [], 0.0
    for i in range(k):
        if len(ids) == depth:
            r.frame_wait(ids.pop(0))
        t0 = time.perf_counter()
        ids.append(r.submit_frame(ring_np[i % depth]))
        host += time.perf_counter() - t0
    for fid in ids:
        r.frame_wait(fid)
    return host
run(4 * depth)
t0 = time.perf_counter()
host = run(frames)
dt = time.perf_counter() - t0
print("devices", devs, "frames/s %.1f" % (frames / dt), "host ms per submit %.3f" % (1e3 * host / frames))
