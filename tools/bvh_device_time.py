"""Development aid: wall time of ycge_mesh_build_device against ycge_mesh_upload_triangles (host build) for one mesh scene."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yetanotherconsolegameengine_b200 import api
scene = sys.argv[1] if len(sys.argv) > 1 else "dragon"
lib = api.load_lib()
s = api.HostScene(scene)
cfg = api.Config(); cfg.fb_w, cfg.fb_h, cfg.ss = 8, 4, 1
lib.ycge_default_params(C.byref(cfg.params))
ctx = C.c_void_p(); assert lib.ycge_create(C.byref(cfg), C.byref(ctx)) == 0
tris = np.ascontiguousarray(s.mesh_triangles(0), np.float32)
mat = s.mesh(0).contents.material
for rep in range(4):
    t0 = time.perf_counter(); assert lib.ycge_mesh_build_device(ctx, 2, len(tris), tris.ctypes.data, C.byref(mat)) == 0; t1 = time.perf_counter()
    print(f"{scene}: {len(tris)} triangles, device build call {1e3 * (t1 - t0):.2f} ms")
t0 = time.perf_counter(); assert lib.ycge_mesh_upload_triangles(ctx, 1, len(tris), tris.ctypes.data, C.byref(mat)) == 0; t1 = time.perf_counter()
print(f"host build + upload {1e3 * (t1 - t0):.2f} ms")
