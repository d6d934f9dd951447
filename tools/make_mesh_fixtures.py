"""Generates tests/golden/meshes/*.ymesh from the reference's OBJ assets (run in the build container, where
/root/reference exists).  The .ymesh files hold exactly what the host mirror's MeshLoader.ParseObj produced (float32
positions as parsed, fan-triangulated face indices), so every machine loads bit-identical mesh input without the
reference checkout.  They are test/bench INPUT data, not reference source code."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yetanotherconsolegameengine_b200 import api  # noqa: E402

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/ConsoleGame/assets"
DST = os.path.join(ROOT, "tests", "golden", "meshes")
os.makedirs(DST, exist_ok=True)
h = api.load_host()
h.ycgeh_obj_to_ymesh.argtypes = [C.c_char_p, C.c_char_p]
for name in ("stanford-bunny", "teapot", "cow"):
    src, dst = os.path.join(SRC, name + ".obj"), os.path.join(DST, name + ".ymesh")
    rc = h.ycgeh_obj_to_ymesh(src.encode(), dst.encode())
    print(name, "->", dst, "rc", rc, os.path.getsize(dst) if rc == 0 else h.ycgeh_last_error().decode())
