"""MeshLoader.FromObj (RayTracing/MeshLoader.cs:11-97: fan triangulation, NormalizeAllUsedVertices :105-149, scale + translate) and
MeshScenes.AddMeshAutoGround (Scenes/MeshScenes.cs:173-184) with TryReadObjBoundsNormalized (:186-331: the largest connected
component, centred on its area-unweighted triangle centroid mean, scaled to unit extent -- a DIFFERENT normalisation, used only
for the ground height) transcribed into numpy binary32 arithmetic; the triangles A, B, C of the cow, the bunny and the teapot
must equal the host mirror's -- the ones uploaded to the GPU -- bit for bit.  Where the reference checkout is present, the OBJ
text parser is checked too (vertex and face counts, every coordinate) against the committed binary twins of the assets.
"""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN
from yetanotherconsolegameengine_b200 import api

F = np.float32
ASSETS = "/root/reference/ConsoleGame/assets"


def read_ymesh(path):
    b = open(path, "rb").read()
    assert b[:4] == b"YMSH"
    nv, nf = struct.unpack_from("<ii", b, 4)
    xyz = np.frombuffer(b, np.float32, nv * 3, 12).reshape(nv, 3).copy()
    faces = np.frombuffer(b, np.int32, nf, 12 + nv * 12).reshape(-1, 3).copy()
    return xyz, faces


def parse_obj(path):  # MeshLoader.cs:23-56
    pos, faces = [], []
    for line in open(path):
        if not line.strip() or line[0] == "#":
            continue
        tok = line.split()
        if tok[0] == "v" and len(tok) >= 4:
            pos.append([F(float(tok[1])), F(float(tok[2])), F(float(tok[3]))])
        elif tok[0] == "f" and len(tok) >= 4:
            idx = []
            for t in tok[1:]:
                s = t.split("/")[0]
                i = int(s) if s else 0               # ParseIndex :99-104 (an empty token gives index 0)
                idx.append((i - 1 if i > 0 else len(pos) + i) if s else 0)
            for i in range(2, len(idx)):
                faces.append([idx[0], idx[i - 1], idx[i]])
    return np.array(pos, F), np.array(faces, np.int32)


def normalize_all_used(pos, faces, target=F(1.0)):  # NormalizeAllUsedVertices :105-149
    used = np.unique(faces)
    mn, mx = pos[used].min(0), pos[used].max(0)
    c = (mn + mx) * F(0.5)
    ext = max(F(mx[0] - mn[0]), F(mx[1] - mn[1]), F(mx[2] - mn[2]))
    if ext <= 0:
        ext = F(1)
    s = F(target / ext)
    return ((pos - c) * s).astype(F)


def bounds_normalized_largest_component(pos, faces):  # TryReadObjBoundsNormalized :186-331
    parent, rank = list(range(len(pos))), [0] * len(pos)

    def find(x):
        while x != parent[x]:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    def union(x, y):
        rx, ry = find(x), find(y)
        if rx == ry:
            return
        if rank[rx] < rank[ry]:
            parent[rx] = ry
        elif rank[rx] > rank[ry]:
            parent[ry] = rx
        else:
            parent[ry] = rx
            rank[rx] += 1

    for a, b, c in faces.tolist():
        union(a, b)
        union(b, c)
    comp = {}
    for i, f in enumerate(faces.tolist()):
        comp.setdefault(find(f[0]), []).append(i)  # a Dictionary without removals enumerates in insertion order: first maximum wins
    best = max(comp.values(), key=len)
    kept = faces[np.array(best)]
    A, B, C = pos[kept[:, 0]], pos[kept[:, 1]], pos[kept[:, 2]]
    third = F(F(1) / F(3))
    terms = (((A + B).astype(F) + C).astype(F) * third).astype(F)          # (A.X + B.X + C.X) * (1.0f / 3.0f)
    centre = np.cumsum(terms, axis=0, dtype=F)[-1]                         # cx += ... in face order, binary32
    centre = (centre * F(F(1) / F(len(kept)))).astype(F)
    rel = (pos[np.unique(kept)] - centre).astype(F)
    rmin, rmax = rel.min(0), rel.max(0)
    ext = max(F(rmax[0] - rmin[0]), F(rmax[1] - rmin[1]), F(rmax[2] - rmin[2]))
    if ext <= 0:
        ext = F(1)
    s = F(F(1) / ext)
    return (rmin * s).astype(F), (rmax * s).astype(F)


@pytest.mark.parametrize("scene,asset", [("teapot", "teapot"), ("cow", "cow"), ("bunny", "stanford-bunny")])
def test_mesh_loader_and_auto_ground_match_a_literal_transcription(scene, asset):
    xyz, faces = read_ymesh(os.path.join(GOLDEN, "meshes", asset + ".ymesh"))
    obj = os.path.join(ASSETS, asset + ".obj")
    if os.path.exists(obj):  # the reference checkout is here: the text parser against the committed binary twin
        p2, f2 = parse_obj(obj)
        assert p2.shape == xyz.shape and np.array_equal(p2.view(np.uint32), xyz.view(np.uint32)), "OBJ vertex parse"
        assert np.array_equal(f2, faces), "OBJ fan triangulation"
    target = np.array([0.0, 0.5, 1.0], F)                                   # targetPos of the mesh scenes (MeshScenes.cs:108-133)
    scale = F(1.0)
    mn_n, _ = bounds_normalized_largest_component(xyz, faces)
    y_translate = F(F(target[1] - F(mn_n[1] * scale)) + F(0.01))            # AddMeshAutoGround :180-182
    t = np.array([target[0], y_translate, target[2]], F)
    pos = normalize_all_used(xyz, faces)
    pos = ((pos * scale).astype(F) + t).astype(F)                           # FromObj :62-68 (t != 0)
    tris = np.concatenate([pos[faces[:, 0]], pos[faces[:, 1]], pos[faces[:, 2]]], axis=1)
    s = api.HostScene(scene)
    got = s.mesh_triangles(0)
    assert got.shape == tris.shape
    assert np.array_equal(got.view(np.uint32), tris.view(np.uint32)), f"{scene}: {int((got != tris).any(1).sum())} triangles differ"
    s.close()


REF_SO = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libycge_ref.so")


@pytest.mark.parametrize("scene,asset", [("teapot", "teapot"), ("cow", "cow"), ("bunny", "stanford-bunny")])
def test_mesh_loader_equals_the_reference_source(scene, asset, tmp_path):
    """The same triangles from THE REFERENCE'S OWN TEXT: MeshLoader.cs (FromObj with its OBJ reader, ParseIndex,
    NormalizeAllUsedVertices) and MeshScenes.TryReadObjBoundsNormalized + the ground placement of AddMeshAutoGround, rewritten
    into C++ syntactically at build time (oracle/ref_transpile.py -> oracle/_ref), read an OBJ file and return the triangles a mesh
    scene gets; the host mirror's -- the ones uploaded to the GPU -- must equal them bit for bit.  The OBJ text is written here
    from the committed binary twin of the asset (9 significant digits: every binary32 survives the round trip), or is the
    reference's own file where the checkout is present."""
    import ctypes as C
    so = os.path.normpath(REF_SO)
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libycge_ref.so is not built (needs /root/reference at build time)")
    ref = C.CDLL(so)
    ref.ref_mesh_from_obj.argtypes = [C.c_char_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    xyz, faces = read_ymesh(os.path.join(GOLDEN, "meshes", asset + ".ymesh"))
    path = str(tmp_path / (asset + ".obj"))
    with open(path, "w") as f:
        f.write("# written from the binary twin\n")
        for v in xyz:
            f.write("v %.9g %.9g %.9g\n" % (float(v[0]), float(v[1]), float(v[2])))
        for a, b, c in faces.tolist():
            f.write("f %d/%d %d//%d %d\n" % (a + 1, a + 1, b + 1, b + 1, c + 1))     # v/vt, v//vn and plain v: ParseIndex takes what precedes the first '/'
    paths = [path]
    if os.path.exists(os.path.join(ASSETS, asset + ".obj")):
        paths.append(os.path.join(ASSETS, asset + ".obj"))                           # the reference's own file (quads fan-triangulated by FromObj itself)
    s = api.HostScene(scene)
    want = s.mesh_triangles(0)
    target = np.array([0.0, 0.5, 1.0], F)                                            # targetPos of the mesh scenes (MeshScenes.cs:108-133), scale 1
    for pth in paths:
        n = ref.ref_mesh_from_obj(pth.encode(), 1.0, target.ctypes.data, None, 0, None)
        assert n == len(want), (pth, n, len(want))
        got, tr = np.empty((n, 9), F), np.empty(3, F)
        assert ref.ref_mesh_from_obj(pth.encode(), 1.0, target.ctypes.data, got.ctypes.data, n, tr.ctypes.data) == n
        assert tr[0] == 0.0 and tr[2] == 1.0 and 0.0 < tr[1] < 2.0
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{scene} ({pth}): {int((got != want).any(1).sum())} triangles differ"
    s.close()
