"""Host mirror (libycge_host.so): the C# host side restated in C++ — scene factories of BuildSceneTable(), MeshLoader,
the BVH builders whose trees are uploaded, Framebuffer/Chexel, ANSITerminalRenderer.Render's byte stream."""
import numpy as np
import pytest

from yetanotherconsolegameengine_b200 import api
from oracle_binding import Oracle


SCENES = {  # name -> (objects, lights)   Scenes.cs:269-406, MeshScenes.cs:108-143
    "cornell": (8, 1), "mirror_spheres": (4, 2), "cylinders_disks_triangles": None, "boxes": (4, 2), "volume_grid_test": None,
    "bunny": (2, 2), "teapot": (2, 2), "cow": (2, 2), "texture_gallery": (6, 2),
}


@pytest.mark.parametrize("name", list(SCENES))
def test_scene_factories_build(name):
    s = api.HostScene(name)
    c = s.counts()
    assert c["objects"] > 0 and c["lights"] > 0
    if SCENES[name]:
        assert (c["objects"], c["lights"]) == SCENES[name]
    pos, yaw, pitch, fov = s.default_camera()
    assert fov == 45.0
    s.close()


def test_mesh_loader_counts_and_normalisation():
    for name, ntri in (("bunny", 69451), ("teapot", 6320), ("cow", 5804)):
        s = api.HostScene(name)
        assert s.counts()["triangles"] == ntri
        t = s.mesh_triangles(0).reshape(-1, 3)
        ext = t.max(0) - t.min(0)
        assert abs(float(ext.max()) - 1.0) < 1e-5  # NormalizeAllUsedVertices: unit max extent (MeshLoader.cs:107-148)
        assert float(t[:, 1].min()) > -0.05          # auto-ground: rests just above y = 0 (MeshScenes.cs:173-184)
        s.close()


def check_tree(tree, n_items, max_leaf):
    boxes, lrsc, leaf = tree["boxes"], tree["lrsc"], tree["leaf"]
    assert sorted(leaf.tolist()) == list(range(n_items)), "every item in exactly one leaf"
    seen = np.zeros(len(boxes), bool)
    stack = [tree["root"]]
    while stack:
        i = stack.pop()
        assert not seen[i]
        seen[i] = True
        l, r, st, cnt = lrsc[i]
        if cnt > 0:
            assert cnt <= max_leaf
            continue
        for ch in (l, r):
            if ch >= 0:
                assert np.all(boxes[ch, :3] >= boxes[i, :3]) and np.all(boxes[ch, 3:] <= boxes[i, 3:]), "child box inside parent box"
                stack.append(ch)
        kids = [ch for ch in (l, r) if ch >= 0]
        lo = np.min([boxes[k, :3] for k in kids], 0)
        hi = np.max([boxes[k, 3:] for k in kids], 0)
        assert np.array_equal(lo, boxes[i, :3]) and np.array_equal(hi, boxes[i, 3:]), "parent box = union of children"
    assert seen.all()


@pytest.mark.parametrize("name", ["cornell", "boxes", "cylinders_disks_triangles", "knot:40x12", "teapot", "voxel_world:64x64"])
def test_bvh_invariants_and_builder_parity(name):
    """Host-built trees (uploaded to the GPU) vs the oracle's own restatement of BVH.cs:258-459 / MeshBVH.cs:371-576."""
    s = api.HostScene(name)
    o = Oracle(s, 8, 4, 1, use_host_trees=False, mesh_form="triangles")
    top_h, top_o = s.bvh_arrays(-1), o.bvh_arrays(-1)
    check_tree(top_h, s.counts()["objects"], 4)
    for k in ("boxes", "lrsc", "leaf"):
        assert np.array_equal(top_h[k], top_o[k]), f"top-level {k}"
    assert top_h["root"] == top_o["root"]
    for m in range(s.n_meshes):
        mh, mo = s.bvh_arrays(m), o.bvh_arrays(m)
        check_tree(mh, s.mesh(m).contents.n_tris, 8)
        for k in ("boxes", "lrsc", "leaf"):
            assert np.array_equal(mh[k], mo[k]), f"mesh {m} {k}"
    o.close()
    s.close()


def ansi_render_py(cells):  # ANSITerminalRenderer.Render (ANSITerminalRenderer.cs:86-153), without the resize prologue
    out = bytearray()
    cur_f = cur_b = -1
    for y in range(cells.shape[0]):
        out += b"\x1b[%d;1H" % (y + 1)
        for x in range(cells.shape[1]):
            c = cells[y, x]
            f, b = int(c["fg_ansi"]), int(c["bg_ansi"])
            if f != cur_f and b != cur_b:
                out += b"\x1b[38;5;%d;48;5;%dm" % (f, b)
                cur_f, cur_b = f, b
            elif f != cur_f:
                out += b"\x1b[38;5;%dm" % f
                cur_f = f
            elif b != cur_b:
                out += b"\x1b[48;5;%dm" % b
                cur_b = b
            out += chr(int(c["glyph"])).encode("utf-8")
    out += b"\x1b[0m"
    return bytes(out)


def test_ansi_byte_stream():
    rng = np.random.default_rng(7)
    cells = np.zeros((9, 31), api.CELL_DTYPE)
    cells["glyph"] = 0x2580
    cells["fg_ansi"] = rng.integers(16, 232, cells.shape)
    cells["bg_ansi"] = rng.integers(16, 232, cells.shape)
    cells["fg_ansi"][2, :] = 20  # runs: no escape when nothing changes
    cells["bg_ansi"][2, 5:] = 21
    cells["glyph"][0, 0] = ord("A")
    cells["glyph"][0, 1] = 0x00E9
    got = api.ansi_from_cells(cells)
    assert got == ansi_render_py(cells)
    assert got.endswith(b"\x1b[0m") and got.startswith(b"\x1b[1;1H")


def test_vg01_world_file_round_trip(tmp_path):
    """the 'VG01' world file (WorldManager.cs:399-440 reader, :612-629 writer).  A file written in the
    writer's layout loads into the same chunk VolumeGrids, top-level tree and camera as the in-memory builder."""
    path = str(tmp_path / "world.vg01")
    assert api.load_host().ycgeh_write_synthetic_world(path.encode(), 64, 64) == 0
    raw = open(path, "rb").read()
    assert raw[:4] == b"VG01" and np.frombuffer(raw[4:16], np.int32).tolist() == [64, 64, 64]
    assert len(raw) == 16 + 64 * 64 * 64 * 8
    a = api.HostScene("voxel_world:64x64")
    b = api.HostScene("voxel_world_file:" + path)
    assert a.n_volumes == b.n_volumes and a.n_volumes > 0
    for i in range(a.n_volumes):
        va, vb = a.volume(i).contents, b.volume(i).contents
        assert (va.nx, va.ny, va.nz) == (vb.nx, vb.ny, vb.nz) and list(va.min_corner) == list(vb.min_corner)
        n = ((va.nx + 7) // 8) * ((va.ny + 7) // 8) * ((va.nz + 7) // 8) * 512
        assert np.array_equal(np.ctypeslib.as_array(va.mat, (n,)), np.ctypeslib.as_array(vb.mat, (n,)))
        assert np.array_equal(np.ctypeslib.as_array(va.meta, (n,)), np.ctypeslib.as_array(vb.meta, (n,)))
    ta, tb = a.bvh_arrays(-1), b.bvh_arrays(-1)
    for k in ("boxes", "lrsc", "leaf"):
        assert np.array_equal(ta[k], tb[k])
    a.close(); b.close()
    # the reader's errors (WorldManager.cs:403-417)
    bad = str(tmp_path / "bad.vg01")
    open(bad, "wb").write(b"VG02" + raw[4:])
    with pytest.raises(Exception, match="VG01"):
        api.HostScene("voxel_world_file:" + bad)
    open(bad, "wb").write(raw[: len(raw) // 2])
    with pytest.raises(Exception, match="[Tt]runcated"):
        api.HostScene("voxel_world_file:" + bad)
    with pytest.raises(Exception, match="not found"):
        api.HostScene("voxel_world_file:" + str(tmp_path / "missing.vg01"))


def test_texture_test_scene_and_png_decoder(tmp_path):
    """BuildTextureTestScene (Scenes.cs:337-358): one textured box, ambient 0.5, no lights.  new Texture(path) decodes through
    OpenCV in the reference (ImreadModes.Color, BGR2RGBA: Texture.cs:25-49); the mirror's zlib-only PNG decoder must give the
    same pixels as an independent decoder for every colour type and row filter PNG writers emit."""
    s = api.HostScene("texture_test")
    assert s.name == "texture_test-standin" and s.n_textures == 1 and s.counts()["objects"] == 1 and s.counts()["lights"] == 0
    m = s.flat.contents.materials[s.flat.contents.objects[0].mat_a]
    assert (m.tex_id, m.tex_weight, m.uv_scale) == (0, 1.0, 1.0) and list(m.albedo) == [0.5, 0.5, 0.5]
    s.close()
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, size=(19, 23, 4), dtype=np.uint8)
    base[:, :12] = (base[:, :12] // 64) * 64  # flat runs make the writer pick different row filters
    for mode in ("RGB", "RGBA", "L", "LA", "P"):
        im = Image.fromarray(base, "RGBA").convert(mode)
        d = tmp_path / mode
        d.mkdir()
        im.save(str(d / "image.png"))
        rgb = np.asarray(im.convert("RGB")).astype(np.uint32)
        want = rgb[..., 0] | (rgb[..., 1] << 8) | (rgb[..., 2] << 16) | np.uint32(255 << 24)
        t = api.HostScene("texture_test", asset_dir=str(d))
        assert t.name == "texture_test"
        assert np.array_equal(t.texture(0), want), mode
        t.set_texture(0, want[:5, :7])
        assert t.texture(0).shape == (5, 7)
        t.close()
