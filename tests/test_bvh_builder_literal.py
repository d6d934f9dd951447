"""The two tree builders pinned by a third implementation in another language: BVH.BuildRecursive (Objects/BVH.cs:258-459, leaf <= 4,
16 SAH bins, and its quirk: the partition re-derives origin / extent from the UNSORTED first and last items, :394-396) and
MeshBVH.BuildRecursive (Objects/MeshBVH.cs:371-576, leaf <= 8, consistent pivot), with every TryGetBounds that feeds them,
transcribed from the C# source into numpy binary32 scalars.  The node arrays (boxes, left / right / start / count), the root and
the leaf order must equal the host mirror's trees -- the ones the GPU traverses -- field by field; the oracle's own builder is
compared with the mirror's in test_host.py.  Array.Sort (the builders' fallback) is reached by two scenes, on ranges of at most
16 items, where .NET's introsort is an insertion sort: that path is transcribed too (Builder.sort_range).
"""
import ctypes as C
import os

import numpy as np
import pytest

from yetanotherconsolegameengine_b200 import api
from oracle_binding import load_oracle

F = np.float32
INF = F(np.inf)
BINS = 16


def surface_area(b):  # BVH.cs:462-466
    dx, dy, dz = F(b[3] - b[0]), F(b[4] - b[1]), F(b[5] - b[2])
    return F(F(2) * F(F(F(dx * dy) + F(dx * dz)) + F(dy * dz)))


def surround(box, o):  # BVH.cs:253-257
    for k in range(3):
        if o[k] < box[k]:
            box[k] = o[k]
        if o[3 + k] > box[3 + k]:
            box[3 + k] = o[3 + k]


def intro_sort(keys, key):
    """ArraySortHelper<T>.IntrospectiveSort (dotnet/runtime, .NET 8, ArraySortHelper.cs) on a Python list, in place: depth limit
    2 * (floor(log2 n) + 1); partitions of <= 16 finish with SwapIfGreater pairs or an insertion sort; median-of-three pivot moved
    to hi - 1; heap sort when the depth limit is spent.  Unstable, which is the point: equal centroids come out in THIS order."""
    def gt(i, j):
        return key(keys[i]) > key(keys[j])

    def swap_if_greater(i, j):
        if gt(i, j):
            keys[i], keys[j] = keys[j], keys[i]

    def insertion(lo, n):
        for i in range(lo, lo + n - 1):
            t = keys[i + 1]
            j = i
            while j >= lo and key(t) < key(keys[j]):
                keys[j + 1] = keys[j]
                j -= 1
            keys[j + 1] = t

    def down_heap(i, n, lo):
        d = keys[lo + i - 1]
        while i <= n // 2:
            child = 2 * i
            if child < n and key(keys[lo + child - 1]) < key(keys[lo + child]):
                child += 1
            if not key(d) < key(keys[lo + child - 1]):
                break
            keys[lo + i - 1] = keys[lo + child - 1]
            i = child
        keys[lo + i - 1] = d

    def heap_sort(lo, n):
        for i in range(n // 2, 0, -1):
            down_heap(i, n, lo)
        for i in range(n, 1, -1):
            keys[lo], keys[lo + i - 1] = keys[lo + i - 1], keys[lo]
            down_heap(1, i - 1, lo)

    def partition(lo, n):
        hi = n - 1
        mid = hi >> 1
        swap_if_greater(lo, lo + mid)
        swap_if_greater(lo, lo + hi)
        swap_if_greater(lo + mid, lo + hi)
        pivot = keys[lo + mid]
        keys[lo + mid], keys[lo + hi - 1] = keys[lo + hi - 1], keys[lo + mid]
        left, right = 0, hi - 1
        while left < right:
            left += 1
            while key(keys[lo + left]) < key(pivot):
                left += 1
            right -= 1
            while key(pivot) < key(keys[lo + right]):
                right -= 1
            if left >= right:
                break
            keys[lo + left], keys[lo + right] = keys[lo + right], keys[lo + left]
        if left != hi - 1:
            keys[lo + left], keys[lo + hi - 1] = keys[lo + hi - 1], keys[lo + left]
        return left

    def intro(lo, n, depth):
        while n > 1:
            if n <= 16:
                if n == 2:
                    swap_if_greater(lo, lo + 1)
                elif n == 3:
                    swap_if_greater(lo, lo + 1)
                    swap_if_greater(lo, lo + 2)
                    swap_if_greater(lo + 1, lo + 2)
                else:
                    insertion(lo, n)
                return
            if depth == 0:
                heap_sort(lo, n)
                return
            depth -= 1
            p = partition(lo, n)
            intro(lo + p + 1, n - (p + 1), depth)
            n = p

    if len(keys) > 1:
        intro(0, len(keys), 2 * (int(len(keys)).bit_length() - 1 + 1))


class Builder:
    def __init__(self, items, leaf_size, consistent_pivot, sort_lib):
        self.arr = items  # list of dicts: index, box[6], c[3]
        self.leaf_size, self.consistent, self.sort_lib = leaf_size, consistent_pivot, sort_lib
        self.nodes, self.leaf, self.sorts, self.sort_sizes = [], [], 0, []
        self.root = self.build(0, len(items))

    def sort_range(self, start, count, axis):
        """Array.Sort(arr, start, count, by centroid[axis]).  Up to 16 elements .NET's IntroSort (ArraySortHelper<T>.IntroSort,
        dotnet/runtime, .NET 8) is two or three SwapIfGreater calls or an insertion sort; those paths are transcribed here, so the
        trees of the scenes below confirm them independently of the oracle.  Longer ranges (the museum's 62 objects reach one) go
        through the transcription of the full introsort above AND the oracle's restatement, which must agree."""
        self.sorts += 1
        self.sort_sizes.append(count)
        key = lambda it: it["c"][axis]
        a = self.arr

        def swap_if_greater(i, j):
            if key(a[start + i]) > key(a[start + j]):
                a[start + i], a[start + j] = a[start + j], a[start + i]

        if count < 2:
            return
        if count == 2:
            swap_if_greater(0, 1)
        elif count == 3:
            swap_if_greater(0, 1)
            swap_if_greater(0, 2)
            swap_if_greater(1, 2)
        elif count <= 16:
            for i in range(count - 1):  # InsertionSort
                t = a[start + i + 1]
                j = i
                while j >= 0 and key(t) < key(a[start + j]):
                    a[start + j + 1] = a[start + j]
                    j -= 1
                a[start + j + 1] = t
        else:
            keys = np.array([key(a[start + i]) for i in range(count)], F)
            payload = np.arange(count, dtype=np.int32)
            self.sort_lib.yo_dotnet_sort_floats(keys.ctypes.data_as(C.c_void_p), payload.ctypes.data_as(C.c_void_p), count)
            by_oracle = [a[start + int(j)] for j in payload]
            seg = a[start:start + count]
            intro_sort(seg, key)
            assert [it["index"] for it in seg] == [it["index"] for it in by_oracle], "the oracle's introsort and the transcription below disagree"
            a[start:start + count] = seg

    def build(self, start, count):
        arr = self.arr
        if count <= 0:
            return -1
        if count <= self.leaf_size:
            box = list(arr[start]["box"])
            for i in range(1, count):
                surround(box, arr[start + i]["box"])
            base = len(self.leaf)
            self.leaf += [arr[start + i]["index"] for i in range(count)]
            self.nodes.append(dict(box=box, left=-1, right=-1, start=base, count=count))
            return len(self.nodes) - 1
        cmin, cmax = list(arr[start]["c"]), list(arr[start]["c"])
        for i in range(start + 1, start + count):
            for k in range(3):
                c = arr[i]["c"][k]
                if c < cmin[k]:
                    cmin[k] = c
                if c > cmax[k]:
                    cmax[k] = c
        ext = [F(cmax[k] - cmin[k]) for k in range(3)]
        axis = 0
        if ext[1] > ext[0] and ext[1] >= ext[2]:
            axis = 1
        elif ext[2] > ext[0] and ext[2] >= ext[1]:
            axis = 2
        split_bin, best_axis, best_cost = -1, axis, INF
        for ax in range(3):
            extent = ext[ax]
            if not extent > 0:
                continue
            origin, inv_extent = cmin[ax], F(F(1) / extent)
            counts = [0] * BINS
            boxes = [[INF, INF, INF, -INF, -INF, -INF] for _ in range(BINS)]
            for i in range(start, start + count):
                b = int(F(F(F(arr[i]["c"][ax] - origin) * inv_extent) * F(BINS - 1)))
                b = min(max(b, 0), BINS - 1)
                counts[b] += 1
                surround(boxes[b], arr[i]["box"])
            left_count, right_count, left_area, right_area = [0] * BINS, [0] * BINS, [F(0)] * BINS, [F(0)] * BINS
            cur, acc = [INF, INF, INF, -INF, -INF, -INF], 0
            with np.errstate(invalid="ignore", over="ignore"):
                for b in range(BINS):
                    if counts[b] > 0:
                        surround(cur, boxes[b])
                    acc += counts[b]
                    left_count[b], left_area[b] = acc, surface_area(cur)
                cur, acc = [INF, INF, INF, -INF, -INF, -INF], 0
                for b in range(BINS - 1, -1, -1):
                    if counts[b] > 0:
                        surround(cur, boxes[b])
                    acc += counts[b]
                    right_count[b], right_area[b] = acc, surface_area(cur)
                for b in range(BINS - 1):
                    lc, rc = left_count[b], right_count[b + 1]
                    if lc == 0 or rc == 0:
                        continue
                    cost = F(F(left_area[b] * F(lc)) + F(right_area[b + 1] * F(rc)))
                    if cost < best_cost:
                        best_cost, best_axis, split_bin = cost, ax, b
        if split_bin < 0:
            self.sort_range(start, count, best_axis)
            mid = start + (count >> 1)
        else:
            if self.consistent:  # MeshBVH.cs:511-513
                origin, extent = cmin[best_axis], ext[best_axis]
                inv_extent = F(F(1) / extent)
            else:                # BVH.cs:394-396: from the unsorted first / last item
                origin = arr[start]["c"][best_axis]
                extent = F(arr[start + count - 1]["c"][best_axis] - origin)
                inv_extent = F(F(1) / extent) if extent != 0 else F(0)
            i0, i1 = start, start + count - 1
            while i0 <= i1:
                c0 = arr[i0]["c"][best_axis]
                b0 = int(F(F(F(c0 - origin) * inv_extent) * F(BINS - 1))) if (self.consistent or inv_extent != 0) else 0
                if b0 <= split_bin:
                    i0 += 1
                else:
                    arr[i0], arr[i1] = arr[i1], arr[i0]
                    i1 -= 1
            mid = i0
            if mid == start or mid == start + count:
                self.sort_range(start, count, best_axis)
                mid = start + (count >> 1)
        my = len(self.nodes)
        self.nodes.append(None)
        left = self.build(start, mid - start)
        right = self.build(mid, start + count - mid)
        if left >= 0 and right >= 0:
            lb, rb = self.nodes[left]["box"], self.nodes[right]["box"]
            box = [min(lb[k], rb[k]) for k in range(3)] + [max(lb[3 + k], rb[3 + k]) for k in range(3)]
        else:
            box = list(self.nodes[left if left >= 0 else right]["box"])
        self.nodes[my] = dict(box=box, left=left, right=right, start=0, count=0)
        return my


def assert_tree_equal(b, tree, what):
    assert b.root == tree["root"], what
    assert len(b.nodes) == len(tree["boxes"]), (what, len(b.nodes), len(tree["boxes"]))
    got_boxes = np.array([n["box"] for n in b.nodes], F)
    assert np.array_equal(got_boxes.view(np.uint32), np.ascontiguousarray(tree["boxes"]).view(np.uint32)), what + ": node boxes"
    got = np.array([[n["left"], n["right"], n["start"], n["count"]] for n in b.nodes], np.int32)
    assert np.array_equal(got, tree["lrsc"]), what + ": left / right / start / count"
    assert np.array_equal(np.array(b.leaf, np.int32), tree["leaf"]), what + ": leaf order"


def triangle_item(i, t):  # MeshBVH.TryComputeBounds :336-349 and the centroid :56-58; Triangle ctor :55-66 uses the same rule
    a, b, c = t[0:3], t[3:6], t[6:9]
    e = F(1e-4)
    box = [F(min(a[k], min(b[k], c[k])) - e) for k in range(3)] + [F(max(a[k], max(b[k], c[k])) + e) for k in range(3)]
    return dict(index=i, box=box, c=[F(F(0.5) * F(box[k] + box[3 + k])) for k in range(3)])


def object_item(i, o, mesh_roots, volumes):  # TryGetBounds of every Hittable
    p = [F(v) for v in o.p]
    e = F(1e-4)
    if o.kind == 1:  # Plane, Surfaces.cs:30-36: a +-1e6 box with centroid exactly 0
        return dict(index=i, box=[F(-1e6)] * 3 + [F(1e6)] * 3, c=[F(0)] * 3)
    if o.kind in (0, 2):  # Sphere BoundedObjects.cs:20-28, Disk Surfaces.cs:96-105: centre -+ (r, r, r)
        r = p[3] if o.kind == 0 else p[6]
        box = [F(p[k] - r) for k in range(3)] + [F(p[k] + r) for k in range(3)]
    elif o.kind == 3:  # XYRect :175-181
        box = [p[0], p[2], F(p[4] - e), p[1], p[3], F(p[4] + e)]
    elif o.kind == 4:  # XZRect :247-253
        box = [p[0], F(p[4] - e), p[2], p[1], F(p[4] + e), p[3]]
    elif o.kind == 5:  # YZRect :319-325
        box = [F(p[4] - e), p[0], p[2], F(p[4] + e), p[1], p[3]]
    elif o.kind == 6:  # Box BoundedObjects.cs:92-97
        box = p[0:6]
    elif o.kind == 7:  # CylinderY :140-145
        box = [F(p[0] - p[3]), p[4], F(p[2] - p[3]), F(p[0] + p[3]), p[5], F(p[2] + p[3])]
    elif o.kind == 8:  # Triangle
        return dict(triangle_item(i, p[0:9]), index=i)
    elif o.kind == 9:  # Mesh -> MeshBVH.TryGetBounds: the root node's box
        box = [F(v) for v in mesh_roots[o.ref_id]]
    elif o.kind == 10:  # VolumeGrid.TryGetBounds (VolumeGrid.cs:386-403): minCorner .. minCorner + n * voxelSize
        v = volumes[o.ref_id]
        box = [F(v.min_corner[k]) for k in range(3)] + [F(F(v.min_corner[k]) + F(F(getattr(v, "nx ny nz".split()[k])) * F(v.voxel_size[k]))) for k in range(3)]
    else:
        raise NotImplementedError(o.kind)
    return dict(index=i, box=list(box), c=[F(F(0.5) * F(box[k] + box[3 + k])) for k in range(3)])


@pytest.mark.parametrize("scene_name", ["knot:12x5", "knot:40x10", "teapot"])
def test_mesh_builder_matches_a_literal_python_transcription(scene_name):
    lib = load_oracle()
    lib.yo_set_sort_mode(0)
    s = api.HostScene(scene_name)
    tris = s.mesh_triangles(0)
    items = [triangle_item(i, [F(v) for v in tris[i].reshape(-1)]) for i in range(len(tris))]
    b = Builder(items, leaf_size=8, consistent_pivot=True, sort_lib=lib)
    assert_tree_equal(b, s.bvh_arrays(0), scene_name)
    assert b.sorts == s.bvh_arrays(0)["sort_fallbacks"], "the sort fallback must be reached exactly as often as in the mirror's build"
    s.close()


@pytest.mark.parametrize("scene_name", ["cornell", "mirror_spheres", "cylinders_disks_triangles", "boxes", "test", "volume_grid_test", "teapot",
                                        "voxel_world:64x64", "texture_gallery", "all_meshes:40x10", "voxel_island:64x128", "museum"])
def test_top_level_builder_matches_a_literal_python_transcription(scene_name):
    lib = load_oracle()
    lib.yo_set_sort_mode(0)
    s = api.HostScene(scene_name)
    flat = s.flat.contents
    mesh_roots = {}
    for i in range(s.n_meshes):
        t = s.bvh_arrays(i)
        mesh_roots[i] = t["boxes"][t["root"]]
    volumes = {i: s.volume(i).contents for i in range(s.n_volumes)}
    items = [object_item(i, flat.objects[i], mesh_roots, volumes) for i in range(flat.n_objects)]
    b = Builder(items, leaf_size=4, consistent_pivot=False, sort_lib=lib)
    tree = s.bvh_arrays(-1)
    assert_tree_equal(b, tree, scene_name)
    assert b.sorts == tree["sort_fallbacks"]
    if scene_name == "museum":
        assert max(b.sort_sizes) > 16, "the museum is the scene that reaches the long-range introsort"
    if scene_name in ("cornell", "volume_grid_test"):
        assert b.sorts > 0  # these two scenes do reach Array.Sort (identical centroids along every axis with extent)
    s.close()


@pytest.mark.parametrize("n,distinct", [(17, 3), (64, 5), (333, 7), (2000, 11), (5000, 4000)])
def test_introsort_transcription_equals_the_oracle_restatement(n, distinct):
    """The same unstable order from both restatements of Array.Sort on keys with many ties (what decides the tree when centroids
    coincide): the permutation, not just the sorted keys."""
    lib = load_oracle()
    rng = np.random.default_rng(n)
    keys = rng.integers(0, distinct, n).astype(F)
    payload = np.arange(n, dtype=np.int32)
    k2 = keys.copy()
    lib.yo_dotnet_sort_floats(k2.ctypes.data_as(C.c_void_p), payload.ctypes.data_as(C.c_void_p), n)
    items = [dict(index=i, k=keys[i]) for i in range(n)]
    intro_sort(items, lambda it: it["k"])
    assert [it["index"] for it in items] == payload.tolist()
    assert (np.diff(k2) >= 0).all()
    if distinct < n // 4:
        assert payload.tolist() != sorted(range(n), key=lambda i: (keys[i], i)), "the order must differ from a stable sort, or the case proves nothing"


def _mesh_tree(make_scene, serial, task_items=None):
    old = os.environ.pop("YCGE_HOST_SERIAL_BVH", None)
    os.environ.pop("YCGE_HOST_BVH_TASK_ITEMS", None)
    if serial:
        os.environ["YCGE_HOST_SERIAL_BVH"] = "1"
    if task_items:
        os.environ["YCGE_HOST_BVH_TASK_ITEMS"] = str(task_items)
    try:
        s = make_scene()
        t = s.bvh_arrays(0)
        out = dict(root=t["root"], boxes=t["boxes"].copy(), lrsc=t["lrsc"].copy(), leaf=t["leaf"].copy(), sort_fallbacks=t["sort_fallbacks"])
        s.close()
        return out
    finally:
        os.environ.pop("YCGE_HOST_SERIAL_BVH", None)
        os.environ.pop("YCGE_HOST_BVH_TASK_ITEMS", None)
        if old is not None:
            os.environ["YCGE_HOST_SERIAL_BVH"] = old


@pytest.mark.parametrize("name", ["teapot", "knot:300x40", "bunny", "degenerate"])
def test_threaded_mesh_builder_gives_the_serial_tree(name):
    """csrc/bvh_build_parallel.hpp: the top of the tree cut like the serial builder cuts it, subtrees built on other threads, pieces
    concatenated in pre-order.  Node for node, leaf reference for leaf reference and fallback for fallback the serial tree
    (which the literal transcription above pins) -- also on a mesh of stacked duplicate triangles, where ranges of identical
    centroids force the Array.Sort fallback inside the tasks and in the top cuts."""
    if name == "degenerate":
        rng = np.random.default_rng(3)
        base = rng.uniform(-1, 1, (90, 3)).astype(F)
        centres = np.repeat(base, 70, axis=0)                                       # 6300 triangles, 70 copies at each of 90 places
        verts = np.concatenate([centres + F(0.01) * np.array(d, F) for d in ((1, 0, 0), (0, 1, 0), (0, 0, 1))]).astype(F)
        n = len(centres)
        faces = np.stack([np.arange(n), np.arange(n) + n, np.arange(n) + 2 * n], axis=1).astype(np.int32)
        make = lambda: api.HostScene.from_triangles("degenerate", verts, faces)
    else:
        make = lambda: api.HostScene(name)
    ser = _mesh_tree(make, serial=True)
    for task_items in (None, 16):                                                  # default task size; top cuts down to 16 items
        par = _mesh_tree(make, serial=False, task_items=task_items)
        assert par["root"] == ser["root"] and par["sort_fallbacks"] == ser["sort_fallbacks"], (name, task_items)
        for k in ("boxes", "lrsc", "leaf"):
            assert np.array_equal(par[k], ser[k]), (name, task_items, k)
    if name == "degenerate":
        assert ser["sort_fallbacks"] > 0
