"""The drop-in boundary: libycge.so loads, exports every symbol include/ycge.h declares, the ctypes mirrors match the C
structs byte for byte, the product never links or imports the oracle, and it fails loudly without a CUDA device."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT
from yetanotherconsolegameengine_b200 import api

HEADER = os.path.join(ROOT, "include", "ycge.h")
PKG = os.path.join(ROOT, "yetanotherconsolegameengine_b200")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"YCGE_API\s+[\w\s\*]+?\b(ycge_\w+)\s*\(", src)))


def test_header_symbols_all_exported():
    decl = declared_symbols()
    assert len(decl) >= 25
    assert sorted(api.ABI_SYMBOLS) == decl, "api.ABI_SYMBOLS and include/ycge.h disagree"
    lib = api.load_lib()
    for name in decl:
        assert hasattr(lib, name), f"libycge.so does not export {name}"
    out = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = set(re.findall(r"\s[TW]\s+(ycge_\w+)", out))
    assert exported == set(decl), f"exported but undeclared / declared but missing: {exported ^ set(decl)}"


def test_every_entry_point_cites_the_reference_interface():
    src = open(HEADER).read()
    assert "RaytraceEntity.cs:12-18" in src  # the seam (IConsoleRenderer)
    for cite in ("RaytraceRenderer.cs:157-267", "RaytraceRenderer.cs:110-138", "RaytraceRenderer.cs:140-148", "RaytraceRenderer.cs:150-153",
                 "Scenes/Scene.cs:66-69", "MeshBVH.cs:41-130", "VolumeGrid.cs:55-93", "RaytraceRenderer.cs:74-108"):
        assert cite in src, cite
    # every exported function line carries a citation comment or sits under a documented group
    for name in declared_symbols():
        assert re.search(r"\b" + name + r"\b", src)


C_PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include "ycge.h"
#define S(T) printf(#T " %zu\n", sizeof(T))
#define O(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
  S(ycge_material); S(ycge_object); S(ycge_light); S(ycge_bvh); S(ycge_scene); S(ycge_mesh_soa); S(ycge_volume);
  S(ycge_params); S(ycge_config); S(ycge_cell); S(ycge_stats);
  O(ycge_object, p); O(ycge_scene, lights); O(ycge_scene, bvh); O(ycge_mesh_soa, material); O(ycge_mesh_soa, bvh);
  O(ycge_volume, palette); O(ycge_params, seed_salt); O(ycge_config, params); O(ycge_cell, fg); O(ycge_cell, attr);
  O(ycge_stats, ms_trace); O(ycge_stats, kernel_launches); O(ycge_stats, rays_total);
  return 0; }
"""


def test_ctypes_structs_match_c_layout(tmp_path):
    src = tmp_path / "probe.c"
    src.write_text(C_PROBE)
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout
    c = dict(line.rsplit(" ", 1) for line in out.strip().splitlines())
    py = {"ycge_material": api.Material, "ycge_object": api.Object, "ycge_light": api.Light, "ycge_bvh": api.Bvh, "ycge_scene": api.Scene,
          "ycge_mesh_soa": api.MeshSoa, "ycge_volume": api.Volume, "ycge_params": api.Params, "ycge_config": api.Config, "ycge_stats": api.Stats}
    for k, v in c.items():
        if "." in k:
            t, f = k.split(".")
            if t == "ycge_cell":
                assert api.CELL_DTYPE.fields[f][1] == int(v), k
            else:
                assert getattr(py[t], f).offset == int(v), k
        elif k == "ycge_cell":
            assert api.CELL_DTYPE.itemsize == int(v) == 32
        else:
            assert C.sizeof(py[k]) == int(v), k
    assert int(c["ycge_material"]) == 64 and int(c["ycge_object"]) == 80


def test_default_params_are_the_reference_constants():
    p = api.default_params()  # RaytraceRenderer.cs:31-43,65; ToneMapper.cs:8-21
    assert (p.diffuse_bounces, p.max_mirror_bounces, p.max_refractions, p.atrous_iterations) == (1, 2, 2, 3)
    f32 = lambda v: C.c_float(v).value
    assert p.mirror_threshold == f32(0.9) and p.eps == f32(1e-4) and p.taa_alpha == f32(0.01)
    assert p.motion_trans_reset == f32(0.0025) and p.motion_rot_reset == f32(0.0025) and p.diffuse_sigma_deg == 25.0
    assert (p.c_phi, p.n_phi, p.z_phi, p.a_phi) == (3.0, f32(0.35), 2.0, f32(0.2))
    assert (p.tone_exposure, p.tone_gamma, p.ae_key, p.ae_speed, p.ae_min, p.ae_max) == (1.0, f32(2.2), f32(0.18), f32(0.2), f32(0.1), 1.5)
    assert (p.saturation, p.vibrance, p.auto_exposure, p.seed_salt) == (2.0, 0.0, 1, 0x9E3779B97F4A7C15)


def test_product_never_touches_the_oracle():
    # no oracle symbols linked into the product libraries ...
    for lib in (api.LIB_PATH, api.HOST_LIB_PATH):
        out = subprocess.run(["nm", "-D", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
        assert not re.search(r"\byo_\w+", out), lib
        ldd = subprocess.run(["ldd", lib], stdout=subprocess.PIPE, text=True).stdout
        assert "oracle" not in ldd
    # ... and no product source file mentions it as a dependency
    for dirpath, _, files in os.walk(PKG):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".cs")) or fn == "Makefile":
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle_binding" not in txt and "libycge_oracle" not in txt and "ycge_oracle.cpp" not in txt, os.path.join(dirpath, fn)


def test_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = api.load_lib()
    cfg = api.Config()
    cfg.fb_w, cfg.fb_h, cfg.ss = 8, 4, 1
    lib.ycge_default_params(C.byref(cfg.params))
    ctx = C.c_void_p()
    rc = lib.ycge_create(C.byref(cfg), C.byref(ctx))
    assert rc == -2 and not ctx.value  # YCGE_ERR_CUDA: no CPU path
    assert b"no CPU path" in lib.ycge_last_error(None)
    scene = api.HostScene("cornell")
    with pytest.raises(api.YcgeError):
        api.CudaRaytraceRenderer(scene, 8, 4, 1)


def test_missing_extension_raises(monkeypatch):
    monkeypatch.setattr(api, "_lib", None)
    monkeypatch.setattr(api, "LIB_PATH", os.path.join(PKG, "does_not_exist.so"))
    with pytest.raises(ImportError):
        api.load_lib()


def test_default_params_are_the_reference_constants():
    """ycge_default_params (and the oracle's) against the constants extracted from RaytraceRenderer.cs:31-43,:65,:222,:281 and
    ToneMapper.cs:8-21 by tools/extract_scene_literals.py (tests/golden/scene_literals.json)."""
    import ctypes as C
    import json
    import os
    import sys
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    from oracle_binding import load_oracle
    gold = json.load(open(os.path.join(here, "golden", "scene_literals.json")))["params"]
    for fill in (api.load_lib().ycge_default_params, load_oracle().yo_default_params):
        p = api.Params()
        fill(C.byref(p))
        for k, want in gold.items():
            got = getattr(p, k)
            if isinstance(want, bool):
                assert bool(got) == want, k
            elif isinstance(want, int):
                assert int(got) == want, k
            else:
                assert float(np.float32(got)) == want, (k, got, want)
        assert set(gold) == {f[0] for f in api.Params._fields_}, "every field of ycge_params has a reference constant behind it"


def test_csharp_binding_declares_only_exported_entry_points_with_matching_layouts():
    """host_cs/CudaRaytraceRenderer.cs cannot be compiled here (no .NET toolchain), so it is checked as text: every [DllImport]
    extern is an exported symbol of libycge.so with the same number of parameters as the header's declaration, and the struct
    sizes its static constructor asserts are the C ABI's."""
    import ctypes as C
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cs = open(os.path.join(root, "yetanotherconsolegameengine_b200", "host_cs", "CudaRaytraceRenderer.cs"), encoding="utf-8").read()
    hdr = open(os.path.join(root, "include", "ycge.h")).read()
    externs = re.findall(r"\[DllImport\(Lib\)\]\s*private static extern \w+ (ycge_\w+)\(([^)]*)\);", cs)
    assert len(externs) >= 15
    n_params = lambda text: 0 if text.strip() in ("", "void") else text.count(",") + 1
    for name, params in externs:
        assert name in api.ABI_SYMBOLS, f"{name} is not exported"
        m = re.search(r"YCGE_API[^;]*?\b" + name + r"\s*\(([^;]*?)\)\s*;", hdr, re.S)
        assert m, name
        assert n_params(params) == n_params(m.group(1)), (name, params, m.group(1))
    sizes = dict(re.findall(r"Marshal\.SizeOf<(\w+)>\(\) != (\d+)", cs))
    assert sizes == {"YMaterial": str(C.sizeof(api.Material)), "YObject": str(C.sizeof(api.Object)), "YCell": str(api.CELL_DTYPE.itemsize)}
