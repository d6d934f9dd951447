"""ref_transpile.py — rewrites the reference's own C# SOURCE TEXT into C++ mechanically (test infrastructure).

    python oracle/ref_transpile.py /root/reference/ConsoleGame oracle/_ref/ref_generated.hpp

The reference is C#/.NET and no .NET toolchain exists here, so "run the reference" is not available.  Its hot-path arithmetic,
however, is plain imperative code over structs, arrays and floats, and C# and C++ share that statement syntax almost token
for token.  This script reads the files under /root/reference at BUILD time and applies only SYNTACTIC rewrites (listed in
`rewrite`): access modifiers dropped, `new T(...)` -> `T(...)`, `MathF.Max` -> `MathF::Max`, `1f` -> `1.0f`, `ref T x` ->
`T &x`, lambdas, array declarations, operators -> friends, properties -> methods ...  No arithmetic expression is touched:
every `a * b + c`, every cast, every loop bound and every comparison of the output is the reference author's text.  The
result is compiled (-std=c++23 -ffp-contract=off) together with oracle/ref_shims.hpp (the sliver of the .NET library the
text calls) and oracle/ref_harness.cpp (C entry points) into oracle/_ref/libycge_ref.so.  Nothing generated is committed
(oracle/_ref/ is git-ignored): the reference's sources never enter this repository.

What is transpiled (and then compared with the oracle, bit for bit, by tests/test_reference_transpiled.py):
  RayTracing/Vec3.cs (whole), RayTracing/RaytraceSampler.cs (whole: blue noise, Rng, PerFrameSeed, SplitMix64,
  CosineSampleHemisphere), RayTracing/ToneMapper.cs (whole), Renderer/Chexel.cs (whole), the ANSI-256 quantiser of
  Renderer/ANSITerminalRenderer.cs and its Render with the Append* helpers (the byte stream), RayTracing/TemporalAA.cs (constructor,
  ShouldResetHistory, CommitCamera, Resize), RayTracing/MeshLoader.cs (whole) with MeshScenes.TryReadObjBoundsNormalized and the
  ground placement of AddMeshAutoGround, Renderer/Texture.cs (SampleBilinear on static images) with Renderer/RGBA32.cs, Rng.cs,
  Win32TerminalRenderer.MapAttributes, Scenes/WorldGeneration (the island generator: GenMath, TerrainNoise, RiverNetworkGlobal, BiomeMap,
  Layering, StrataMap, FloraPlacer, settings, WorldConfig, the passes of WorldManager.GenerateAndSaveWorld), and of RayTracing/RaytraceRenderer.cs: Resize, Luma, TemporalBlendWithClamp, ApplyAtrousDenoise,
  the BSDF helpers, and -- verbatim -- the tail of TryFlipAndBlit from the TAA call to the cell loop (:218-264), i.e. the
  reference's own buffer juggling, including the swap at :718 that makes the second a-trous iteration run in place.
"""
import os
import re
import sys


def strip_namespace(src):
    src = src.lstrip("﻿")
    src = re.sub(r"^\s*using [\w.= ]+;[ \t]*$", "", src, flags=re.M)
    m = re.search(r"namespace\s+[\w.]+\s*\{", src)
    if not m:
        return src
    end = src.rindex("}")
    return src[m.end():end]


def block_end(s, i):
    """index just past the brace block that starts at s[i] == '{' (char / string literals and comments skipped)"""
    depth = 0
    j = i
    while j < len(s):
        ch = s[j]
        if s.startswith("//", j):
            j = s.index("\n", j)
            continue
        if ch == "'" or ch == '"':
            k = j + 1
            while s[k] != ch:
                k += 2 if s[k] == "\\" else 1
            j = k + 1
            continue
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                return j + 1
        j += 1
    raise ValueError("unbalanced braces")


def type_body(src, name):
    m = re.search(r"\b(?:struct|class)\s+" + re.escape(name) + r"\b[^{;]*\{", src)
    if not m:
        raise KeyError(name)
    i = m.end() - 1
    return src[i + 1:block_end(src, i) - 1]


def members(body):
    """split a type body into its members: (text, name).  A member ends at ';' or at the end of its brace block."""
    out, i, n = [], 0, len(body)
    while i < n:
        while i < n and body[i] in " \t\r\n":
            i += 1
        if i >= n:
            break
        if body.startswith("//", i):
            i = body.index("\n", i) if "\n" in body[i:] else n
            continue
        if body[i] == "[":  # attribute
            i = body.index("]", i) + 1
            continue
        j = i
        while j < n and body[j] not in "{;":
            if body[j] == "=" and body[j + 1] == ">":  # expression-bodied member: ends at ';'
                j = body.index(";", j)
                break
            j += 1
        if j < n and body[j] == "{":
            k = block_end(body, j)
            # `T[] x = new T[] { ... };` : an initialiser, the member ends at the following ';'
            if re.search(r"=\s*new\b[^;{]*$", body[i:j]) or re.search(r"=\s*$", body[i:j]):
                k = body.index(";", k) + 1
            text = body[i:k]
        else:
            text = body[i:j + 1]
            k = j + 1
        head = text.split("{")[0].split("=>")[0]
        head = head.split("=")[0] if "(" not in head.split("=")[0] else head
        m = re.search(r"(operator\s*[^\s(]+|\w+)\s*(?:\(|$|;)", head.strip().rstrip(";").strip() + ";")
        names = re.findall(r"(operator\s*\S+?|\w+)\s*(?=\()", head)
        if names:
            name = names[0]
        else:
            toks = re.findall(r"\w+", head)
            name = toks[-1] if toks else ""
        out.append((text, name))
        i = k
    return out


def rewrite(t, struct_name=None, statics=()):
    """the syntactic C# -> C++ rewrites; nothing here touches an arithmetic expression"""
    t = re.sub(r"^\s*\[[A-Za-z][^\]\n]*\]\s*$", "", t, flags=re.M)                       # attributes
    t = re.sub(r"\b(public|private|internal|protected)\s+", "", t)                        # access modifiers
    t = re.sub(r"\b(readonly|sealed|unsafe)\s+", "", t)
    t = re.sub(r"\bunchecked\s*\{", "{", t)
    t = re.sub(r"\bvar\b", "auto", t)
    t = re.sub(r"\bnull\b", "nullptr", t)
    t = re.sub(r"\bulong\b", "uint64_t", t)
    t = re.sub(r"\buint\b", "uint32_t", t)
    t = re.sub(r"\blong\b", "int64_t", t)
    t = re.sub(r"\bchar\b", "char16_t", t)
    t = re.sub(r"'([^\x00-\x7f])'", r"u'\1'", t)                                          # '▀' -> u'▀'
    t = re.sub(r"(?<![\w.])(?<![eE][-+])(\d+)f\b", r"\1.0f", t)                            # 1f -> 1.0f
    t = re.sub(r"=\s*default;", "= {};", t)
    t = re.sub(r"\bthrow new \w+\(.*\);", 'throw std::runtime_error("reference exception");', t)
    # arrays
    t = re.sub(r"static (\w+)\[,\] (\w+) = new \1\[(\w+), (\w+)\]\s*\{", r"static constexpr \1 \2[\3][\4] = {", t)                  # byte[,] table
    t = re.sub(r"static (\w+)\[\] (\w+) = new \1\[\]\s*\{", r"static inline const std::vector<\1> \2 = {", t)
    t = re.sub(r"static (\w+)\[\] (\w+) = new \1\[(\w+)\];", r"static inline std::vector<\1> \2 = std::vector<\1>(\3);", t)
    t = re.sub(r"\b(\w+)\[\] (\w+) = new \1\[[^\]]*\]\s*\{", r"std::vector<\1> \2 = {", t)                                          # float[] k = new float[5] { ... }
    t = re.sub(r"\b(\w+)\[\] (\w+) = new \1\[([^\]]*)\];", r"std::vector<\1> \2(\3);", t)                                           # float[] a = new float[n];
    t = re.sub(r"\bconst (\w+) (\w+)\s*=", r"static constexpr \1 \2 =", t)                                                          # C# const members are static
    t = re.sub(r"\.Length\b", ".size()", t)
    # new
    t = re.sub(r"\bnew ((?:Fast2D<\w+>|Vec3|Material|Ray|HitRecord|Chexel|ChexelColor|RGBA32|RaytraceSampler\.Rng|PathWorkItem|PrimaryGBuffer)\s*\()", r"\1", t)  # value types (and the Fast2D handle) are constructed in place
    t = re.sub(r"\bHittable\[\] (\w+);", r"std::vector<Hittable *> \1;", t)
    t = re.sub(r"\b(\w+) = new Hittable\[(\w+)\];", r"\1.assign(\2, nullptr);", t)
    t = re.sub(r"\bfaces\[(\w+)\]\.", r"faces[\1]->", t)
    t = re.sub(r"^(\s*)((?:[\w.<>]+ \w+ = |if \().*)\bout (Vec3|float|int|bool) (\w+)\)", r"\1\3 \4;\n\1\2\4)", t, flags=re.M)   # inline `out T x` at a call site: declared just before the statement
    # parameters passed by reference
    t = re.sub(r"\b(?:ref|out) ([\w.<>]+) (\w+)(?=\s*[,)])", r"\1 &\2", t)
    t = re.sub(r"(?<=[(,\s])(?:ref|out) (?=\w+\s*[,)])", "", t)                                                                    # call sites
    # lambdas
    t = re.sub(r"(?<![\w)])(\w+)\s*=>\s*\{", r"[&](int \1) {", t)
    # static members of known types
    for s in ("MathF", "Math", "RaytraceSampler", "Vec3", "ChexelColor", "ToneMapper") + tuple(statics):
        t = re.sub(r"\b" + s + r"\.(?=[A-Z])", s + "::", t)
    t = re.sub(r"\bfloat\.(?=[A-Z])", "Single::", t)
    t = re.sub(r"\bint\.(?=[A-Z])", "Int32::", t)
    t = re.sub(r"\bVec3::Zero\b(?!\()", "Vec3::Zero()", t)
    t = re.sub(r"\bBlueNoise8x8\[(\w+), (\w+)\]", r"BlueNoise8x8[\1][\2]", t)
    t = re.sub(r"\bthis\.", "this->", t)
    t = re.sub(r"\breturn this;", "return *this;", t)
    # members
    if struct_name:
        # conversions: from the struct -> conversion operator; to the struct -> the converting constructor already exists
        t = re.sub(r"static implicit operator (\w+)\(" + struct_name + r" (\w+)\)\s*\{", r"operator \1() const { const " + struct_name + r" &\2 = *this;", t)
        t = re.sub(r"static implicit operator " + struct_name + r"\([^)]*\)\s*\{[^}]*\}", "", t)
    t = re.sub(r"^(\s*)static (?!constexpr|inline)(\w+) (\w+) = ", r"\1static inline \2 \3 = ", t, flags=re.M)
    t = re.sub(r"static (\w+) (\w+) => ([^;]+);", r"static \1 \2() { return \3; }", t)                                              # static T Zero => expr;
    t = re.sub(r"\bstatic (\w+) operator\s*([^\s(]+)\s*\(", r"friend \1 operator\2(", t)
    t = re.sub(r"\babstract ([^;{]*\));", r"virtual \1 = 0;", t)
    t = re.sub(r"Func<(\w+), (\w+), (\w+), (\w+)>", r"std::function<\4(\1, \2, \3)>", t)
    t = re.sub(r"Func<(\w+), (\w+), (\w+)>", r"std::function<\3(\1, \2)>", t)
    t = re.sub(r"\bfloat\.IsInfinity\(", "Single::IsInfinity(", t)
    t = re.sub(r"\(pos, n, u\) =>\s*\{", "[=](Vec3 pos, Vec3 n, float u) {", t)
    t = re.sub(r"\(pos, n, u\) => ([^;]+);", r"[=](Vec3 pos, Vec3 n, float u) { return \1; };", t)
    t = re.sub(r"\bTexture (\w+)", r"Texture *\1", t)
    t = re.sub(r"\babstract\s+", "virtual ", t)
    t = re.sub(r"\boverride\s+", "", t)
    return t


def emit_struct(src, name, statics=()):
    body = rewrite(type_body(src, name), name, statics)
    body = re.sub(r"^(\s*)((?:float|int|bool|double|Vec3|ChexelColor|ConsoleColor|char16_t|uint64_t|uint32_t) \w+);", r"\1\2 = {};", body, flags=re.M)  # C# zero-initialises fields
    return "struct %s {\n    %s() = default;\n%s\n};\n" % (name, name, body)



def bvh_class(src, name, elem, var_names):
    """MeshBVH.cs / BVH.cs whole: constructor (Item list), BuildRecursive (binned SAH, partition, Array.Sort fallback), Hit, BoxHitFast (+ TriHit).
    `elem` = the element class (Triangle / Hittable: reference types -> pointers), `var_names` = the identifiers that hold one."""
    mb = type_body(src, name)
    mb = re.sub(r"^.*Vector128.*$", "", mb, flags=re.M)
    while True:
        m = re.search(r"if \(Sse\w*\.IsSupported[^)]*\)", mb)
        if not m:
            break
        b0 = mb.index("{", m.end())
        e0 = block_end(mb, b0)
        m2 = re.match(r"\s*else\s*\{", mb[e0:])  # `if (Sse...) {...} else {scalar}` keeps the scalar block
        if m2:
            b1 = e0 + m2.end() - 1
            e1 = block_end(mb, b1)
            mb = mb[:m.start()] + mb[b1 + 1:e1 - 1] + mb[e1:]
        else:
            mb = mb[:m.start()] + mb[e0:]
    E = elem
    mb = re.sub(r"^\s*(?:private |public )?delegate [^;]*;", "", mb, flags=re.M)
    mb = re.sub(r"\bHitFunc\[\] (\w+);", r"std::vector<std::function<bool(Ray, float, float, HitRecord &, float, float)>> \1;", mb)
    mb = re.sub(r"= Array\.Empty<HitFunc>\(\)", "= {}", mb)
    mb = re.sub(r"(\w+) = new HitFunc\[(\w+)\];", r"\1.resize(\2);", mb)
    mb = re.sub(r"(\w+)\[(\w+)\] = (\w+)\[\2\]\.Hit;", r"\1[\2] = [p_ = \3[\2]](Ray r_, float a_, float b_, HitRecord &h_, float u_, float v_) { return p_->Hit(r_, a_, b_, h_, u_, v_); };", mb)  # method group -> delegate
    mb = re.sub(r"IEnumerable<" + E + r"> (\w+)", r"const std::vector<" + E + r" *> &\1", mb)
    mb = re.sub(r"\bobjects\.Count\(\)", "(int)objects.size()", mb)
    mb = re.sub(r"foreach \(" + E + r" (\w+) in (\w+)\)", r"for (" + E + r" *\1 : \2)", mb)
    mb = re.sub(r"\bList<" + E + r">", "List<" + E + " *>", mb)
    vn = "|".join(var_names)
    mb = re.sub(r"\b" + E + r" (" + vn + r")\b(?! :)", E + r" *\1", mb)
    mb = re.sub(r"\b(" + vn + r")\.(?=[A-Z])", r"\1->", mb)
    mb = re.sub(r"new (List<[\w *]+>)\(", r"\1(", mb)
    mb = re.sub(r"([(,]\s*)(List<[\w *]+>) (\w+)(?=[,)])", r"\1\2 &\3", mb)           # List<T> parameters: reference types
    mb = re.sub(r"([(,]\s*)(\w+)\[\] (\w+)(?=[,)])", r"\1std::vector<\2> &\3", mb)       # T[] parameters likewise
    mb = re.sub(r"\b(nodes|leafIndices|tris|items|objs)\.Count\b(?!\()", r"\1.Count()", mb)   # List<T>.Count is a property
    mb = re.sub(r"\bin Ray (\w+)", r"const Ray &\1", mb)
    mb = re.sub(r"^(\s*)(?:private |public )?(float|int|Material)\[\] ([\w, ]+);", r"\1std::vector<\2> \3;", mb, flags=re.M)   # float[] ax, ay, az;
    mb = re.sub(r"= Array\.Empty<(\w+)>\(\)", r"= std::vector<\1>()", mb)
    mb = re.sub(r"\bnew (float|int|Material)\[([^\]]+)\](?!\s*\{)", r"std::vector<\1>(\2)", mb)
    mb = re.sub(r"^(\s*)(float|int)\[\] ", r"\1std::vector<\2> ", mb, flags=re.M)
    mb = re.sub(r"(\w+)\[\] (\w+) = (\w+)\.ToArray\(\);", r"std::vector<\1> \2 = \3.ToArray();", mb)
    mb = re.sub(r"Span<int> (\w+) = stackalloc int\[(\d+)\];", r"int \1[\2];", mb)
    mb = re.sub(r"\(a, b\) => (a\.\w+)\.CompareTo\((b\.\w+)\)", r"[](const Item &a, const Item &b) { return SingleCompareTo(\1, \2); }", mb)
    mb = re.sub(r"Array\.Sort\((\w+), (\w+), (\w+), Comparer<Item>\.Create\((\w+)\)\);", r"Array::Sort(\1, \2, \3, \4);", mb)
    mb = re.sub(r"\bnew (NodeTmp|Item)\(\)", r"\1()", mb)
    mb = re.sub(r"\.Add\(default\)", ".Add({})", mb)
    mb = re.sub(r"\bout (\w+\.\w+)", r"\1", mb)                                        # `out it.MinX` at a call site
    mb = re.sub(r"\bref (\w+\[[^\]]+\])", r"\1", mb)                                   # `ref lminx[b]` at a call site
    mb = re.sub(r"(struct (?:NodeTmp|Item)\s*\{[^}]*\})", r"\1;", mb)
    mb = re.sub(r"\|\s*MethodImplOptions\.\w+", "", mb)
    mb = rewrite(mb)
    nested = re.findall(r"struct (?:NodeTmp|Item)\s*\{[^}]*\};", mb)  # hoisted: C++ wants them declared before the signatures that name them
    for nsrc in nested:
        mb = mb.replace(nsrc, "")
    return "struct %s : Hittable {\n%s\n%s\n};\n" % (name, "\n".join(nested), mb)


def main(ref, out_path):
    rd = lambda p: strip_namespace(open(os.path.join(ref, p), encoding="utf-8-sig").read())
    out = ["// GENERATED by oracle/ref_transpile.py from the reference's C# sources under " + ref + " -- do not edit, do not commit.",
           "#pragma once", '#include "../ref_shims.hpp"', "namespace refcs {", ""]
    v3 = emit_struct(rd("RayTracing/Vec3.cs"), "Vec3")
    # two things the C# COMPILER supplies: overload resolution picks the float constructor for int arguments (`new Vec3(1, 1, 1)`), and
    # `a += b` is `a = a + b` for a type that defines operator +
    v3 = v3.replace("    Vec3() = default;", "    Vec3() = default;\n    Vec3(int x, int y, int z) : Vec3((float)x, (float)y, (float)z) {}\n    Vec3 &operator+=(Vec3 b) { *this = *this + b; return *this; }")
    out.append(v3)
    chex = rd("Renderer/Chexel.cs")
    out.append(emit_struct(chex, "ChexelColor"))
    out.append(emit_struct(chex, "Chexel"))
    out.append("struct Framebuffer { Fast2D<Chexel> cells; int Width, Height, ViewportX = 0, ViewportY = 0; Framebuffer(int w, int h) : cells(w, h), Width(w), Height(h) {} "
               "void SetChexel(int x, int y, Chexel c) { cells[x, y] = c; } Chexel GetChexel(int x, int y) { return cells[x, y]; } };\n")
    samp = rd("RayTracing/RaytraceSampler.cs")
    body = rewrite(type_body(samp, "RaytraceSampler"), None)
    body = re.sub(r"^(\s*)(uint64_t state);", r"\1\2 = 0;", body, flags=re.M)
    body = body.replace("struct Rng\n", "struct Rng\n").replace("};", "};")
    body = re.sub(r"(struct Rng\s*\{)", r"\1\n            Rng() = default;", body)
    body = re.sub(r"(struct Rng\s*\{.*?\n        \})", r"\1;", body, flags=re.S)  # nested struct needs its ';'
    out.append("struct RaytraceSampler {\n%s\n};\n" % body)
    tm = rd("RayTracing/ToneMapper.cs")
    tb = rewrite(type_body(tm, "ToneMapper"), None)
    tb = re.sub(r"float EffectiveExposure\s*\{\s*get \{ return effectiveExposure; \}\s*\}", "", tb)
    tb = re.sub(r"(Fast2D<bool> optionalSkyMask) = nullptr", r"\1 = Fast2D<bool>()", tb)
    tb = re.sub(r"if \(threadpool == nullptr\) \{[^}]*\}", "", tb)   # a FixedThreadFor VALUE is never null
    tb = re.sub(r"if \(threadpool == nullptr\) return [^;]*;", "", tb)
    tb = re.sub(r"FixedThreadFor threadpool\b", "FixedThreadFor &threadpool", tb)
    out.append("struct ToneMapper {\n%s\n};\n" % tb)
    # ---- Rng.cs (ConsoleRayTracing.Rng: dead code in the engine, named by the north star): whole, in a namespace of its own (RaytraceSampler has a Rng too)
    out.append("namespace rngcs {\n%s}\n" % emit_struct(rd("Rng.cs"), "Rng"))
    # ---- Win32TerminalRenderer.MapAttributes (:109-112): the console attribute word of a cell (fg | bg << 4)
    w32 = [t for t, n in members(type_body(rd("Renderer/Win32TerminalRenderer.cs"), "Win32TerminalRenderer")) if n == "MapAttributes"]
    out.append("struct Win32Ref {\n%s\n};\n" % rewrite("\n".join(w32), None))
    # ---- TemporalAA.cs: what TryFlipAndBlit asks it (:171 ShouldResetHistory, :266 CommitCamera, :128 Resize) and its constructor; the
    # blend itself lives in RaytraceRenderer.TemporalBlendWithClamp (BlendIntoHistory is not on the path)
    taa_members = [t for t, n in members(type_body(rd("RayTracing/TemporalAA.cs"), "TemporalAA")) if n not in ("BlendIntoHistory", "GetHistory", "History", "HistoryValid")]
    taa = rewrite("\n".join(taa_members), None)
    taa = re.sub(r"^(\s*)((?:int|bool|float) \w+);", r"\1\2 = {};", taa, flags=re.M)
    out.append("struct TemporalAA {\n%s\n};\n" % taa)
    ansi = rd("Renderer/ANSITerminalRenderer.cs")
    ab = type_body(ansi, "ANSITerminalRenderer")
    want = {"s_cubeSrgb", "s_cubeLinear", "s_graySrgb", "s_grayLinear", "ChexelToAnsi256", "ToCubeLevelSrgb", "LinearToSrgb8", "Dist2Srgb"}
    sel = [t for t, n in members(ab) if n in want]
    out.append("struct AnsiRef {\n%s\n};\n" % rewrite("\n".join(sel), None))
    # ---- ANSITerminalRenderer.Render (:86-153) with everything it calls (GetChexelForPoint, Append*, EnsureCapacity): the byte stream the
    # terminal receives.  Not taken: the constructor (console mode, cursor), Flush (WriteFile / stdout: the harness keeps the bytes instead).
    want = {"frameBuffers", "consoleWidth", "consoleHeight", "defaultFg", "defaultBg", "onResize", "zeroSeq", "outBuf", "outLen", "GetChexelForPoint", "Render",
            "EnsureCapacity", "AppendAscii", "AppendBytes", "AppendInt", "AppendCharUtf8"}
    rt = "\n".join(t for t, n in members(ab) if n in want and "ITerminalRenderer." not in t)          # not the explicit interface properties
    rt = re.sub(r"\(byte\[\] (\w+)\)", r"(std::vector<byte> &\1)", rt)                                   # an array parameter: a reference
    rt = re.sub(r"List<Framebuffer> frameBuffers;", "List<Framebuffer *> frameBuffers;", rt)             # Framebuffer is a class: a list of references
    rt = re.sub(r"Framebuffer fb = frameBuffers\[i\];", "Framebuffer &fb = *frameBuffers[i];", rt)
    rt = re.sub(r"\bframeBuffers\.Count\b(?!\()", "frameBuffers.Count()", rt)                           # List<T>.Count is a property
    rt = re.sub(r"Action<int, int> onResize;", "std::function<void(int, int)> onResize;", rt)
    rt = re.sub(r"onResize\?\.Invoke\(([^;]*)\);", r"if (onResize) onResize(\1);", rt)                  # null-conditional call
    rt = re.sub(r"Array\.Resize\(ref (\w+), (\w+)\);", r"Array::Resize(\1, \2);", rt)
    rt = re.sub(r"\bstring (\w+)\)", r"const std::string &\1)", rt)
    rt = rewrite(rt, None, statics=("Console", "Buffer"))
    rt = re.sub(r"std::vector<byte> outBuf\(([^;]*)\);", r"std::vector<byte> outBuf = std::vector<byte>(\1);", rt)  # a member initialiser
    rt = re.sub(r"^(\s*)((?:int|ConsoleColor) \w+);", r"\1\2 = {};", rt, flags=re.M)                    # C# zero-initialises fields
    out.append("struct AnsiRenderRef : AnsiRef {\n    std::vector<byte> flushed;\n    void Flush() { flushed.insert(flushed.end(), outBuf.begin(), outBuf.begin() + outLen); }\n%s\n};\n" % rt)
    # ---- the analytic primitives and what their Hit needs
    # ---- Renderer/RGBA32.cs (the packed pixel: the int constructor and toVec3) and Renderer/Texture.cs: the static-image half of SampleBilinear
    # (:108-163 without the live-frame branch, removed mechanically), Lerp, Frac; pixels / width / height are filled by the harness
    rg = type_body(rd("Renderer/RGBA32.cs"), "RGBA32")
    rsel = [t for t, n in members(rg) if n in ("a", "g", "b", "r", "toVec3") or (n == "RGBA32" and "(int value)" in t)]
    out.append("struct RGBA32 {\n%s\n};\n" % rewrite("\n".join(rsel), None))
    tx = type_body(rd("Renderer/Texture.cs"), "Texture")
    tsel = "\n".join(t for t, n in members(tx) if n in ("pixels", "width", "height", "SampleBilinear", "Lerp", "Frac"))
    m_dyn = re.search(r"if \(isDynamic && dynamicReader != null\)", tsel)
    b0 = tsel.index("{", m_dyn.end())
    tsel = tsel[:m_dyn.start()] + tsel[block_end(tsel, b0):]
    tsel = re.sub(r"\bint\[\] pixels;", "std::vector<int> pixels;", tsel)
    tsel = re.sub(r"if \(pixels == null\)", "if (pixels.empty())", tsel)
    tsel = re.sub(r"static float Frac\(float x\) => ([^;]+);", r"static float Frac(float x) { return \1; }", tsel)
    tsel = rewrite(tsel, None)
    tsel = re.sub(r"^(\s*)(int \w+);", r"\1\2 = {};", tsel, flags=re.M)
    out.append("struct Texture {\n%s\n};\n" % tsel)
    out.append(emit_struct(rd("RayTracing/Ray.cs"), "Ray"))
    out.append(emit_struct(rd("RayTracing/Material.cs"), "Material"))
    out.append(emit_struct(rd("RayTracing/HitRecord.cs"), "HitRecord"))
    hb = rewrite(type_body(rd("RayTracing/Objects/Hittable.cs"), "Hittable"))
    out.append("struct Hittable {\n    virtual ~Hittable() {}\n%s\n};\n" % hb)
    for path, names in (("RayTracing/Objects/Surfaces.cs", ("Plane", "Disk", "XYRect", "XZRect", "YZRect")), ("RayTracing/Objects/BoundedObjects.cs", ("Sphere", "Box", "CylinderY")),
                        ("RayTracing/Objects/Triangle.cs", ("Triangle",))):
        src = rd(path)
        for nm in names:
            body = type_body(src, nm)
            if nm == "Triangle":  # the scalar path is the specified one: drop the SSE fields and the SSE branch, mechanically
                body = re.sub(r"^.*Vector128.*$", "", body, flags=re.M)
                while True:
                    m = re.search(r"if \(Sse\.IsSupported[^)]*\)", body)
                    if not m:
                        break
                    b0 = body.index("{", m.end())
                    body = body[:m.start()] + body[block_end(body, b0):]
                body = re.sub(r"\|\s*MethodImplOptions\.AggressiveOptimization", "", body)
            out.append("struct %s : Hittable {\n%s\n};\n" % (nm, rewrite(body)))
    out.append(bvh_class(rd("RayTracing/Objects/MeshBVH.cs"), "MeshBVH", "Triangle", ("t", "tr")))
    out.append(bvh_class(open(os.path.join(ref, "RayTracing/Objects/BVH.cs"), encoding="utf-8-sig").read(), "BVH", "Hittable", ("h",)))
    # ---- VolumeGrid.cs: everything but the constructor (it takes a C# tuple array; the harness fills the same fields from the flat
    # description, whose mat / meta arrays already are in the reference's bricked-Morton order), the GC handles and disposal
    vsrc = type_body(rd("RayTracing/Objects/VolumeGrid.cs"), "VolumeGrid")
    skip = {"VolumeGrid", "mat", "meta", "matHandle", "metaHandle", "BoundsMin", "BoundsMax", "Dispose", "disposed"}
    vtxt = "\n".join(t for t, n in members(vsrc) if n not in skip and not t.lstrip().startswith("~"))
    vtxt = re.sub(r"^(\s*)(?:private |public )?(?:readonly )?(int|float|bool) (\w+);", r"\1\2 \3 = {};", vtxt, flags=re.M)
    vtxt = re.sub(r"\bint\* (\w+);", r"int *\1 = nullptr;", vtxt)
    out.append("struct VolumeGrid : Hittable {\n%s\n};\n" % rewrite(vtxt))
    msrc = type_body(rd("RayTracing/Mesh.cs"), "Mesh")
    msel = [t for t, n in members(msrc) if n != "FromObj"]
    mtxt = "\n".join(msel)
    mtxt = re.sub(r"\bHittable bvh;", "Hittable *bvh = nullptr;", mtxt)
    mtxt = re.sub(r"List<Triangle> (\w+)", r"const std::vector<Triangle *> &\1", mtxt)
    mtxt = re.sub(r"\bbvh\.(?=[A-Z])", "bvh->", mtxt)
    mtxt = rewrite(mtxt).replace("bvh = new MeshBVH(", "bvh = new MeshBVH(")
    out.append("struct Mesh : Hittable {\n%s\n};\n" % mtxt)
    # ---- MeshLoader.cs (FromObj with its OBJ reader, ParseIndex, NormalizeAllUsedVertices) and, of Scenes/MeshScenes.cs, TryReadObjBoundsNormalized
    # (:186-331) with the three lines of AddMeshAutoGround that place the mesh on the ground (:175-182): what decides the triangles a mesh scene uploads
    def mesh_text(t):
        t = re.sub(r"using \(var sr = new StreamReader\(path\)\)", "if (StreamReader sr(path); true)", t)                     # using (...) { } : a scope
        t = re.sub(r"(\w+)\.Split\(\(char\[\]\)null, StringSplitOptions\.RemoveEmptyEntries\)", r"\1.SplitWhitespace()", t)
        t = re.sub(r"\bstring\[\] (\w+) =", r"std::vector<String> \1 =", t)
        t = re.sub(r"\bfloat\.Parse\(", "SingleParse(", t)
        t = re.sub(r"\bint\.Parse\(", "Int32Parse(", t)
        t = re.sub(r"\bstring\.(?=[A-Z])", "String::", t)
        t = re.sub(r"\bstring\b", "String", t)
        t = re.sub(r"List<\(int a, int b, int c\)>|List<\(int, int, int\)>", "List<Face3>", t)                               # the value tuple -> a struct with the same field names
        t = re.sub(r"\.Add\(\((vIdx\[0\], vIdx\[i - 1\], vIdx\[i\])\)\);", r".Add(Face3{\1});", t)
        t = re.sub(r"= \((remap\[f\.a\], remap\[f\.b\], remap\[f\.c\])\);", r"= Face3{\1};", t)
        t = re.sub(r"\bvar \((\w+), (\w+), (\w+)\) =", r"auto [\1, \2, \3] =", t)                                           # deconstruction -> structured binding
        t = re.sub(r"Vec3\? translate = null", "std::optional<Vec3> translate = std::nullopt", t)
        t = re.sub(r"translate \?\? (new Vec3\([^)]*\))", r"translate.value_or(\1)", t)
        t = re.sub(r"\bref Vec3\[\] (\w+)", r"std::vector<Vec3> &\1", t)
        t = re.sub(r"\bVec3\[\] (\w+) = (\w+)\.ToArray\(\);", r"std::vector<Vec3> \1 = \2.ToArray();", t)
        t = re.sub(r"Dictionary<int, List<int>>", "Dictionary<int, RList<int>>", t)                                          # List<int> instances shared between the dictionary and a local
        t = re.sub(r"if \(!compToFaces\.TryGetValue\(r, out var list\)\) \{ list = new List<int>\(64\);", "RList<int> list;\n                if (!compToFaces.TryGetValue(r, list)) { list = RList<int>(64);", t)
        t = re.sub(r"new ((?:List|HashSet|Dictionary|RList)<[^()]*>)\(", r"\1(", t)                                          # collections are constructed in place
        t = re.sub(r"List<Triangle> (\w+) = List<Triangle>\(", r"List<Triangle *> \1 = List<Triangle *>(", t)
        t = re.sub(r"\b(positions|faces|fi|keptFaces|usedVerts)\.Count\b(?!\()", r"\1.Count()", t)                          # Count is a property
        t = re.sub(r"\bkv\.Value\.Count\b", "kv.Value.Count()", t)
        t = re.sub(r"foreach \((?:int|var) (\w+) in (\w+)\)", r"for (auto &\1 : \2)", t)
        t = re.sub(r"^(\s*)int Find\(int x\) \{(.*)\}\s*$", r"\1auto Find = [&](int x) -> int {\2};", t, flags=re.M)                # local functions -> lambdas
        t = re.sub(r"^(\s*)void Union\(int x, int y\) \{(.*)\}\s*$", r"\1auto Union = [&](int x, int y) -> void {\2};", t, flags=re.M)
        t = re.sub(r"\bstatic Mesh FromObj\(", "static List<Triangle *> FromObj(", t)                                        # the triangle list itself: Mesh (-> MeshBVH) is built from it elsewhere
        t = re.sub(r"return new Mesh\(tris, mn, mx\);", "return tris;", t)
        t = rewrite(t, None, statics=("CultureInfo", "File"))
        return re.sub(r"\bfaces\[(\w+)\]->", r"faces[\1].", t)   # `faces` is a list of value tuples here (rewrite() knows the name from Box: a list of references)
    ml = mesh_text(type_body(rd("RayTracing/MeshLoader.cs"), "MeshLoader"))
    out.append("struct MeshLoaderRef {\n%s\n};\n" % ml)
    msb = type_body(rd("RayTracing/Scenes/MeshScenes.cs"), "MeshScenes")
    msel = [t for t, n in members(msb) if n in ("TryReadObjBoundsNormalized", "ParseIndex")]
    ag = [t for t, n in members(msb) if n == "AddMeshAutoGround"][0]
    a3, b3 = ag.index("Vec3 mnN, mxN;"), ag.index("s.Objects.Add(")
    msel.append("static Vec3 AutoGroundTranslate(string objPath, float scale, Vec3 targetPos)\n{\n" + ag[a3:b3] + "\n    return translate;\n}\n")
    out.append("struct MeshScenesRef {\n%s\n};\n" % mesh_text("\n".join(msel)))
    # ---- Scenes/WorldGeneration: the island generator behind BuildMinecraftLike (VolumeScenes.cs:569) -- GenMath, TerrainNoise, BiomeMap, Layering,
    # StrataMap, RiverNetworkGlobal, FloraPlacer, the settings, WorldConfig, and the three passes of WorldManager.GenerateAndSaveWorld (:510-606,
    # up to the file writer): every (block id, meta) of the world
    wg = "RayTracing/Scenes/WorldGeneration/"
    def nested_static_classes(t):
        while True:
            m_ = re.search(r"(?:public |internal )?static class (\w+)\s*\{", t)
            if not m_:
                return t
            b0_ = m_.end() - 1
            e0_ = block_end(t, b0_)
            t = t[:m_.start()] + "struct " + m_.group(1) + " {" + t[b0_ + 1:e0_] + ";" + t[e0_:]
    def gen_text(t):
        t = re.sub(r"Console\.WriteLine\([^;]*\);", "(void)0;", t)                                                                  # progress messages
        t = re.sub(r"\(int x, int z, int h\)\[\]|var order = new \(int x, int z, int h\)\[([^\]]*)\];", lambda m_: "std::vector<OrderItem> order(%s);" % m_.group(1) if m_.group(1) else "std::vector<OrderItem>", t)
        t = re.sub(r"order\[k\+\+\] = \((x, z, ground\[x, z\])\);", r"order[k++] = OrderItem{\1};", t)
        t = re.sub(r"Array\.Sort\(order, \(a, b\) => a\.h\.CompareTo\(b\.h\)\);", "Array::Sort(order, [](const OrderItem &a, const OrderItem &b) { return Int32CompareTo(a.h, b.h); });", t)
        t = re.sub(r"new \(int, int\)\[(\w+), (\w+), (\w+)\]", r"Array3<Cell2>(\1, \2, \3)", t)
        t = re.sub(r"\(int, int\)\[,,\] (\w+)", r"Array3<Cell2> \1", t)
        t = re.sub(r"\(int, int\) (\w+);", r"Cell2 \1;", t)
        t = re.sub(r"\((WorldGenSettings\.Blocks\.\w+|top|sub), (0|1|meta)\)", r"Cell2{\1, \2}", t)                                    # (id, meta) tuple literals
        t = re.sub(r"new (\w+)\[(\w+), (\w+)\]", r"Array2<\1>(\2, \3)", t)
        t = re.sub(r"\bout (\w+)\[,\] (\w+)", r"Array2<\1> &\2", t)
        t = re.sub(r"\b(\w+)\[,\] (\w+)", r"Array2<\1> \2", t)
        t = re.sub(r"RiverNetworkGlobal\.Compute\(nx, nz, config, ground, out var carveDepth, out var riverWater\);",
                   "Array2<float> carveDepth; Array2<int> riverWater;\n            RiverNetworkGlobal::Compute(nx, nz, config, ground, carveDepth, riverWater);", t)
        t = re.sub(r"static float\[\] (\w+)\(", r"static std::vector<float> \1(", t)
        t = re.sub(r"return new float\[\] \{", "return std::vector<float>{", t)
        t = re.sub(r"\(float\[\] g,", "(const std::vector<float> &g,", t)
        t = re.sub(r"static (\w+) (\w+)\(([^)]*)\) => ([^;]+);", r"static \1 \2(\3) { return \4; }", t)                                   # expression-bodied methods
        t = re.sub(r"(?:internal|public) enum (\w+)", r"enum \1", t)
        t = re.sub(r"\bWorldGenSettings\.(\w+)\.(\w+)", r"WorldGenSettings::\1::\2", t)
        t = rewrite(t, None, statics=("GenMath", "IslandSettings", "TerrainNoise", "BiomeMap", "Layering", "StrataMap", "RiverNetworkGlobal", "FloraPlacer", "Biome"))
        return t
    def static_class(path, name):
        body = nested_static_classes(type_body(rd(path), name))
        return "struct %s {\n%s\n};\n" % (name, gen_text(body))
    out.append(static_class(wg + "WorldGenSettings.cs", "WorldGenSettings"))
    out.append(static_class(wg + "IslandSettings.cs", "IslandSettings"))
    bsrc = rd(wg + "Biome.cs")
    out.append(gen_text(bsrc[bsrc.index("internal enum"):bsrc.rindex("}") + 1]) + ";\n")
    wc = gen_text(type_body(rd(wg + "WorldConfig.cs"), "WorldConfig"))
    wc = re.sub(r"^(\s*)((?:int|Vec3) \w+);", r"\1\2 = {};", wc, flags=re.M)
    out.append("struct WorldConfig {\n%s\n};\n" % wc)
    for fn, nm in (("GenMath.cs", "GenMath"), ("StrataMap.cs", "StrataMap"), ("BiomeMap.cs", "BiomeMap"), ("Layering.cs", "Layering"), ("TerrainNoise.cs", "TerrainNoise"),
                   ("RiverNetworkGlobal.cs", "RiverNetworkGlobal"), ("FloraPlacer.cs", "FloraPlacer")):
        out.append(static_class(wg + fn, nm))
    gsw = [t for t, n in members(type_body(rd(wg + "WorldManager.cs"), "WorldManager")) if n == "GenerateAndSaveWorld"][0]
    a4, b4 = gsw.index("int nx = config.ChunksX"), gsw.index("// --- Write file ---")
    # ... and the VG01 writer that follows them (:607-631), verbatim but for the directory creation and the stream constructor
    wtxt = gsw[b4:gsw.rindex("}")]
    wtxt = re.sub(r"string dir = Path\.GetDirectoryName\(filename\);", "", wtxt)
    wtxt = re.sub(r"if \(!string\.IsNullOrEmpty\(dir\) && !Directory\.Exists\(dir\)\) Directory\.CreateDirectory\(dir\);", "", wtxt)
    wtxt = re.sub(r"using \(var bw = new BinaryWriter\(File\.Open\(filename, FileMode\.Create, FileAccess\.Write, FileShare\.None\)\)\)", "if (BinaryWriter bw(filename); true)", wtxt)
    out.append("struct WorldGenRef {\nstatic Array3<Cell2> GenerateCells(WorldConfig config, const String &filename = String())\n{\n" + gen_text(gsw[a4:b4])
               + "\n    if (filename != nullptr) {\n" + gen_text(wtxt) + "\n    }\n    return worldCells;\n}\n};\n")
    out.append(emit_struct(rd("RayTracing/Objects/PointLight.cs"), "PointLight"))
    out.append(emit_struct(rd("RayTracing/Objects/AmbientLight.cs"), "AmbientLight"))
    ssrc = type_body(rd("RayTracing/Scenes/Scene.cs"), "Scene")
    want = {"Objects", "Lights", "BackgroundTop", "BackgroundBottom", "Ambient", "bvh", "RebuildBVH", "Hit", "Occluded"}
    stxt = "\n".join(t for t, n in members(ssrc) if n in want)
    stxt = re.sub(r"List<Hittable> Objects = new List<Hittable>\(\);", "List<Hittable *> Objects;", stxt)
    stxt = re.sub(r"List<PointLight> Lights = new List<PointLight>\(\);", "List<PointLight> Lights;", stxt)
    stxt = re.sub(r"\bBVH bvh;", "BVH *bvh = nullptr;", stxt)
    stxt = re.sub(r"new BVH\(Objects\)", "new BVH(Objects.v)", stxt)
    stxt = re.sub(r"\bbvh\.(?=[A-Z])", "bvh->", stxt)
    stxt = re.sub(r"new AmbientLight\(", "AmbientLight(", stxt)
    out.append("struct SceneRef {\n    bool IsVolumeScene = false; // `scene is VolumeScene`\n    virtual ~SceneRef() {}\n%s\n};\n" % rewrite(stxt))
    sc = rd("RayTracing/Scenes/Scenes.cs")
    sel = [t for t, n in members(type_body(sc, "Scenes")) if n in ("Solid", "Emissive", "Checker")]
    out.append("struct ScenesRef {\n%s\n};\n" % rewrite("\n".join(sel)))
    rr = rd("RayTracing/RaytraceRenderer.cs")
    rb = type_body(rr, "RaytraceRenderer")
    mem = members(rb)
    fields = {"taaAlpha", "taaHistory", "taaHistoryValid", "prevNormal", "prevDepth", "prevSky", "spatialA", "spatialB", "gAlbedo", "gNormal", "gDepth", "skyMask",
              "toneMapper", "threadpool", "procCount", "ss", "fbW", "fbH", "Pi", "InvPi", "DiffuseSigmaDeg", "Eps", "MirrorThreshold"}
    fields |= {"hiW", "hiH", "fovDeg", "rays", "currentHdr", "pixelPool", "frameCounter", "frameBuffer", "scene"}
    fields |= {"MotionTransReset", "MotionRotReset", "taa"}
    fields |= {"DiffuseBounces", "IndirectSamples", "MaxMirrorBounces", "MaxRefractions", "SeedSalt", "MaxLuminance", "PrimaryGBuffer", "PathWorkItem"}
    funcs = {"Resize", "Luma", "TemporalBlendWithClamp", "ApplyAtrousDenoise", "OrenNayarBRDF", "FresnelSchlick", "Refract", "Reflect", "Lerp", "SampleAlbedo", "ForwardFromYawPitch",
             "MakeJitteredRay", "TraceFull", "ComputeTransmittanceToLight", "CosineSampleHemisphere"}
    sel = [t for t, n in mem if n in fields] + [t for t, n in mem if n in funcs]
    flip = [t for t, n in mem if n == "TryFlipAndBlit"][0]
    a, b = flip.index("Fast2D<Vec3> blendedHdr = TemporalBlendWithClamp("), flip.index("taa.CommitCamera")
    tail = flip[a:b]
    tail = re.sub(r"(?<=[(,\s])([A-Za-z_]\w*):\s+(?=[\w\d.\-])", "", tail)  # named arguments (all in declaration order) -> positional
    sel.append("void PostTail(Fast2D<Vec3> currentHdr, bool resetHistory, Fast2D<Chexel> target, Framebuffer &fb)\n{\n" + tail + "\n    lastDenoised = denoisedHdr;\n}\n")
    # ... and its head, verbatim as well (:159, :175-216): frame counter, per-frame jitter rotations, ray generation, the per-pixel trace loop
    a2, b2 = flip.index("long frame = Interlocked.Increment(ref frameCounter);"), flip.index("Fast2D<Vec3> blendedHdr = TemporalBlendWithClamp(")
    head = re.search(r"float aspect = [^;]*;", flip).group(0) + "\n" + flip[a2:b2]
    head = head.replace("Interlocked.Increment(ref frameCounter)", "(++frameCounter)")       # an atomic increment that returns the new value
    head = re.sub(r"\bunchecked\(", "(", head)
    head = re.sub(r"\(px, py, threadId\) =>\s*\{", "[&](int px, int py, int threadId) {", head)
    head = head.replace("TraceFull(scene,", "TraceFull(*scene,")
    sel.append("void TraceStage(Vec3 camPosSnapshot, float yawSnapshot, float pitchSnapshot)\n{\n" + head + "\n}\n")
    body = "\n".join(sel)
    body = re.sub(r"^\s*(?:private |public )?(?:readonly )?Scene scene;", "SceneRef *scene = nullptr;", body, flags=re.M)
    # Resize (:110-138): TemporalAA and Framebuffer are classes -> references
    body = re.sub(r"^\s*(?:private |public )?TemporalAA taa;", "TemporalAA *taa = nullptr;", body, flags=re.M)
    body = re.sub(r"\btaa\.(?=[A-Z])", "taa->", body)
    body = re.sub(r"void Resize\(Framebuffer framebuffer,", "void Resize(Framebuffer *framebuffer,", body)
    body = re.sub(r"\bframebuffer\.(?=[A-Z])", "framebuffer->", body)
    body = re.sub(r"\bScene scene\b", "SceneRef &scene", body)
    body = re.sub(r"\bscene is [\w.]*VolumeScene\b", "scene.IsVolumeScene", body)
    body = re.sub(r"\bscene\.Lights\.Count\b", "scene.Lights.Count()", body)
    body = re.sub(r"\bDiffuseTexture\.(?=[A-Z])", "DiffuseTexture->", body)
    body = re.sub(r"new (PathWorkItem|PrimaryGBuffer) \{([^}]*)\}", lambda m_: m_.group(1) + "{" + re.sub(r"(^|,)\s*(\w+) =", r"\1 .\2 =", m_.group(2)) + "}", body)  # object initialisers -> designated initialisers
    nested = re.findall(r"struct (?:PathWorkItem|PrimaryGBuffer)\s*\{[^}]*\}", body)  # hoisted, and kept aggregates (designated initialisers)
    for nsrc in nested:
        body = body.replace(nsrc, "")
    nested = [rewrite(x) + ";" for x in nested]
    body = rewrite(body, None)
    body = re.sub(r"^(\s*)((?:int|bool) \w+);", r"\1\2 = {};", body, flags=re.M)
    out.append("struct RendererRef {\n    Fast2D<Vec3> lastDenoised;\n%s\n%s\n};\n" % ("\n".join(nested), body))
    out.append("} // namespace refcs\n")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    open(out_path, "w", encoding="utf-8").write("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/ConsoleGame", sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ref_generated.hpp"))
