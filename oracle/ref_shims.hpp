// ref_shims.hpp — the .NET base-class-library surface the transpiled reference sources use (test infrastructure).
//
// oracle/ref_transpile.py rewrites the reference's own C# text (/root/reference/ConsoleGame/...) into C++ mechanically; what
// that text CALLS in the .NET runtime is supplied here, by hand, and is small: MathF / Math (IEEE-754:2019 Max / Min as .NET
// implements them, banker's Math.Round, and the transcendentals — the one documented substitution: MathF.Exp / Log / Pow /
// Sin / Cos / Tan forward to include/ycge_detmath.h exactly as the product and the oracle do, because the platform libm
// behind them is not bit-reproducible), float / int constants, Fast2D<T> (a reference type: assignment aliases, == compares
// identity — the in-place a-trous iteration of RaytraceRenderer.cs:718 depends on that), FixedThreadFor run serially.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <optional>
#include <unordered_map>
#include <vector>
#include "../include/ycge_detmath.h"

namespace refcs {

using byte = uint8_t;
using ushort = uint16_t;

struct Single {
    static constexpr float PositiveInfinity = std::numeric_limits<float>::infinity();
    static constexpr float NegativeInfinity = -std::numeric_limits<float>::infinity();
    static constexpr float MaxValue = std::numeric_limits<float>::max();
    static constexpr float NaN = std::numeric_limits<float>::quiet_NaN();
    static bool IsFinite(float v) { return std::isfinite(v); }
    static bool IsNaN(float v) { return v != v; }
    static bool IsInfinity(float v) { return std::isinf(v); }
};
struct Int32 { static constexpr int MaxValue = 2147483647; static constexpr int MinValue = -2147483647 - 1; };

struct SinCosResult { float Sin, Cos; };
struct MathF {
    static constexpr float PI = 3.14159274f;
    static float Max(float a, float b) { if (a != b) return (a != a) ? a : (b < a ? a : b); return std::signbit(b) ? a : b; }
    static float Min(float a, float b) { if (a != b) return (a != a) ? a : (a < b ? a : b); return std::signbit(a) ? a : b; }
    static float Abs(float a) { return std::fabs(a); }
    static float Sqrt(float a) { return std::sqrt(a); }
    static float Floor(float a) { return std::floor(a); }
    static float Ceiling(float a) { return std::ceil(a); }
    static float Round(float a) { return std::nearbyintf(a); } // MidpointRounding.ToEven under the default rounding mode
    static float CopySign(float a, float b) { return std::copysign(a, b); }
    static float Exp(float a) { return ycge_expf(a); }
    static float Log(float a) { return ycge_logf(a); }
    static float Pow(float a, float b) { return ycge_powf(a, b); }
    static float Sin(float a) { return ycge_sinf(a); }
    static float Cos(float a) { return ycge_cosf(a); }
    static float Tan(float a) { return ycge_tanf(a); }
    static SinCosResult SinCos(float a) { SinCosResult r; ycge_sincosf(a, &r.Sin, &r.Cos); return r; }
};
struct Math {
    static int Max(int a, int b) { return a > b ? a : b; }
    static int Min(int a, int b) { return a < b ? a : b; }
    static double Max(double a, double b) { if (a != b) return (a != a) ? a : (b < a ? a : b); return std::signbit(b) ? a : b; }
    static double Min(double a, double b) { if (a != b) return (a != a) ? a : (a < b ? a : b); return std::signbit(a) ? a : b; }
    static float Max(float a, float b) { if (a != b) return (a != a) ? a : (b < a ? a : b); return std::signbit(b) ? a : b; }   // Math.Max(float, float): C# keeps binary32
    static float Min(float a, float b) { if (a != b) return (a != a) ? a : (a < b ? a : b); return std::signbit(a) ? a : b; }
    static int Abs(int a) { return a < 0 ? -a : a; }
    static double Abs(double a) { return std::fabs(a); }
    static double Round(double a) { return std::nearbyint(a); } // MidpointRounding.ToEven, the default rounding mode
    static double Pow(double a, double b) { return std::pow(a, b); }
    static double Clamp(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
    static double Sqrt(double a) { return std::sqrt(a); }
};

// Fast2D.cs: `public sealed class` -> a handle with reference semantics
template <class T> struct Fast2D {
    struct Impl { int w, h; std::unique_ptr<T[]> d; };
    std::shared_ptr<Impl> p;
    int Width = 0, Height = 0;
    Fast2D() {}
    Fast2D(std::nullptr_t) {}
    Fast2D(int width, int height) : p(new Impl{width, height, std::unique_ptr<T[]>(new T[(size_t)width * height]())}), Width(width), Height(height) {
        if (width <= 0 || height <= 0) throw std::out_of_range("Fast2D");
    }
    T &operator[](int x, int y) const { return p->d[(size_t)x + (size_t)y * p->w]; }
    bool operator==(const Fast2D &o) const { return p.get() == o.p.get(); }
    bool operator!=(const Fast2D &o) const { return p.get() != o.p.get(); }
    bool operator==(std::nullptr_t) const { return !p; }
    bool operator!=(std::nullptr_t) const { return (bool)p; }
    T *Buffer() const { return p->d.get(); }
};

// Renderer/FixedThreadFor.cs: For(from, to, body) -- every worker index once, one thread each (the bodies write disjoint rows)
inline int &ref_threads() { static int n = 1; return n; }
struct FixedThreadFor {
    void For(int from, int to, const std::function<void(int)> &body) {
        if (ref_threads() <= 1 || to - from <= 1) { for (int i = from; i < to; i++) body(i); return; }
        std::vector<std::thread> th;
        for (int i = from; i < to; i++) th.emplace_back([&body, i] { body(i); });
        for (auto &t : th) t.join();
    }
};

struct PixelThreadPool { // Renderer/PixelThreadPool.cs: For2D(w, h, body(px, py, threadId)) -- every pixel once, striped over the worker threads (pixels are independent)
    void For2D(int w, int h, const std::function<void(int, int, int)> &body) {
        const int n = ref_threads();
        if (n <= 1) { for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) body(x, y, 0); return; }
        std::vector<std::thread> th;
        for (int k = 0; k < n; k++) th.emplace_back([&body, k, n, w, h] { for (long long p = k; p < (long long)w * h; p += n) body((int)(p % w), (int)(p / w), k); });
        for (auto &t : th) t.join();
    }
};
enum ConsoleColor : int { Black = 0 };
// System.Console: the window size ANSITerminalRenderer.Render compares its own with (set by the harness)
struct Console { static inline int WindowWidth = 0, WindowHeight = 0; };
struct Buffer { // System.Buffer.BlockCopy on byte arrays
    static void BlockCopy(const std::vector<byte> &src, int src_offset, std::vector<byte> &dst, int dst_offset, int count) { std::memcpy(dst.data() + dst_offset, src.data() + src_offset, (size_t)count); }
};

// System.Collections.Generic.List<T>: a reference type in C# -- the transpiler passes it by reference
template <class T> struct List {
    std::vector<T> v;
    List() {}
    explicit List(int capacity) { v.reserve((size_t)capacity); }
    void Add(const T &x) { v.push_back(x); }
    int Count() const { return (int)v.size(); }
    std::vector<T> ToArray() const { return v; }
    T &operator[](int i) { return v[(size_t)i]; }
    const T &operator[](int i) const { return v[(size_t)i]; }
};
// ---- what MeshLoader.FromObj / MeshScenes.TryReadObjBoundsNormalized call: strings, a line reader, number parsing, and the
// collections whose enumeration order the text relies on (HashSet<int> / Dictionary<K, V> without removals enumerate in insertion order)
struct String {
    std::string s;
    bool null = true;
    String() {}
    String(const char *c) : s(c), null(false) {}
    String(std::string v) : s(std::move(v)), null(false) {}
    size_t size() const { return s.size(); }                       // .Length
    char operator[](int i) const { return s[(size_t)i]; }
    bool operator==(const char *c) const { return !null && s == c; }
    bool operator!=(std::nullptr_t) const { return !null; }
    bool operator==(std::nullptr_t) const { return null; }
    static bool IsWhite(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }
    static bool IsNullOrEmpty(const String &v) { return v.null || v.s.empty(); }
    static bool IsNullOrWhiteSpace(const String &v) { if (v.null) return true; for (char c : v.s) if (!IsWhite(c)) return false; return true; }
    std::vector<String> SplitWhitespace() const {                   // Split((char[])null, StringSplitOptions.RemoveEmptyEntries)
        std::vector<String> out;
        size_t i = 0;
        while (i < s.size()) {
            while (i < s.size() && IsWhite(s[i])) i++;
            size_t j = i;
            while (j < s.size() && !IsWhite(s[j])) j++;
            if (j > i) out.push_back(String(s.substr(i, j - i)));
            i = j;
        }
        return out;
    }
    std::vector<String> Split(char sep) const {                      // Split('/'): empty entries kept
        std::vector<String> out;
        size_t i = 0;
        for (;;) {
            size_t j = s.find(sep, i);
            out.push_back(String(s.substr(i, j == std::string::npos ? std::string::npos : j - i)));
            if (j == std::string::npos) break;
            i = j + 1;
        }
        return out;
    }
};
struct NumberFormatInfo {};
struct CultureInfo { NumberFormatInfo NumberFormat; static const CultureInfo InvariantCulture; };
inline const CultureInfo CultureInfo::InvariantCulture{};
struct File { static bool Exists(const String &p) { FILE *f = std::fopen(p.s.c_str(), "rb"); if (!f) return false; std::fclose(f); return true; } };
struct StreamReader {
    FILE *f;
    explicit StreamReader(const String &p) : f(std::fopen(p.s.c_str(), "rb")) {}
    ~StreamReader() { if (f) std::fclose(f); }
    String ReadLine() {                                              // a line without its terminator (\n, \r\n), null at the end of the file
        if (!f) return String();
        std::string line;
        int c = std::fgetc(f);
        if (c == EOF) return String();
        while (c != EOF && c != '\n') { line.push_back((char)c); c = std::fgetc(f); }
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return String(line);
    }
};
struct BinaryWriter { // System.IO.BinaryWriter over a new file: Write(char) = the character's UTF-8 bytes, Write(int) = 4 bytes little endian
    FILE *f;
    explicit BinaryWriter(const String &p) : f(std::fopen(p.s.c_str(), "wb")) { if (!f) throw std::runtime_error("cannot create file"); }
    ~BinaryWriter() { if (f) std::fclose(f); }
    void Write(char c) { std::fputc((unsigned char)c, f); }   // ASCII only here ('V', 'G', '0', '1')
    void Write(int v) { unsigned char b[4] = {(unsigned char)(v & 255), (unsigned char)((v >> 8) & 255), (unsigned char)((v >> 16) & 255), (unsigned char)((v >> 24) & 255)}; std::fwrite(b, 1, 4, f); }
};
template <class Fmt> float SingleParse(const String &t, const Fmt &) { return std::strtof(t.s.c_str(), nullptr); }   // float.Parse(s, invariant): correctly rounded (.NET Core 3.0+), as strtof
template <class Fmt> int Int32Parse(const String &t, const Fmt &) { return (int)std::strtol(t.s.c_str(), nullptr, 10); }
template <class T> struct HashSet {
    std::vector<T> order;
    std::unordered_map<T, int> index;
    bool Add(const T &x) { if (index.count(x)) return false; index[x] = (int)order.size(); order.push_back(x); return true; }
    int Count() const { return (int)order.size(); }
    typename std::vector<T>::const_iterator begin() const { return order.begin(); }
    typename std::vector<T>::const_iterator end() const { return order.end(); }
};
template <class K, class V> struct KeyValuePair { K Key; V Value; };
template <class K, class V> struct Dictionary {
    std::vector<KeyValuePair<K, V>> items;
    std::unordered_map<K, int> index;
    Dictionary() {}
    explicit Dictionary(int) {}
    bool TryGetValue(const K &k, V &out) const { auto it = index.find(k); if (it == index.end()) return false; out = items[(size_t)it->second].Value; return true; }
    V &operator[](const K &k) { auto it = index.find(k); if (it != index.end()) return items[(size_t)it->second].Value; index[k] = (int)items.size(); items.push_back({k, V()}); return items.back().Value; }
    typename std::vector<KeyValuePair<K, V>>::iterator begin() { return items.begin(); }
    typename std::vector<KeyValuePair<K, V>>::iterator end() { return items.end(); }
};
// a List<T> held by several names at once (a C# class instance): copies alias
template <class T> struct RList {
    std::shared_ptr<std::vector<T>> p;
    RList() {}
    explicit RList(int capacity) : p(new std::vector<T>()) { p->reserve((size_t)capacity); }
    void Add(const T &x) { p->push_back(x); }
    int Count() const { return (int)p->size(); }
    T &operator[](int i) const { return (*p)[(size_t)i]; }
};
struct Face3 { int a, b, c; }; // the value tuple (int a, int b, int c)
// ---- WorldGeneration: rectangular arrays (C# arrays are references: copies alias; zero-initialised), the (int, int) block tuple, sbyte
using sbyte = int8_t;
struct Cell2 { int Item1 = 0, Item2 = 0; };
struct OrderItem { int x, z, h; };
inline int Int32CompareTo(int a, int b) { return a < b ? -1 : (a > b ? 1 : 0); }
template <class T> struct Array2 {
    std::shared_ptr<std::vector<T>> p;
    int n0 = 0, n1 = 0;
    Array2() {}
    Array2(int a, int b) : p(new std::vector<T>((size_t)a * b)), n0(a), n1(b) {}
    T &operator[](int i, int j) const { return (*p)[(size_t)i * n1 + j]; }
    int GetLength(int d) const { return d == 0 ? n0 : n1; }
};
template <class T> struct Array3 {
    std::shared_ptr<std::vector<T>> p;
    int n0 = 0, n1 = 0, n2 = 0;
    Array3() {}
    Array3(int a, int b, int c) : p(new std::vector<T>((size_t)a * b * c)), n0(a), n1(b), n2(c) {}
    T &operator[](int i, int j, int k) const { return (*p)[((size_t)i * n1 + j) * n2 + k]; }
};
template <class T> using Comparison = std::function<int(const T &, const T &)>;
inline int SingleCompareTo(float a, float b) { // System.Single.CompareTo
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    if (a != a) return (b != b) ? 0 : -1;
    return 1;
}
// System.Array.Sort(keys, index, length, comparer) = dotnet/runtime ArraySortHelper<T>.IntrospectiveSort (.NET 8): insertion sort up to
// 16 elements, heap sort at depth 0, else median-of-three partition.  Unstable: the PERMUTATION it produces is part of the reference's
// behaviour (the builders' fallback split), hence restated here as runtime library.
struct Array {
    template <class T, class Cmp> static void Sort(std::vector<T> &keys, Cmp cmp) { Sort(keys, 0, (int)keys.size(), cmp); }                     // Array.Sort(array, comparison): the same introsort
    template <class T> static void Resize(std::vector<T> &a, int n) { a.resize((size_t)n); } // Array.Resize(ref a, n): contents kept, zero-filled tail
    template <class T, class Cmp> static void Sort(std::vector<T> &keys, int index, int length, Cmp cmp) {
        if (length < 2) return;
        int lg = 0;
        for (unsigned v = (unsigned)length; v > 1; v >>= 1) lg++;
        Intro(keys.data() + index, length, 2 * (lg + 1), cmp);
    }
    template <class T, class Cmp> static void SwapIfGreater(T *a, int i, int j, Cmp &cmp) { if (cmp(a[i], a[j]) > 0) std::swap(a[i], a[j]); }
    template <class T, class Cmp> static void Insertion(T *a, int n, Cmp &cmp) {
        for (int i = 0; i + 1 < n; i++) {
            T t = a[i + 1];
            int j = i;
            for (; j >= 0 && cmp(t, a[j]) < 0; j--) a[j + 1] = a[j];
            a[j + 1] = t;
        }
    }
    template <class T, class Cmp> static void DownHeap(T *a, int i, int n, Cmp &cmp) {
        T d = a[i - 1];
        while (i <= n / 2) {
            int ch = 2 * i;
            if (ch < n && cmp(a[ch - 1], a[ch]) < 0) ch++;
            if (!(cmp(d, a[ch - 1]) < 0)) break;
            a[i - 1] = a[ch - 1];
            i = ch;
        }
        a[i - 1] = d;
    }
    template <class T, class Cmp> static void Heap(T *a, int n, Cmp &cmp) {
        for (int i = n / 2; i >= 1; i--) DownHeap(a, i, n, cmp);
        for (int i = n; i > 1; i--) { std::swap(a[0], a[i - 1]); DownHeap(a, 1, i - 1, cmp); }
    }
    template <class T, class Cmp> static int Partition(T *a, int n, Cmp &cmp) {
        int hi = n - 1, mid = hi >> 1;
        SwapIfGreater(a, 0, mid, cmp); SwapIfGreater(a, 0, hi, cmp); SwapIfGreater(a, mid, hi, cmp);
        T pivot = a[mid];
        std::swap(a[mid], a[hi - 1]);
        int l = 0, r = hi - 1;
        while (l < r) {
            while (cmp(a[++l], pivot) < 0) {}
            while (cmp(pivot, a[--r]) < 0) {}
            if (l >= r) break;
            std::swap(a[l], a[r]);
        }
        if (l != hi - 1) std::swap(a[l], a[hi - 1]);
        return l;
    }
    template <class T, class Cmp> static void Intro(T *a, int n, int depth, Cmp &cmp) {
        while (n > 1) {
            if (n <= 16) {
                if (n == 2) { SwapIfGreater(a, 0, 1, cmp); return; }
                if (n == 3) { SwapIfGreater(a, 0, 1, cmp); SwapIfGreater(a, 0, 2, cmp); SwapIfGreater(a, 1, 2, cmp); return; }
                Insertion(a, n, cmp);
                return;
            }
            if (depth == 0) { Heap(a, n, cmp); return; }
            depth--;
            int p = Partition(a, n, cmp);
            Intro(a + p + 1, n - (p + 1), depth, cmp);
            n = p;
        }
    }
};

} // namespace refcs
