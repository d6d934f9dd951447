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
#include <vector>
#include "../include/ycge_detmath.h"

namespace refcs {

using byte = uint8_t;

struct Single {
    static constexpr float PositiveInfinity = std::numeric_limits<float>::infinity();
    static constexpr float NegativeInfinity = -std::numeric_limits<float>::infinity();
    static constexpr float MaxValue = std::numeric_limits<float>::max();
    static bool IsFinite(float v) { return std::isfinite(v); }
    static bool IsNaN(float v) { return v != v; }
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

struct FixedThreadFor { // Renderer/FixedThreadFor.cs: For(from, to, body) -- every worker index once; serial here (the bodies write disjoint rows)
    void For(int from, int to, const std::function<void(int)> &body) { for (int i = from; i < to; i++) body(i); }
};

enum ConsoleColor : int { Black = 0 };

} // namespace refcs
