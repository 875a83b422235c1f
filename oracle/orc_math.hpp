// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// load this code, and only as the checker / reported CPU baseline.
//
// PARITY PINS.  The reference ships no test, golden vector or known-answer fixture for the pssmlt / drmlt / PathSampler
// path and its own build cannot run here (Boost/Eigen/Xerces/OpenEXR absent).  Its SOURCES do
// compile, though, from where they lie under /root/reference, once a few arithmetic-free Boost/Eigen headers are stood in
// for (oracle/ref/): oracle/_ref/libref_leaf.so and libref_path.so are the reference's own libcore + librender + libbidir +
// BSDF / emitter / sensor / integrator plugins.  This restatement is pinned against them (tests/test_ref_pins.py, fixtures
// tests/golden/ref_*.npz written by tools/make_ref_golden.py):
//   * numerical leaves (warps, Fresnel, TriAccel, DiscreteDistribution, microfacet, transition kernels): bit for bit;
//   * BSDF plugins sample / eval / pdf: bit for bit (GGX 2 ulp; plastic 1e-7, its derived constants are stored as float);
//   * PathSampler::sampleSplats (MMLT / BDPT / PT) on replayed primary-sample vectors, 14 scene x technique cases:
//     f(u), strategy, splat count, pixel, RGB -- 1e-12 where only diffuse surfaces are hit, >= 99.95 % within 1e-4 else;
//   * whole jobs of the reference's DRMLT / PSSMLT integrators: acceptance-rate counters, b, equal-mutation relMSE.
// NOT pinned sample by sample: the chain step's accept / reject sequence -- the reference draws from SFMT streams seeded
// from /dev/urandom in call order, this restatement (like the product) addresses uniforms by key.
// That -ffp-contract=off is set in the Makefile is what makes the bit-for-bit comparisons possible (the reference's
// default x86-64 build has no FMA contraction).
//
// orc_math.hpp: vectors, RGB spectrum, frames, warps.  Float = double (the reference's default
// CMake build is double precision: data/cmake/MitsubaBuildOptions.cmake:48-70).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>
#include <limits>

namespace orc {

typedef double Float;
static const Float PI = 3.14159265358979323846;
static const Float INV_PI = 0.31830988618379067154;
static const Float INV_TWOPI = 0.15915494309189533577;
static const Float DELTA_EPSILON = 1e-3;            // include/mitsuba/core/constants.h:31
static const Float RCPOVERFLOW = 0x1p-1024;         // constants.h:58 (double build)
static const Float INF = std::numeric_limits<Float>::infinity();

struct Vec2 { Float x, y; Vec2() : x(0), y(0) {} Vec2(Float a, Float b) : x(a), y(b) {} };

struct Vec3 {
    Float x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    explicit Vec3(Float a) : x(a), y(a), z(a) {}
    Vec3(Float a, Float b, Float c) : x(a), y(b), z(c) {}
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Vec3 operator+(const Vec3 &o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
    Vec3 operator-(const Vec3 &o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
    Vec3 operator-() const { return Vec3(-x, -y, -z); }
    Vec3 operator*(Float s) const { return Vec3(x * s, y * s, z * s); }
    Vec3 operator/(Float s) const { const Float r = (Float) 1 / s; return Vec3(x * r, y * r, z * r); }   // TVector3::operator/ (vector.h:535-542)
    Vec3 &operator+=(const Vec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    Vec3 &operator*=(Float s) { x *= s; y *= s; z *= s; return *this; }
    Vec3 &operator/=(Float s) { const Float r = (Float) 1 / s; x *= r; y *= r; z *= r; return *this; }   // vector.h:545-556
};
inline Vec3 operator*(Float s, const Vec3 &v) { return v * s; }
inline Float dot(const Vec3 &a, const Vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Float absDot(const Vec3 &a, const Vec3 &b) { return std::abs(dot(a, b)); }
inline Vec3 cross(const Vec3 &a, const Vec3 &b) {
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline Float lengthSquared(const Vec3 &a) { return dot(a, a); }
inline Float length(const Vec3 &a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(const Vec3 &a) { return a / length(a); }

// RGB spectrum, SPECTRUM_SAMPLES=3 (data/cmake/MitsubaBuildOptions.cmake:54-62)
struct RGB {
    Float r, g, b;
    RGB() : r(0), g(0), b(0) {}
    explicit RGB(Float v) : r(v), g(v), b(v) {}
    RGB(Float a, Float b_, Float c) : r(a), g(b_), b(c) {}
    RGB operator+(const RGB &o) const { return RGB(r + o.r, g + o.g, b + o.b); }
    RGB operator-(const RGB &o) const { return RGB(r - o.r, g - o.g, b - o.b); }
    RGB operator*(const RGB &o) const { return RGB(r * o.r, g * o.g, b * o.b); }
    RGB operator/(const RGB &o) const { return RGB(r / o.r, g / o.g, b / o.b); }
    RGB operator*(Float s) const { return RGB(r * s, g * s, b * s); }
    RGB operator/(Float s) const { const Float rc = (Float) 1 / s; return RGB(r * rc, g * rc, b * rc); }   // TSpectrum::operator/ (spectrum.h:415-425)
    RGB &operator+=(const RGB &o) { r += o.r; g += o.g; b += o.b; return *this; }
    RGB &operator*=(const RGB &o) { r *= o.r; g *= o.g; b *= o.b; return *this; }
    RGB &operator*=(Float s) { r *= s; g *= s; b *= s; return *this; }
    RGB &operator/=(Float s) { const Float rc = (Float) 1 / s; r *= rc; g *= rc; b *= rc; return *this; }
    bool isZero() const { return r == 0 && g == 0 && b == 0; }
    Float max() const { return std::max(r, std::max(g, b)); }
    // include/mitsuba/core/spectrum.h:734-736
    Float luminance() const { return r * 0.212671f + g * 0.715160f + b * 0.072169f; }
    // spectrum.h:467 isValid(): finite and non-negative
    bool isValid() const {
        return std::isfinite(r) && std::isfinite(g) && std::isfinite(b) && r >= 0 && g >= 0 && b >= 0;
    }
    RGB safe_sqrt() const {
        return RGB(std::sqrt(std::max(0.0, r)), std::sqrt(std::max(0.0, g)), std::sqrt(std::max(0.0, b)));
    }
};
inline RGB operator*(Float s, const RGB &v) { return v * s; }

inline Float safe_sqrt(Float v) { return std::sqrt(std::max((Float) 0, v)); }
inline Float safe_acos(Float v) { return std::acos(std::min((Float) 1, std::max((Float) -1, v))); }
inline Float signum(Float v) { return std::signbit(v) ? -1.0 : 1.0; }

// src/libcore/util.cpp:600-609
inline void coordinateSystem(const Vec3 &a, Vec3 &b, Vec3 &c) {
    if (std::abs(a.x) > std::abs(a.y)) {
        Float invLen = 1.0 / std::sqrt(a.x * a.x + a.z * a.z);
        c = Vec3(a.z * invLen, 0.0, -a.x * invLen);
    } else {
        Float invLen = 1.0 / std::sqrt(a.y * a.y + a.z * a.z);
        c = Vec3(0.0, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}

// include/mitsuba/core/frame.h:37-120
struct Frame {
    Vec3 s, t, n;
    Frame() {}
    explicit Frame(const Vec3 &n_) : n(n_) { coordinateSystem(n, s, t); }
    Vec3 toLocal(const Vec3 &v) const { return Vec3(dot(v, s), dot(v, t), dot(v, n)); }
    Vec3 toWorld(const Vec3 &v) const { return s * v.x + t * v.y + n * v.z; }
    static Float cosTheta(const Vec3 &v) { return v.z; }
    static Float cosTheta2(const Vec3 &v) { return v.z * v.z; }
    static Float sinTheta2(const Vec3 &v) { return 1.0 - v.z * v.z; }
    static Float tanTheta(const Vec3 &v) {
        Float temp = 1 - v.z * v.z;
        if (temp <= 0.0) return 0.0;
        return std::sqrt(temp) / v.z;
    }
};

// src/libcore/util.cpp:610-616 computeShadingFrame
inline void computeShadingFrame(const Vec3 &n, const Vec3 &dpdu, Frame &frame) {
    frame.n = n;
    frame.s = normalize(dpdu - frame.n * dot(frame.n, dpdu));
    frame.t = cross(frame.n, frame.s);
}

// ---- warps: src/libcore/warp.cpp:44-100
inline Vec2 squareToUniformDiskConcentric(const Vec2 &sample) {
    Float r1 = 2.0 * sample.x - 1.0, r2 = 2.0 * sample.y - 1.0;
    Float phi, r;
    if (r1 == 0 && r2 == 0) {
        r = phi = 0;
    } else if (r1 * r1 > r2 * r2) {
        r = r1;
        phi = (PI / 4.0) * (r2 / r1);
    } else {
        r = r2;
        phi = (PI / 2.0) - (r1 / r2) * (PI / 4.0);
    }
    return Vec2(r * std::cos(phi), r * std::sin(phi));
}
inline Vec3 squareToCosineHemisphere(const Vec2 &sample) {
    Vec2 p = squareToUniformDiskConcentric(sample);
    Float z = safe_sqrt(1.0 - p.x * p.x - p.y * p.y);
    if (z == 0) z = 1e-10f;
    return Vec3(p.x, p.y, z);
}
inline Float squareToCosineHemispherePdf(const Vec3 &d) { return INV_PI * Frame::cosTheta(d); }
inline Vec2 squareToUniformTriangle(const Vec2 &sample) {
    Float a = safe_sqrt(1.0 - sample.x);
    return Vec2(1 - a, a * sample.y);
}

// ---- Fresnel: src/libcore/util.cpp:659-693, 765-789
inline Float fresnelDielectricExt(Float cosThetaI_, Float &cosThetaT_, Float eta) {
    if (eta == 1) { cosThetaT_ = -cosThetaI_; return 0.0; }
    Float scale = (cosThetaI_ > 0) ? 1 / eta : eta,
          cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0) { cosThetaT_ = 0.0; return 1.0; }
    Float cosThetaI = std::abs(cosThetaI_);
    Float cosThetaT = std::sqrt(cosThetaTSqr);
    Float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    Float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5 * (Rs * Rs + Rp * Rp);
}
inline RGB fresnelConductorExact(Float cosThetaI, const RGB &eta, const RGB &k) {
    Float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1 - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
    RGB temp1 = eta * eta - k * k - RGB(sinThetaI2),
        a2pb2 = (temp1 * temp1 + k * k * eta * eta * 4).safe_sqrt(),
        a = ((a2pb2 + temp1) * 0.5).safe_sqrt();
    RGB term1 = a2pb2 + RGB(cosThetaI2), term2 = a * (2 * cosThetaI);
    RGB Rs2 = (term1 - term2) / (term1 + term2);
    RGB term3 = a2pb2 * cosThetaI2 + RGB(sinThetaI4), term4 = term2 * sinThetaI2;
    RGB Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
    return 0.5 * (Rp2 + Rs2);
}

} // namespace orc
