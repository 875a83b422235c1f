// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
//
// orc_bsdf.hpp: the five BSDF plugins on the hot path, in the local shading frame.
// Restates src/bsdfs/diffuse.cpp:109-150, src/bsdfs/dielectric.cpp:220-340,
// src/bsdfs/conductor.cpp:223-285, src/bsdfs/roughconductor.cpp:250-412,
// src/bsdfs/microfacet.h:191-700 (Beckmann + GGX, sampleAll / sampleVisible),
// src/bsdfs/twosided.cpp:107-195; plastic.cpp:240-420; roughplastic.cpp:325-491 with rtrans.h:136-146 and spline.cpp:23-60.
#pragma once
#include "orc_math.hpp"
#include "../include/drmlt_b200.h"
#include <array>
#include <deque>
#include <mutex>

namespace orc {

enum ETransportMode { ERadiance = 0, EImportance = 1 };
enum EMeasure { EInvalidMeasure = 0, ESolidAngle = 1, ELength = 2, EArea = 3, EDiscrete = 4 };
enum EBSDFType {
    EDiffuseReflection = 0x1, EGlossyReflection = 0x2, EDeltaReflection = 0x4, EDeltaTransmission = 0x8, EGlossyTransmission = 0x10,
    ESmooth = EDiffuseReflection | EGlossyReflection | EGlossyTransmission, EDelta = EDeltaReflection | EDeltaTransmission
};

struct BSDFRecord {
    Vec3 wi, wo;
    int mode;
    int sampledType = 0;
    Float eta = 1.0;
    BSDFRecord(const Vec3 &wi_, int mode_) : wi(wi_), mode(mode_) {}
    BSDFRecord(const Vec3 &wi_, const Vec3 &wo_, int mode_) : wi(wi_), wo(wo_), mode(mode_) {}
    void reverse() { std::swap(wi, wo); mode = 1 - mode; }   // include/mitsuba/render/bsdf.h reverse()
};

inline RGB rgb3(const float *v) { return RGB(v[0], v[1], v[2]); }

// A material as the BSDF functions see it: the POD of the ABI plus its two colour parameters in Float -- the constants of the
// struct, or, for a textured parameter (DR_MAT_TEX_*), the bitmap lookup at the intersection (Scene::material, orc_scene.hpp),
// which must not be rounded to float32.
struct Mat : dr_material {
    RGB R, T;
    Mat(const dr_material &m) : dr_material(m), R(rgb3(m.reflectance)), T(rgb3(m.transmittance)) {}
};
// the material at an intersection: m_reflectance->eval(bRec.its) etc. of the BSDF plugins
template <class SceneT, class ItsT> inline Mat materialAt(const SceneT &sc, const ItsT &its) {
    Mat m(sc.mats[its.material]);
    const uint32_t tr = (m.flags >> 8) & 0xfffu, tt = m.flags >> 20;
    if (tr) m.R = sc.textures[tr - 1].eval(its.uv);
    if (tt) m.T = sc.textures[tt - 1].eval(its.uv);
    return m;
}

inline bool bsdfHasSmooth(const dr_material &m) { return m.type == DR_BSDF_DIFFUSE || m.type == DR_BSDF_ROUGHCONDUCTOR || m.type == DR_BSDF_ROUGHDIELECTRIC || m.type == DR_BSDF_PLASTIC || m.type == DR_BSDF_ROUGHPLASTIC; }
inline bool bsdfNonSymmetric(const dr_material &m) { return m.type == DR_BSDF_DIELECTRIC || m.type == DR_BSDF_ROUGHDIELECTRIC; }   // dielectric.cpp:201, roughdielectric.cpp:253
// BSDF::EUsesSampler: sample() draws one more number from the vertex's sampler (roughdielectric.cpp:464, 555)
inline bool bsdfUsesSampler(const dr_material &m) { return m.type == DR_BSDF_ROUGHDIELECTRIC; }
// BSDF::ETransmission | BSDF::EBackSide test used by DirectSamplingRecord (records.inl:160-164)
inline bool bsdfTransmissiveOrBackside(const dr_material &m) { return m.type == DR_BSDF_DIELECTRIC || m.type == DR_BSDF_ROUGHDIELECTRIC || (m.flags & DR_MAT_TWOSIDED); }
inline int bsdfMeasure(int sampledType) { return (sampledType & EDelta) ? EDiscrete : ESolidAngle; }

// ---------------------------------------------------------------- microfacet.h
struct Microfacet {
    bool ggx, sampleVis;
    Float alpha;
    Microfacet(const dr_material &m)
        : ggx((m.flags & DR_MAT_GGX) != 0), sampleVis((m.flags & DR_MAT_SAMPLE_VISIBLE) != 0),
          alpha(std::max(average3((Float) m.alpha), (Float) 1e-4f)) {}   // microfacet.h:67-72
    // alpha is a ConstantFloatTexture evaluated to a Spectrum and averaged: (a + a + a) * (1.0f / 3), a FLOAT third
    // (roughconductor.cpp:273, spectrum.h:481-486) -- the effective roughness is alpha * (1 + 3e-8)
    static Float average3(Float a) { Float r = 0.0; r += a; r += a; r += a; return r * (1.0f / 3); }
    void scaleAlpha(Float value) { alpha *= value; }          // :178-183

    Float eval(const Vec3 &m) const {   // :191-237
        if (Frame::cosTheta(m) <= 0) return 0.0;
        Float cosTheta2 = Frame::cosTheta2(m);
        Float beckmannExponent = ((m.x * m.x) / (alpha * alpha) + (m.y * m.y) / (alpha * alpha)) / cosTheta2;
        Float result;
        if (!ggx) {
            result = std::exp(-beckmannExponent) / (PI * alpha * alpha * cosTheta2 * cosTheta2);
        } else {
            Float root = (1.0 + beckmannExponent) * cosTheta2;
            result = 1.0 / (PI * alpha * alpha * root * root);
        }
        if (result * Frame::cosTheta(m) < 1e-20f) result = 0;
        return result;
    }
    Float smithG1(const Vec3 &v, const Vec3 &m) const {   // :476-513
        if (dot(v, m) * Frame::cosTheta(v) <= 0) return 0.0;
        Float tanTheta = std::abs(Frame::tanTheta(v));
        if (tanTheta == 0.0) return 1.0;
        if (!ggx) {
            Float a = 1.0 / (alpha * tanTheta);
            if (a >= 1.6f) return 1.0;
            Float aSqr = a * a;
            return (3.535f * a + 2.181f * aSqr) / (1.0 + 2.276f * a + 2.577f * aSqr);
        } else {
            Float root = alpha * tanTheta;
            return 2.0 / (1.0 + std::hypot((Float) 1.0, root));
        }
    }
    Float G(const Vec3 &wi, const Vec3 &wo, const Vec3 &m) const { return smithG1(wi, m) * smithG1(wo, m); }
    Float pdfAll(const Vec3 &m) const { return eval(m) * Frame::cosTheta(m); }
    Float pdfVisible(const Vec3 &wi, const Vec3 &m) const {   // :469-473
        if (Frame::cosTheta(wi) == 0) return 0.0;
        return smithG1(wi, m) * absDot(wi, m) * eval(m) / std::abs(Frame::cosTheta(wi));
    }
    Float pdf(const Vec3 &wi, const Vec3 &m) const { return sampleVis ? pdfVisible(wi, m) : pdfAll(m); }

    Vec3 sampleAll(const Vec2 &sample, Float &pdf) const {   // :287-397 (isotropic)
        Float sinPhiM = std::sin(2.0 * PI * sample.y), cosPhiM = std::cos(2.0 * PI * sample.y);
        Float alphaSqr = alpha * alpha, cosThetaM;
        if (!ggx) {
            Float tanThetaMSqr = alphaSqr * -std::log(1.0 - sample.x);
            cosThetaM = 1.0 / std::sqrt(1.0 + tanThetaMSqr);
            pdf = (1.0 - sample.x) / (PI * alpha * alpha * cosThetaM * cosThetaM * cosThetaM);
        } else {
            Float tanThetaMSqr = alphaSqr * sample.x / (1.0 - sample.x);
            cosThetaM = 1.0 / std::sqrt(1.0 + tanThetaMSqr);
            Float temp = 1 + tanThetaMSqr / alphaSqr;
            pdf = INV_PI / (alpha * alpha * cosThetaM * cosThetaM * cosThetaM * temp * temp);
        }
        if (pdf < 1e-20f) pdf = 0;
        Float sinThetaM = std::sqrt(std::max((Float) 0, 1 - cosThetaM * cosThetaM));
        return Vec3(sinThetaM * cosPhiM, sinThetaM * sinPhiM, cosThetaM);
    }

    // Numerical inverse of erf used by the Beckmann visible-normal sampler (src/libcore/math.cpp:25-80,
    // Giles' single-precision polynomial as used by the reference)
    static Float erfinv(Float x) {
        Float w = -std::log((1.0 - x) * (1.0 + x));
        Float p;
        if (w < 5.0) {
            w = w - 2.5;
            p = 2.81022636e-08;
            p = 3.43273939e-07 + p * w; p = -3.5233877e-06 + p * w; p = -4.39150654e-06 + p * w;
            p = 0.00021858087 + p * w; p = -0.00125372503 + p * w; p = -0.00417768164 + p * w;
            p = 0.246640727 + p * w; p = 1.50140941 + p * w;
        } else {
            w = std::sqrt(w) - 3.0;
            p = -0.000200214257;
            p = 0.000100950558 + p * w; p = 0.00134934322 + p * w; p = -0.00367342844 + p * w;
            p = 0.00573950773 + p * w; p = -0.0076224613 + p * w; p = 0.00943887047 + p * w;
            p = 1.00167406 + p * w; p = 2.83297682 + p * w;
        }
        return p * x;
    }
    static Float erfApprox(Float x) {   // src/libcore/math.cpp erf(): A&S 7.1.26-style rational
        Float a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741, a4 = -1.453152027, a5 = 1.061405429, p = 0.3275911;
        Float sign = signum(x);
        x = std::abs(x);
        Float t = 1.0 / (1.0 + p * x);
        Float y = 1.0 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * std::exp(-x * x);
        return sign * y;
    }

    Vec2 sampleVisible11(Float thetaI, Vec2 sample, Float epsilon) const {   // :573-690
        const Float SQRT_PI_INV = 1 / std::sqrt(PI);
        Vec2 slope;
        if (!ggx) {
            if (thetaI < 1e-4f) {
                Float r = std::sqrt(-std::log(1.0 - sample.x));
                return Vec2(r * std::cos(2 * PI * sample.y), r * std::sin(2 * PI * sample.y));
            }
            Float tanThetaI = std::tan(thetaI), cotThetaI = 1 / tanThetaI;
            Float a = -1, c = erfApprox(cotThetaI);
            Float sample_x = std::max(sample.x, (Float) 1e-6f);
            Float fit = 1 + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
            Float b = c - (1 + c) * std::pow(1 - sample_x, fit);
            Float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * std::exp(-cotThetaI * cotThetaI));
            int it = 0;
            while (++it < 10) {
                if (!(b >= a && b <= c)) b = 0.5 * (a + c);
                Float invErf = erfinv(b);
                Float value = normalization * (1 + b + SQRT_PI_INV * tanThetaI * std::exp(-invErf * invErf)) - sample_x;
                Float derivative = normalization * (1 - invErf * tanThetaI);
                if (std::abs(value) < 1e-5f) break;
                if (value > 0) c = b; else a = b;
                b -= value / derivative;
            }
            slope.x = erfinv(b);
            slope.y = erfinv(2.0 * std::max(sample.y, (Float) 1e-6f) - 1.0);
        } else {
            if (thetaI < 1e-4f) {
                Float r = safe_sqrt(sample.x / (1 - sample.x));
                return Vec2(r * std::cos(2 * PI * sample.y), r * std::sin(2 * PI * sample.y));
            }
            Float tanThetaI = std::tan(thetaI);
            Float a = 1 / tanThetaI;
            Float G1 = 2.0 / (1.0 + safe_sqrt(1.0 + 1.0 / (a * a)));
            Float A = 2.0 * sample.x / G1 - 1.0;
            if (std::abs(A) == 1) A -= signum(A) * epsilon;
            Float tmp = 1.0 / (A * A - 1.0);
            Float B = tanThetaI;
            Float D = safe_sqrt(B * B * tmp * tmp - (A * A - B * B) * tmp);
            Float slope_x_1 = B * tmp - D, slope_x_2 = B * tmp + D;
            slope.x = (A < 0.0 || slope_x_2 > 1.0 / tanThetaI) ? slope_x_1 : slope_x_2;
            Float S;
            if (sample.y > 0.5) { S = 1.0; sample.y = 2.0 * (sample.y - 0.5); }
            else { S = -1.0; sample.y = 2.0 * (0.5 - sample.y); }
            Float z = (sample.y * (sample.y * (sample.y * (-0.365728915865723) + 0.790235037209296) - 0.424965825137544) + 0.000152998850436920) /
                      (sample.y * (sample.y * (sample.y * (sample.y * 0.169507819808272 - 0.397203533833404) - 0.232500544458471) + 1) - 0.539825872510702);
            slope.y = S * z * std::sqrt(1.0 + slope.x * slope.x);
        }
        return slope;
    }
    Vec3 sampleVisible(const Vec3 &_wi, const Vec2 &sample, Float epsilon) const {   // :417-466
        Vec3 wi = normalize(Vec3(alpha * _wi.x, alpha * _wi.y, _wi.z));
        Float theta = 0, phi = 0;
        if (wi.z < 0.99999) { theta = std::acos(wi.z); phi = std::atan2(wi.y, wi.x); }
        Float sinPhi = std::sin(phi), cosPhi = std::cos(phi);
        Vec2 slope = sampleVisible11(theta, sample, epsilon);
        slope = Vec2(cosPhi * slope.x - sinPhi * slope.y, sinPhi * slope.x + cosPhi * slope.y);
        slope.x *= alpha; slope.y *= alpha;
        Float normalization = 1.0 / std::sqrt(slope.x * slope.x + slope.y * slope.y + 1.0);
        return Vec3(-slope.x * normalization, -slope.y * normalization, normalization);
    }
    Vec3 sample(const Vec3 &wi, const Vec2 &s, Float &pdf, Float epsilon) const {   // :243-252
        if (sampleVis) { Vec3 m = sampleVisible(wi, s, epsilon); pdf = pdfVisible(wi, m); return m; }
        return sampleAll(s, pdf);
    }
};

// ---------------------------------------------------------------- nested (one-sided) models
namespace detail {

inline Vec3 reflectZ(const Vec3 &wi) { return Vec3(-wi.x, -wi.y, wi.z); }
inline Vec3 reflectM(const Vec3 &wi, const Vec3 &m) { return m * (2 * dot(wi, m)) - wi; }

// ---------------------------------------------------------------- smooth plastic (plastic.cpp)
// fresnelDiffuseReflectance(eta, fast = false) (util.cpp:815-867): integral over xi in [0,1] of F(sqrt(xi), eta); the
// reference uses an adaptive Gauss-Lobatto rule (relative error 1e-5), here composite Simpson in x = sqrt(xi).
inline Float fresnelDiffuseReflectance(Float eta) {
    const int n = 1 << 16;
    const Float h = 1.0 / n;
    auto f = [&](Float x) { Float ct; return fresnelDielectricExt(x, ct, eta) * 2.0 * x; };
    Float s = f(0.0) + f(1.0);
    for (int i = 1; i < n; ++i) s += f(i * h) * ((i & 1) ? 4.0 : 2.0);
    return s * h / 3.0;
}
// SmoothPlastic::configure (plastic.cpp:188-205): m_fdrInt -> k[0], m_specularSamplingWeight -> k[1]
inline void preparePlastic(dr_material &m) {
    if (m.type != DR_BSDF_PLASTIC) return;
    m.k[0] = (float) fresnelDiffuseReflectance(1.0 / (Float) m.eta[0]);
    Float dAvg = rgb3(m.reflectance).luminance(), sAvg = rgb3(m.transmittance).luminance();
    m.k[1] = (float) (sAvg / (dAvg + sAvg));
}
struct Plastic {       // SmoothPlastic::configure (plastic.cpp:188-205); dr_material: reflectance = diffuse, transmittance = specular
    Float eta, invEta2, fdrInt, specularSamplingWeight;
    RGB diffuse, specular;
    bool nonlinear;
    explicit Plastic(const Mat &m) {
        eta = m.eta[0]; invEta2 = 1 / (eta * eta);
        // derived once per material by preparePlastic (below), stored as floats like the library's device copy
        fdrInt = m.k[0]; specularSamplingWeight = m.k[1];
        diffuse = m.R; specular = m.T;
        nonlinear = (m.flags & DR_MAT_NONLINEAR) != 0;
    }
    Float probSpecular(Float Fi) const { return (Fi * specularSamplingWeight) / (Fi * specularSamplingWeight + (1 - Fi) * (1 - specularSamplingWeight)); }
    RGB diff() const {
        if (nonlinear) return RGB(diffuse.r / (1 - diffuse.r * fdrInt), diffuse.g / (1 - diffuse.g * fdrInt), diffuse.b / (1 - diffuse.b * fdrInt));
        return diffuse * (1.0 / (1 - fdrInt)) ;
    }
};

// ---------------------------------------------------------------- rough plastic (roughplastic.cpp, rtrans.h)
// The oracle keeps every table it was given in a process-wide, append-only registry; prepareRoughPlastic re-points
// dr_material.table at the registry entry, so that the BSDF functions need nothing but the material.
typedef std::array<double, DR_ROUGH_TABLE_DOUBLES> RoughTable;
inline std::deque<RoughTable> &roughTableRegistry() { static std::deque<RoughTable> r; return r; }
inline uint32_t registerRoughTable(const double *t) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    RoughTable a;
    for (int i = 0; i < DR_ROUGH_TABLE_DOUBLES; ++i) a[i] = t[i];
    roughTableRegistry().push_back(a);
    return (uint32_t) (roughTableRegistry().size() - 1);
}
// RoughPlastic::configure (roughplastic.cpp:273-277): m_specularSamplingWeight -> slot [102] of the material's own copy of the table
inline void prepareRoughPlastic(dr_material &m, const double *sceneTables) {
    if (m.type != DR_BSDF_ROUGHPLASTIC) return;
    RoughTable t;
    for (int i = 0; i < DR_ROUGH_TABLE_DOUBLES; ++i) t[i] = sceneTables[(size_t) m.table * DR_ROUGH_TABLE_DOUBLES + i];
    const Float dAvg = RGB(m.reflectance[0], m.reflectance[1], m.reflectance[2]).luminance();
    const Float sAvg = RGB(m.transmittance[0], m.transmittance[1], m.transmittance[2]).luminance();
    t[102] = sAvg / (dAvg + sAvg);
    m.table = registerRoughTable(t.data());
}
// evalCubicInterp1D(x, values, size, min = 0, max = 1) (src/libcore/spline.cpp:23-60)
inline Float evalCubicInterp1D(Float x, const double *values, size_t size, Float min, Float max) {
    if (!(x >= min && x <= max)) return 0.0;
    Float t = ((x - min) * (size - 1)) / (max - min);
    size_t k = std::max((size_t) 0, std::min((size_t) t, size - 2));
    Float f0 = values[k], f1 = values[k + 1], d0, d1;
    if (k > 0) d0 = 0.5f * (values[k + 1] - values[k - 1]); else d0 = values[k + 1] - values[k];
    if (k + 2 < size) d1 = 0.5f * (values[k + 2] - values[k]); else d1 = values[k + 1] - values[k];
    t = t - (Float) k;
    Float t2 = t * t, t3 = t2 * t;
    return (2 * t3 - 3 * t2 + 1) * f0 + (-2 * t3 + 3 * t2) * f1 + (t3 - 2 * t2 + t) * d0 + (t3 - t2) * d1;
}
struct RoughPlastic {   // dr_material: reflectance = diffuse, transmittance = specular, eta[0], alpha, table
    const double *tab;
    Float eta, invEta2, specularSamplingWeight;
    RGB diffuse, specular;
    bool nonlinear;
    explicit RoughPlastic(const Mat &m) {
        tab = roughTableRegistry()[m.table].data();
        eta = m.eta[0]; invEta2 = 1.0f / (eta * eta);                  // roughplastic.cpp:279
        specularSamplingWeight = tab[102];
        diffuse = m.R; specular = m.T;
        nonlinear = (m.flags & DR_MAT_NONLINEAR) != 0;
    }
    // m_externalRoughTransmittance->eval(cosTheta, alpha) with eta and alpha fixed (rtrans.h:136-146, 192)
    Float T(Float cosTheta) const {
        Float warpedCosTheta = std::pow(std::abs(cosTheta), (Float) 0.25f);
        if (!(cosTheta >= 0)) return 0.0;
        Float result = evalCubicInterp1D(warpedCosTheta, tab, DR_ROUGH_TABLE_THETA, 0.0f, 1.0f);
        return std::min((Float) 1.0f, std::max((Float) 0.0f, result));
    }
    Float probSpecular(Float cosThetaI) const {                       // roughplastic.cpp:407-416
        Float p = 1 - T(cosThetaI);
        return (p * specularSamplingWeight) / (p * specularSamplingWeight + (1 - p) * (1 - specularSamplingWeight));
    }
    RGB eval(const dr_material &m, const Vec3 &wi, const Vec3 &wo) const {   // :339-381
        Microfacet distr(m);
        const Vec3 H = normalize(wo + wi);
        const Float D = distr.eval(H);
        Float ct;
        const Float F = fresnelDielectricExt(dot(wi, H), ct, eta);
        const Float G = distr.G(wi, wo, H);
        Float value = F * D * G / (4.0f * Frame::cosTheta(wi));
        RGB result = specular * value;
        RGB diff = diffuse;
        Float T12 = T(Frame::cosTheta(wi)), T21 = T(Frame::cosTheta(wo));
        Float Fdr = 1 - tab[DR_ROUGH_TABLE_THETA];                     // 1 - m_internalRoughTransmittance->evalDiffuse(alpha)
        if (nonlinear) diff = RGB(diff.r / (1 - diff.r * Fdr), diff.g / (1 - diff.g * Fdr), diff.b / (1 - diff.b * Fdr));
        else diff = diff * (1.0 / (1 - Fdr));
        return result + diff * (INV_PI * Frame::cosTheta(wo) * T12 * T21 * invEta2);
    }
    Float pdf(const dr_material &m, const Vec3 &wi, const Vec3 &wo) const {   // :396-432
        Microfacet distr(m);
        const Vec3 H = normalize(wo + wi);
        Float pS = probSpecular(Frame::cosTheta(wi)), pD = 1 - pS;
        const Float dwh_dwo = 1.0f / (4.0f * dot(wo, H));
        const Float prob = distr.pdf(wi, H);
        Float result = prob * dwh_dwo * pS;
        result += pD * squareToCosineHemispherePdf(wo);
        return result;
    }
};

inline RGB evalNested(const Mat &m, const BSDFRecord &b, int measure) {
    switch (m.type) {
    case DR_BSDF_DIFFUSE:   // diffuse.cpp:109-117
        if (measure != ESolidAngle || Frame::cosTheta(b.wi) <= 0 || Frame::cosTheta(b.wo) <= 0) return RGB(0.0);
        return m.R * (INV_PI * Frame::cosTheta(b.wo));
    case DR_BSDF_CONDUCTOR:   // conductor.cpp:223-237
        if (measure != EDiscrete || Frame::cosTheta(b.wi) <= 0 || Frame::cosTheta(b.wo) <= 0 ||
            std::abs(dot(reflectZ(b.wi), b.wo) - 1) > DELTA_EPSILON) return RGB(0.0);
        return m.R * fresnelConductorExact(Frame::cosTheta(b.wi), rgb3(m.eta), rgb3(m.k));
    case DR_BSDF_ROUGHCONDUCTOR: {   // roughconductor.cpp:258-297
        if (measure != ESolidAngle || Frame::cosTheta(b.wi) <= 0 || Frame::cosTheta(b.wo) <= 0) return RGB(0.0);
        Vec3 H = normalize(b.wo + b.wi);
        Microfacet distr(m);
        const Float D = distr.eval(H);
        if (D == 0) return RGB(0.0);
        const RGB F = fresnelConductorExact(dot(b.wi, H), rgb3(m.eta), rgb3(m.k)) * m.R;
        const Float G = distr.G(b.wi, b.wo, H);
        Float model = D * G / (4.0 * Frame::cosTheta(b.wi));
        return F * model;
    }
    case DR_BSDF_ROUGHPLASTIC:   // roughplastic.cpp:325-381
        if (measure != ESolidAngle || Frame::cosTheta(b.wi) <= 0 || Frame::cosTheta(b.wo) <= 0) return RGB(0.0);
        return RoughPlastic(m).eval(m, b.wi, b.wo);
    case DR_BSDF_PLASTIC: {   // plastic.cpp:240-277
        if (Frame::cosTheta(b.wo) <= 0 || Frame::cosTheta(b.wi) <= 0) return RGB(0.0);
        Plastic p(m);
        Float ct;
        Float Fi = fresnelDielectricExt(Frame::cosTheta(b.wi), ct, p.eta);
        if (measure == EDiscrete) {
            if (std::abs(dot(reflectZ(b.wi), b.wo) - 1) < DELTA_EPSILON) return p.specular * Fi;
        } else if (measure == ESolidAngle) {
            Float Fo = fresnelDielectricExt(Frame::cosTheta(b.wo), ct, p.eta);
            return p.diff() * (squareToCosineHemispherePdf(b.wo) * p.invEta2 * (1 - Fi) * (1 - Fo));
        }
        return RGB(0.0);
    }
    case DR_BSDF_ROUGHDIELECTRIC: {   // roughdielectric.cpp:270-348
        if (measure != ESolidAngle || Frame::cosTheta(b.wi) == 0) return RGB(0.0);
        const Float mEta = m.eta[0], mInvEta = 1 / mEta;
        const bool reflect = Frame::cosTheta(b.wi) * Frame::cosTheta(b.wo) > 0;
        Vec3 H;
        if (reflect) H = normalize(b.wo + b.wi);
        else { Float eta = Frame::cosTheta(b.wi) > 0 ? mEta : mInvEta; H = normalize(b.wi + b.wo * eta); }
        H = H * std::copysign(1.0, Frame::cosTheta(H));
        Microfacet distr(m);
        const Float D = distr.eval(H);
        if (D == 0) return RGB(0.0);
        Float cosThetaT;
        const Float F = fresnelDielectricExt(dot(b.wi, H), cosThetaT, mEta);
        const Float G = distr.G(b.wi, b.wo, H);
        if (reflect) {
            Float value = F * D * G / (4.0 * std::abs(Frame::cosTheta(b.wi)));
            return m.R * value;
        } else {
            Float eta = Frame::cosTheta(b.wi) > 0.0 ? mEta : mInvEta;
            Float sqrtDenom = dot(b.wi, H) + eta * dot(b.wo, H);
            Float value = ((1 - F) * D * G * eta * eta * dot(b.wi, H) * dot(b.wo, H)) / (Frame::cosTheta(b.wi) * sqrtDenom * sqrtDenom);
            Float factor = (b.mode == ERadiance) ? (Frame::cosTheta(b.wi) > 0 ? mInvEta : mEta) : 1.0;
            return m.T * std::abs(value * factor * factor);
        }
    }
    case DR_BSDF_DIELECTRIC: {   // dielectric.cpp:227-253
        if (measure != EDiscrete) return RGB(0.0);
        Float eta = m.eta[0], invEta = 1 / eta, cosThetaT;
        Float F = fresnelDielectricExt(Frame::cosTheta(b.wi), cosThetaT, eta);
        if (Frame::cosTheta(b.wi) * Frame::cosTheta(b.wo) >= 0) {
            if (std::abs(dot(reflectZ(b.wi), b.wo) - 1) > DELTA_EPSILON) return RGB(0.0);
            return m.R * F;
        } else {
            Float scale = -(cosThetaT < 0 ? invEta : eta);
            Vec3 refr(scale * b.wi.x, scale * b.wi.y, cosThetaT);
            if (std::abs(dot(refr, b.wo) - 1) > DELTA_EPSILON) return RGB(0.0);
            Float factor = (b.mode == ERadiance) ? (cosThetaT < 0 ? invEta : eta) : 1.0;
            return m.T * (factor * factor * (1 - F));
        }
    }
    }
    return RGB(0.0);
}

inline Float pdfNested(const Mat &m, const BSDFRecord &b, int measure) {
    switch (m.type) {
    case DR_BSDF_DIFFUSE:   // diffuse.cpp:119-126
        if (measure != ESolidAngle || Frame::cosTheta(b.wi) <= 0 || Frame::cosTheta(b.wo) <= 0) return 0.0;
        return squareToCosineHemispherePdf(b.wo);
    case DR_BSDF_CONDUCTOR:   // conductor.cpp:239-252
        if (measure != EDiscrete || Frame::cosTheta(b.wi) <= 0 || Frame::cosTheta(b.wo) <= 0 ||
            std::abs(dot(reflectZ(b.wi), b.wo) - 1) > DELTA_EPSILON) return 0.0;
        return 1.0;
    case DR_BSDF_ROUGHCONDUCTOR: {   // roughconductor.cpp:299-324
        if (measure != ESolidAngle || Frame::cosTheta(b.wi) <= 0 || Frame::cosTheta(b.wo) <= 0) return 0.0;
        Vec3 H = normalize(b.wo + b.wi);
        Microfacet distr(m);
        if (distr.sampleVis)
            return distr.eval(H) * distr.smithG1(b.wi, H) / (4.0 * Frame::cosTheta(b.wi));
        else
            return distr.pdf(b.wi, H) / (4 * absDot(b.wo, H));
    }
    case DR_BSDF_ROUGHPLASTIC:   // roughplastic.cpp:383-432
        if (measure != ESolidAngle || Frame::cosTheta(b.wi) <= 0 || Frame::cosTheta(b.wo) <= 0) return 0.0;
        return RoughPlastic(m).pdf(m, b.wi, b.wo);
    case DR_BSDF_PLASTIC: {   // plastic.cpp:279-307
        if (Frame::cosTheta(b.wo) <= 0 || Frame::cosTheta(b.wi) <= 0) return 0.0;
        Plastic p(m);
        Float ct;
        Float probSpecular = p.probSpecular(fresnelDielectricExt(Frame::cosTheta(b.wi), ct, p.eta));
        if (measure == EDiscrete) {
            if (std::abs(dot(reflectZ(b.wi), b.wo) - 1) < DELTA_EPSILON) return probSpecular;
        } else if (measure == ESolidAngle) return squareToCosineHemispherePdf(b.wo) * (1 - probSpecular);
        return 0.0;
    }
    case DR_BSDF_ROUGHDIELECTRIC: {   // roughdielectric.cpp:350-420 (both components enabled)
        if (measure != ESolidAngle) return 0.0;
        const Float mEta = m.eta[0], mInvEta = 1 / mEta;
        const bool reflect = Frame::cosTheta(b.wi) * Frame::cosTheta(b.wo) > 0;
        Vec3 H;
        Float dwh_dwo;
        if (reflect) {
            H = normalize(b.wo + b.wi);
            dwh_dwo = 1.0 / (4.0 * dot(b.wo, H));
        } else {
            Float eta = Frame::cosTheta(b.wi) > 0 ? mEta : mInvEta;
            H = normalize(b.wi + b.wo * eta);
            Float sqrtDenom = dot(b.wi, H) + eta * dot(b.wo, H);
            dwh_dwo = (eta * eta * dot(b.wo, H)) / (sqrtDenom * sqrtDenom);
        }
        H = H * std::copysign(1.0, Frame::cosTheta(H));
        Microfacet sampleDistr(m);
        if (!sampleDistr.sampleVis) sampleDistr.scaleAlpha(1.2f - 0.2f * std::sqrt(std::abs(Frame::cosTheta(b.wi))));
        Float prob = sampleDistr.pdf(b.wi * std::copysign(1.0, Frame::cosTheta(b.wi)), H);
        Float cosThetaT;
        Float F = fresnelDielectricExt(dot(b.wi, H), cosThetaT, mEta);
        prob *= reflect ? F : (1 - F);
        return std::abs(prob * dwh_dwo);
    }
    case DR_BSDF_DIELECTRIC: {   // dielectric.cpp:255-276
        if (measure != EDiscrete) return 0.0;
        Float eta = m.eta[0], invEta = 1 / eta, cosThetaT;
        Float F = fresnelDielectricExt(Frame::cosTheta(b.wi), cosThetaT, eta);
        if (Frame::cosTheta(b.wi) * Frame::cosTheta(b.wo) >= 0) {
            if (std::abs(dot(reflectZ(b.wi), b.wo) - 1) > DELTA_EPSILON) return 0.0;
            return F;
        } else {
            Float scale = -(cosThetaT < 0 ? invEta : eta);
            Vec3 refr(scale * b.wi.x, scale * b.wi.y, cosThetaT);
            if (std::abs(dot(refr, b.wo) - 1) > DELTA_EPSILON) return 0.0;
            return 1 - F;
        }
    }
    }
    return 0.0;
}

// `extra`: the number sample() draws from bRec.sampler (EUsesSampler BSDFs; roughdielectric.cpp:555)
inline RGB sampleNested(const Mat &m, BSDFRecord &b, Float &pdf, const Vec2 &sample, Float epsilon, Float extra) {
    switch (m.type) {
    case DR_BSDF_ROUGHPLASTIC: {   // roughplastic.cpp:434-491 (both components enabled)
        if (Frame::cosTheta(b.wi) <= 0) return RGB(0.0);
        RoughPlastic p(m);
        Microfacet distr(m);
        Vec2 s2(sample);
        const Float probSpecular = p.probSpecular(Frame::cosTheta(b.wi));
        bool choseSpecular = true;
        if (s2.y < probSpecular) s2.y /= probSpecular;
        else { s2.y = (s2.y - probSpecular) / (1 - probSpecular); choseSpecular = false; }
        if (choseSpecular) {
            Float unusedPdf;
            const Vec3 mm = distr.sample(b.wi, s2, unusedPdf, epsilon);
            b.wo = reflectM(b.wi, mm);
            b.sampledType = EGlossyReflection;
            if (Frame::cosTheta(b.wo) <= 0) return RGB(0.0);
        } else {
            b.sampledType = EDiffuseReflection;
            b.wo = squareToCosineHemisphere(s2);
        }
        b.eta = 1.0;
        pdf = (Frame::cosTheta(b.wi) <= 0 || Frame::cosTheta(b.wo) <= 0) ? 0.0 : p.pdf(m, b.wi, b.wo);
        if (pdf == 0) return RGB(0.0);
        return p.eval(m, b.wi, b.wo) * (1.0 / pdf);
    }
    case DR_BSDF_PLASTIC: {   // plastic.cpp:368-412 (both components enabled)
        if (Frame::cosTheta(b.wi) <= 0) return RGB(0.0);
        Plastic p(m);
        Float ct;
        Float Fi = fresnelDielectricExt(Frame::cosTheta(b.wi), ct, p.eta);
        Float probSpecular = p.probSpecular(Fi);
        b.eta = 1.0;
        if (sample.x < probSpecular) {
            b.sampledType = EDeltaReflection; b.wo = reflectZ(b.wi); pdf = probSpecular;
            return p.specular * (Fi / probSpecular);
        } else {
            b.sampledType = EDiffuseReflection;
            b.wo = squareToCosineHemisphere(Vec2((sample.x - probSpecular) / (1 - probSpecular), sample.y));
            Float Fo = fresnelDielectricExt(Frame::cosTheta(b.wo), ct, p.eta);
            pdf = (1 - probSpecular) * squareToCosineHemispherePdf(b.wo);
            return p.diff() * (p.invEta2 * (1 - Fi) * (1 - Fo) / (1 - probSpecular));
        }
    }
    case DR_BSDF_ROUGHDIELECTRIC: {   // roughdielectric.cpp:514-611 (both components enabled)
        const Float mEta = m.eta[0], mInvEta = 1 / mEta;
        Microfacet distr(m);
        Microfacet sampleDistr(distr);
        if (!distr.sampleVis) sampleDistr.scaleAlpha(1.2f - 0.2f * std::sqrt(std::abs(Frame::cosTheta(b.wi))));
        Float microfacetPDF;
        const Vec3 mm = sampleDistr.sample(b.wi * std::copysign(1.0, Frame::cosTheta(b.wi)), sample, microfacetPDF, epsilon);
        if (microfacetPDF == 0) return RGB(0.0);
        float temporaryPdf = (float) microfacetPDF;              // sic: single precision in the reference (:543)
        Float cosThetaT;
        Float F = fresnelDielectricExt(dot(b.wi, mm), cosThetaT, mEta);
        RGB weight(1.0);
        bool sampleReflection = true;
        if (extra > F) { sampleReflection = false; temporaryPdf *= 1 - F; }
        else temporaryPdf *= F;
        Float dwh_dwo;
        if (sampleReflection) {
            b.wo = reflectM(b.wi, mm);
            b.eta = 1.0; b.sampledType = EGlossyReflection;
            if (Frame::cosTheta(b.wi) * Frame::cosTheta(b.wo) <= 0) return RGB(0.0);
            weight = weight * m.R;
            dwh_dwo = 1.0 / (4.0 * dot(b.wo, mm));
        } else {
            if (cosThetaT == 0) return RGB(0.0);
            Float e = cosThetaT < 0 ? mInvEta : mEta;            // refract(): util.cpp:775-780 (eta inverted when cosThetaT < 0)
            b.wo = mm * (dot(b.wi, mm) * e + cosThetaT) - b.wi * e;
            b.eta = cosThetaT < 0 ? mEta : mInvEta;
            b.sampledType = EGlossyTransmission;
            if (Frame::cosTheta(b.wi) * Frame::cosTheta(b.wo) >= 0) return RGB(0.0);
            Float factor = (b.mode == ERadiance) ? (cosThetaT < 0 ? mInvEta : mEta) : 1.0;
            weight = weight * m.T * (factor * factor);
            Float sqrtDenom = dot(b.wi, mm) + b.eta * dot(b.wo, mm);
            dwh_dwo = (b.eta * b.eta * dot(b.wo, mm)) / (sqrtDenom * sqrtDenom);
        }
        if (distr.sampleVis) weight = weight * distr.smithG1(b.wo, mm);
        else weight = weight * std::abs(distr.eval(mm) * distr.G(b.wi, b.wo, mm) * dot(b.wi, mm) / (microfacetPDF * Frame::cosTheta(b.wi)));
        temporaryPdf *= std::abs(dwh_dwo);
        pdf = temporaryPdf;
        return weight;
    }
    case DR_BSDF_DIFFUSE:   // diffuse.cpp:139-149
        if (Frame::cosTheta(b.wi) <= 0) return RGB(0.0);
        b.wo = squareToCosineHemisphere(sample);
        b.eta = 1.0; b.sampledType = EDiffuseReflection;
        pdf = squareToCosineHemispherePdf(b.wo);
        return m.R;
    case DR_BSDF_CONDUCTOR:   // conductor.cpp:270-285
        if (Frame::cosTheta(b.wi) <= 0) return RGB(0.0);
        b.sampledType = EDeltaReflection; b.wo = reflectZ(b.wi); b.eta = 1.0; pdf = 1;
        return m.R * fresnelConductorExact(Frame::cosTheta(b.wi), rgb3(m.eta), rgb3(m.k));
    case DR_BSDF_ROUGHCONDUCTOR: {   // roughconductor.cpp:371-417
        if (Frame::cosTheta(b.wi) < 0) return RGB(0.0);
        Microfacet distr(m);
        Float temporaryPdf = 0;
        Vec3 mm = distr.sample(b.wi, sample, temporaryPdf, epsilon);
        if (temporaryPdf == 0) return RGB(0.0);
        b.wo = reflectM(b.wi, mm);
        b.eta = 1.0; b.sampledType = EGlossyReflection;
        if (Frame::cosTheta(b.wo) <= 0) return RGB(0.0);
        RGB F = fresnelConductorExact(dot(b.wi, mm), rgb3(m.eta), rgb3(m.k)) * m.R;
        Float weight;
        if (distr.sampleVis) weight = distr.smithG1(b.wo, mm);
        else weight = distr.eval(mm) * distr.G(b.wi, b.wo, mm) * dot(b.wi, mm) / (temporaryPdf * Frame::cosTheta(b.wi));
        if (weight > 0) {
            pdf = temporaryPdf / (4.0 * dot(b.wo, mm));
            return F * weight;
        }
        return RGB(0.0);
    }
    case DR_BSDF_DIELECTRIC: {   // dielectric.cpp:278-330 (both components enabled)
        Float eta = m.eta[0], invEta = 1 / eta, cosThetaT;
        Float F = fresnelDielectricExt(Frame::cosTheta(b.wi), cosThetaT, eta);
        if (sample.x <= F) {
            b.sampledType = EDeltaReflection; b.wo = reflectZ(b.wi); b.eta = 1.0; pdf = F;
            return m.R;
        } else {
            b.sampledType = EDeltaTransmission;
            Float scale = -(cosThetaT < 0 ? invEta : eta);
            b.wo = Vec3(scale * b.wi.x, scale * b.wi.y, cosThetaT);
            b.eta = cosThetaT < 0 ? eta : invEta;
            pdf = 1 - F;
            Float factor = (b.mode == ERadiance) ? (cosThetaT < 0 ? invEta : eta) : 1.0;
            return m.T * (factor * factor);
        }
    }
    }
    return RGB(0.0);
}

} // namespace detail

// ---------------------------------------------------------------- public: with the twosided adapter
inline RGB bsdfEval(const Mat &m, const BSDFRecord &bRec, int measure = ESolidAngle) {
    if (m.flags & DR_MAT_TWOSIDED) {   // twosided.cpp:107-127
        BSDFRecord b(bRec);
        if (Frame::cosTheta(b.wi) > 0) return detail::evalNested(m, b, measure);
        b.wi.z *= -1; b.wo.z *= -1;
        return detail::evalNested(m, b, measure);
    }
    return detail::evalNested(m, bRec, measure);
}
inline Float bsdfPdf(const Mat &m, const BSDFRecord &bRec, int measure = ESolidAngle) {
    if (m.flags & DR_MAT_TWOSIDED) {   // twosided.cpp:129-141
        BSDFRecord b(bRec);
        if (b.wi.z > 0) return detail::pdfNested(m, b, measure);
        b.wi.z *= -1; b.wo.z *= -1;
        return detail::pdfNested(m, b, measure);
    }
    return detail::pdfNested(m, bRec, measure);
}
inline RGB bsdfSample(const Mat &m, BSDFRecord &bRec, Float &pdf, const Vec2 &sample, Float epsilon, Float extra = 0.5) {
    pdf = 0;
    if (m.flags & DR_MAT_TWOSIDED) {   // twosided.cpp:166-186
        bool flipped = false;
        if (Frame::cosTheta(bRec.wi) < 0) { bRec.wi.z *= -1; flipped = true; }
        RGB result = detail::sampleNested(m, bRec, pdf, sample, epsilon, extra);
        if (flipped) {
            bRec.wi.z *= -1;
            if (!result.isZero() && pdf != 0) bRec.wo.z *= -1;
        }
        return result;
    }
    return detail::sampleNested(m, bRec, pdf, sample, epsilon, extra);
}

} // namespace orc
