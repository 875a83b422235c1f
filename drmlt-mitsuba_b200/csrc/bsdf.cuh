// bsdf.cuh -- device BSDF models in the local shading frame (z = shading normal).
//
// Behavioural parity targets: src/bsdfs/diffuse.cpp:109-150, dielectric.cpp:227-330,
// conductor.cpp:223-285, roughconductor.cpp:258-417 with src/bsdfs/microfacet.h (Beckmann/GGX,
// isotropic, sampleAll and sampleVisible), wrapped by twosided.cpp:107-186 when DR_MAT_TWOSIDED.
// Branching is on the material type (a small enum) -- no virtual dispatch.
#pragma once
#include "scene.h"

enum { MODE_RADIANCE = 0, MODE_IMPORTANCE = 1 };
enum { MEAS_INVALID = 0, MEAS_SOLID_ANGLE = 1, MEAS_AREA = 3, MEAS_DISCRETE = 4 };
enum { BT_DIFFUSE_R = 1, BT_GLOSSY_R = 2, BT_DELTA_R = 4, BT_DELTA_T = 8, BT_SMOOTH = 3, BT_DELTA = 12 };

struct Mat {   // material fetched into registers
    int type; uint32_t flags;
    float3 refl, trans, eta, k;
    float alpha;
};

DR_D Mat load_material(const DevScene &sc, int id) {
    const float4 *p = reinterpret_cast<const float4 *>(sc.materials + id);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    Mat m;
    m.type = __float_as_int(a.x); m.flags = (uint32_t) __float_as_int(a.y);
    m.refl = f3(a.z, a.w, b.x); m.trans = f3(b.y, b.z, b.w);
    m.eta = f3(c.x, c.y, c.z); m.k = f3(c.w, d.x, d.y);
    m.alpha = d.z;
    return m;
}
DR_D bool mat_has_smooth(int type) { return type == DR_BSDF_DIFFUSE || type == DR_BSDF_ROUGHCONDUCTOR; }
DR_D bool mat_non_symmetric(int type) { return type == DR_BSDF_DIELECTRIC; }
DR_D bool mat_transmissive_or_backside(const Mat &m) { return m.type == DR_BSDF_DIELECTRIC || (m.flags & DR_MAT_TWOSIDED); }

// ---- warps (src/libcore/warp.cpp:44-100)
DR_D float2 square_to_disk_concentric(float sx, float sy) {
    float r1 = 2.0f * sx - 1.0f, r2 = 2.0f * sy - 1.0f;
    float phi, r;
    if (r1 == 0.f && r2 == 0.f) { r = phi = 0.f; }
    else if (r1 * r1 > r2 * r2) { r = r1; phi = (DR_PI / 4.0f) * (r2 / r1); }
    else { r = r2; phi = (DR_PI / 2.0f) - (r1 / r2) * (DR_PI / 4.0f); }
    float s, c;
    sincosf(phi, &s, &c);
    return make_float2(r * c, r * s);
}
DR_D float3 square_to_cosine_hemisphere(float sx, float sy) {
    float2 p = square_to_disk_concentric(sx, sy);
    float z = safe_sqrtf(1.0f - p.x * p.x - p.y * p.y);
    if (z == 0.f) z = 1e-10f;
    return f3(p.x, p.y, z);
}

// ---- Fresnel (src/libcore/util.cpp:659-693, 765-789)
DR_D float fresnel_dielectric_ext(float cosThetaI_, float &cosThetaT_, float eta) {
    if (eta == 1.f) { cosThetaT_ = -cosThetaI_; return 0.0f; }
    float scale = (cosThetaI_ > 0.f) ? 1.0f / eta : eta;
    float cosThetaTSqr = 1.0f - (1.0f - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) { cosThetaT_ = 0.0f; return 1.0f; }
    float cosThetaI = fabsf(cosThetaI_), cosThetaT = sqrtf(cosThetaTSqr);
    float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0.f) ? -cosThetaT : cosThetaT;
    return 0.5f * (Rs * Rs + Rp * Rp);
}
DR_D float fresnel_conductor_1(float cosThetaI, float eta, float k) {
    float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1.f - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
    float temp1 = eta * eta - k * k - sinThetaI2;
    float a2pb2 = safe_sqrtf(temp1 * temp1 + 4.f * k * k * eta * eta);
    float a = safe_sqrtf(0.5f * (a2pb2 + temp1));
    float term1 = a2pb2 + cosThetaI2, term2 = 2.f * a * cosThetaI;
    float Rs2 = (term1 - term2) / (term1 + term2);
    float term3 = a2pb2 * cosThetaI2 + sinThetaI4, term4 = term2 * sinThetaI2;
    float Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
    return 0.5f * (Rp2 + Rs2);
}
DR_D float3 fresnel_conductor(float cosThetaI, float3 eta, float3 k) {
    return f3(fresnel_conductor_1(cosThetaI, eta.x, k.x), fresnel_conductor_1(cosThetaI, eta.y, k.y),
              fresnel_conductor_1(cosThetaI, eta.z, k.z));
}

// ---- isotropic microfacet distribution (src/bsdfs/microfacet.h)
struct Microfacet {
    bool ggx, visible;
    float alpha;
    DR_D Microfacet(const Mat &m) : ggx(m.flags & DR_MAT_GGX), visible(m.flags & DR_MAT_SAMPLE_VISIBLE), alpha(fmaxf(m.alpha, 1e-4f)) {}
    DR_D float eval(float3 m) const {
        if (m.z <= 0.f) return 0.0f;
        float cosTheta2 = m.z * m.z;
        float e = ((m.x * m.x) / (alpha * alpha) + (m.y * m.y) / (alpha * alpha)) / cosTheta2;
        float result;
        if (!ggx) result = expf(-e) / (DR_PI * alpha * alpha * cosTheta2 * cosTheta2);
        else { float root = (1.0f + e) * cosTheta2; result = 1.0f / (DR_PI * alpha * alpha * root * root); }
        if (result * m.z < 1e-20f) result = 0.f;
        return result;
    }
    DR_D float smithG1(float3 v, float3 m) const {
        if (dot(v, m) * v.z <= 0.f) return 0.0f;
        float temp = 1.f - v.z * v.z;
        float tanTheta = temp <= 0.f ? 0.f : fabsf(sqrtf(temp) / v.z);
        if (tanTheta == 0.0f) return 1.0f;
        if (!ggx) {
            float a = 1.0f / (alpha * tanTheta);
            if (a >= 1.6f) return 1.0f;
            float aSqr = a * a;
            return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
        } else {
            float root = alpha * tanTheta;
            return 2.0f / (1.0f + sqrtf(1.0f + root * root));
        }
    }
    DR_D float G(float3 wi, float3 wo, float3 m) const { return smithG1(wi, m) * smithG1(wo, m); }
    DR_D float pdf_all(float3 m) const { return eval(m) * m.z; }
    DR_D float pdf_visible(float3 wi, float3 m) const {
        if (wi.z == 0.f) return 0.0f;
        return smithG1(wi, m) * absdot(wi, m) * eval(m) / fabsf(wi.z);
    }
    DR_D float pdf(float3 wi, float3 m) const { return visible ? pdf_visible(wi, m) : pdf_all(m); }
    DR_D float3 sample_all(float sx, float sy, float &pdf) const {
        float sinPhiM, cosPhiM;
        sincosf(2.0f * DR_PI * sy, &sinPhiM, &cosPhiM);
        float alphaSqr = alpha * alpha, cosThetaM;
        if (!ggx) {
            float tanThetaMSqr = alphaSqr * -logf(1.0f - sx);
            cosThetaM = 1.0f / sqrtf(1.0f + tanThetaMSqr);
            pdf = (1.0f - sx) / (DR_PI * alpha * alpha * cosThetaM * cosThetaM * cosThetaM);
        } else {
            float tanThetaMSqr = alphaSqr * sx / (1.0f - sx);
            cosThetaM = 1.0f / sqrtf(1.0f + tanThetaMSqr);
            float temp = 1.f + tanThetaMSqr / alphaSqr;
            pdf = DR_INV_PI / (alpha * alpha * cosThetaM * cosThetaM * cosThetaM * temp * temp);
        }
        if (pdf < 1e-20f) pdf = 0.f;
        float sinThetaM = sqrtf(fmaxf(0.f, 1.f - cosThetaM * cosThetaM));
        return f3(sinThetaM * cosPhiM, sinThetaM * sinPhiM, cosThetaM);
    }
    static DR_D float erfinv_giles(float x) {   // src/libcore/math.cpp:25-52
        float w = -logf((1.0f - x) * (1.0f + x)), p;
        if (w < 5.0f) {
            w = w - 2.5f;
            p = 2.81022636e-08f; p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w; p = -4.39150654e-06f + p * w;
            p = 0.00021858087f + p * w; p = -0.00125372503f + p * w; p = -0.00417768164f + p * w;
            p = 0.246640727f + p * w; p = 1.50140941f + p * w;
        } else {
            w = sqrtf(w) - 3.0f;
            p = -0.000200214257f; p = 0.000100950558f + p * w; p = 0.00134934322f + p * w; p = -0.00367342844f + p * w;
            p = 0.00573950773f + p * w; p = -0.0076224613f + p * w; p = 0.00943887047f + p * w;
            p = 1.00167406f + p * w; p = 2.83297682f + p * w;
        }
        return p * x;
    }
    static DR_D float erf_as(float x) {   // math.cpp:54-70 (A&S 7.1.26)
        float sign = signbit(x) ? -1.f : 1.f;
        x = fabsf(x);
        float t = 1.0f / (1.0f + 0.3275911f * x);
        float y = 1.0f - (((((1.061405429f * t + -1.453152027f) * t) + 1.421413741f) * t + -0.284496736f) * t + 0.254829592f) * t * expf(-x * x);
        return sign * y;
    }
    DR_D float2 sample_visible11(float thetaI, float sx, float sy, float epsilon) const {
        const float SQRT_PI_INV = 0.5641895835477563f;
        float2 slope;
        if (!ggx) {
            if (thetaI < 1e-4f) {
                float r = sqrtf(-logf(1.0f - sx)), s, c;
                sincosf(2.f * DR_PI * sy, &s, &c);
                return make_float2(r * c, r * s);
            }
            float tanThetaI = tanf(thetaI), cotThetaI = 1.f / tanThetaI;
            float a = -1.f, c = erf_as(cotThetaI);
            float sample_x = fmaxf(sx, 1e-6f);
            float fit = 1.f + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
            float b = c - (1.f + c) * powf(1.f - sample_x, fit);
            float normalization = 1.f / (1.f + c + SQRT_PI_INV * tanThetaI * expf(-cotThetaI * cotThetaI));
            int it = 0;
            while (++it < 10) {
                if (!(b >= a && b <= c)) b = 0.5f * (a + c);
                float invErf = erfinv_giles(b);
                float value = normalization * (1.f + b + SQRT_PI_INV * tanThetaI * expf(-invErf * invErf)) - sample_x;
                float derivative = normalization * (1.f - invErf * tanThetaI);
                if (fabsf(value) < 1e-5f) break;
                if (value > 0.f) c = b; else a = b;
                b -= value / derivative;
            }
            slope.x = erfinv_giles(b);
            slope.y = erfinv_giles(2.0f * fmaxf(sy, 1e-6f) - 1.0f);
        } else {
            if (thetaI < 1e-4f) {
                float r = safe_sqrtf(sx / (1.f - sx)), s, c;
                sincosf(2.f * DR_PI * sy, &s, &c);
                return make_float2(r * c, r * s);
            }
            float tanThetaI = tanf(thetaI);
            float a = 1.f / tanThetaI;
            float G1 = 2.0f / (1.0f + safe_sqrtf(1.0f + 1.0f / (a * a)));
            float A = 2.0f * sx / G1 - 1.0f;
            if (fabsf(A) == 1.f) A -= (signbit(A) ? -1.f : 1.f) * epsilon;
            float tmp = 1.0f / (A * A - 1.0f);
            float B = tanThetaI;
            float D = safe_sqrtf(B * B * tmp * tmp - (A * A - B * B) * tmp);
            float slope_x_1 = B * tmp - D, slope_x_2 = B * tmp + D;
            slope.x = (A < 0.0f || slope_x_2 > 1.0f / tanThetaI) ? slope_x_1 : slope_x_2;
            float S;
            if (sy > 0.5f) { S = 1.0f; sy = 2.0f * (sy - 0.5f); }
            else { S = -1.0f; sy = 2.0f * (0.5f - sy); }
            float z = (sy * (sy * (sy * (-0.365728915865723f) + 0.790235037209296f) - 0.424965825137544f) + 0.000152998850436920f) /
                      (sy * (sy * (sy * (sy * 0.169507819808272f - 0.397203533833404f) - 0.232500544458471f) + 1.f) - 0.539825872510702f);
            slope.y = S * z * sqrtf(1.0f + slope.x * slope.x);
        }
        return slope;
    }
    DR_D float3 sample_visible(float3 _wi, float sx, float sy, float epsilon) const {
        float3 wi = normalize(f3(alpha * _wi.x, alpha * _wi.y, _wi.z));
        float theta = 0.f, phi = 0.f;
        if (wi.z < 0.99999f) { theta = acosf(wi.z); phi = atan2f(wi.y, wi.x); }
        float sinPhi, cosPhi;
        sincosf(phi, &sinPhi, &cosPhi);
        float2 slope = sample_visible11(theta, sx, sy, epsilon);
        slope = make_float2(cosPhi * slope.x - sinPhi * slope.y, sinPhi * slope.x + cosPhi * slope.y);
        slope.x *= alpha; slope.y *= alpha;
        float normalization = 1.0f / sqrtf(slope.x * slope.x + slope.y * slope.y + 1.0f);
        return f3(-slope.x * normalization, -slope.y * normalization, normalization);
    }
    DR_D float3 sample(float3 wi, float sx, float sy, float &pdf, float epsilon) const {
        if (visible) { float3 m = sample_visible(wi, sx, sy, epsilon); pdf = pdf_visible(wi, m); return m; }
        return sample_all(sx, sy, pdf);
    }
};

DR_D float3 reflect_z(float3 wi) { return f3(-wi.x, -wi.y, wi.z); }

// ---- nested (one-sided) evaluation
DR_D float3 bsdf_eval_nested(const Mat &m, float3 wi, float3 wo, int mode, int measure) {
    switch (m.type) {
    case DR_BSDF_DIFFUSE:
        if (measure != MEAS_SOLID_ANGLE || wi.z <= 0.f || wo.z <= 0.f) return f3(0.f);
        return m.refl * (DR_INV_PI * wo.z);
    case DR_BSDF_CONDUCTOR:
        if (measure != MEAS_DISCRETE || wi.z <= 0.f || wo.z <= 0.f || fabsf(dot(reflect_z(wi), wo) - 1.f) > DR_DELTA_EPS) return f3(0.f);
        return m.refl * fresnel_conductor(wi.z, m.eta, m.k);
    case DR_BSDF_ROUGHCONDUCTOR: {
        if (measure != MEAS_SOLID_ANGLE || wi.z <= 0.f || wo.z <= 0.f) return f3(0.f);
        float3 H = normalize(wo + wi);
        Microfacet distr(m);
        const float D = distr.eval(H);
        if (D == 0.f) return f3(0.f);
        const float3 F = fresnel_conductor(dot(wi, H), m.eta, m.k) * m.refl;
        const float G = distr.G(wi, wo, H);
        return F * (D * G / (4.0f * wi.z));
    }
    case DR_BSDF_DIELECTRIC: {
        if (measure != MEAS_DISCRETE) return f3(0.f);
        float eta = m.eta.x, invEta = 1.f / eta, cosThetaT;
        float F = fresnel_dielectric_ext(wi.z, cosThetaT, eta);
        if (wi.z * wo.z >= 0.f) {
            if (fabsf(dot(reflect_z(wi), wo) - 1.f) > DR_DELTA_EPS) return f3(0.f);
            return m.refl * F;
        } else {
            float scale = -(cosThetaT < 0.f ? invEta : eta);
            float3 refr = f3(scale * wi.x, scale * wi.y, cosThetaT);
            if (fabsf(dot(refr, wo) - 1.f) > DR_DELTA_EPS) return f3(0.f);
            float factor = (mode == MODE_RADIANCE) ? (cosThetaT < 0.f ? invEta : eta) : 1.0f;
            return m.trans * (factor * factor * (1.f - F));
        }
    }
    }
    return f3(0.f);
}

DR_D float bsdf_pdf_nested(const Mat &m, float3 wi, float3 wo, int measure) {
    switch (m.type) {
    case DR_BSDF_DIFFUSE:
        if (measure != MEAS_SOLID_ANGLE || wi.z <= 0.f || wo.z <= 0.f) return 0.f;
        return DR_INV_PI * wo.z;
    case DR_BSDF_CONDUCTOR:
        if (measure != MEAS_DISCRETE || wi.z <= 0.f || wo.z <= 0.f || fabsf(dot(reflect_z(wi), wo) - 1.f) > DR_DELTA_EPS) return 0.f;
        return 1.0f;
    case DR_BSDF_ROUGHCONDUCTOR: {
        if (measure != MEAS_SOLID_ANGLE || wi.z <= 0.f || wo.z <= 0.f) return 0.f;
        float3 H = normalize(wo + wi);
        Microfacet distr(m);
        if (distr.visible) return distr.eval(H) * distr.smithG1(wi, H) / (4.0f * wi.z);
        return distr.pdf(wi, H) / (4.f * absdot(wo, H));
    }
    case DR_BSDF_DIELECTRIC: {
        if (measure != MEAS_DISCRETE) return 0.f;
        float eta = m.eta.x, invEta = 1.f / eta, cosThetaT;
        float F = fresnel_dielectric_ext(wi.z, cosThetaT, eta);
        if (wi.z * wo.z >= 0.f) {
            if (fabsf(dot(reflect_z(wi), wo) - 1.f) > DR_DELTA_EPS) return 0.f;
            return F;
        } else {
            float scale = -(cosThetaT < 0.f ? invEta : eta);
            float3 refr = f3(scale * wi.x, scale * wi.y, cosThetaT);
            if (fabsf(dot(refr, wo) - 1.f) > DR_DELTA_EPS) return 0.f;
            return 1.f - F;
        }
    }
    }
    return 0.f;
}

struct BsdfSample { float3 wo; float3 weight; float pdf; int sampledType; float eta; };

DR_D void bsdf_sample_nested(const Mat &m, float3 wi, int mode, float sx, float sy, float epsilon, BsdfSample &r) {
    r.weight = f3(0.f); r.pdf = 0.f; r.sampledType = 0; r.eta = 1.f; r.wo = f3(0.f, 0.f, 1.f);
    switch (m.type) {
    case DR_BSDF_DIFFUSE:
        if (wi.z <= 0.f) return;
        r.wo = square_to_cosine_hemisphere(sx, sy);
        r.sampledType = BT_DIFFUSE_R; r.pdf = DR_INV_PI * r.wo.z; r.weight = m.refl;
        return;
    case DR_BSDF_CONDUCTOR:
        if (wi.z <= 0.f) return;
        r.sampledType = BT_DELTA_R; r.wo = reflect_z(wi); r.pdf = 1.f;
        r.weight = m.refl * fresnel_conductor(wi.z, m.eta, m.k);
        return;
    case DR_BSDF_ROUGHCONDUCTOR: {
        if (wi.z < 0.f) return;
        Microfacet distr(m);
        float tpdf = 0.f;
        float3 mm = distr.sample(wi, sx, sy, tpdf, epsilon);
        if (tpdf == 0.f) return;
        r.wo = mm * (2.f * dot(wi, mm)) - wi;
        r.sampledType = BT_GLOSSY_R;
        if (r.wo.z <= 0.f) return;
        float3 F = fresnel_conductor(dot(wi, mm), m.eta, m.k) * m.refl;
        float weight;
        if (distr.visible) weight = distr.smithG1(r.wo, mm);
        else weight = distr.eval(mm) * distr.G(wi, r.wo, mm) * dot(wi, mm) / (tpdf * wi.z);
        if (weight > 0.f) { r.pdf = tpdf / (4.0f * dot(r.wo, mm)); r.weight = F * weight; }
        return;
    }
    case DR_BSDF_DIELECTRIC: {
        float eta = m.eta.x, invEta = 1.f / eta, cosThetaT;
        float F = fresnel_dielectric_ext(wi.z, cosThetaT, eta);
        if (sx <= F) {
            r.sampledType = BT_DELTA_R; r.wo = reflect_z(wi); r.pdf = F; r.weight = m.refl;
        } else {
            r.sampledType = BT_DELTA_T;
            float scale = -(cosThetaT < 0.f ? invEta : eta);
            r.wo = f3(scale * wi.x, scale * wi.y, cosThetaT);
            r.eta = cosThetaT < 0.f ? eta : invEta;
            r.pdf = 1.f - F;
            float factor = (mode == MODE_RADIANCE) ? (cosThetaT < 0.f ? invEta : eta) : 1.0f;
            r.weight = m.trans * (factor * factor);
        }
        return;
    }
    }
}

// ---- public: twosided adapter (twosided.cpp:107-186)
DR_D float3 bsdf_eval(const Mat &m, float3 wi, float3 wo, int mode, int measure) {
    if ((m.flags & DR_MAT_TWOSIDED) && !(wi.z > 0.f)) { wi.z = -wi.z; wo.z = -wo.z; }
    return bsdf_eval_nested(m, wi, wo, mode, measure);
}
DR_D float bsdf_pdf(const Mat &m, float3 wi, float3 wo, int measure) {
    if ((m.flags & DR_MAT_TWOSIDED) && !(wi.z > 0.f)) { wi.z = -wi.z; wo.z = -wo.z; }
    return bsdf_pdf_nested(m, wi, wo, measure);
}
DR_D void bsdf_sample(const Mat &m, float3 wi, int mode, float sx, float sy, float epsilon, BsdfSample &r) {
    bool flipped = false;
    if ((m.flags & DR_MAT_TWOSIDED) && wi.z < 0.f) { wi.z = -wi.z; flipped = true; }
    bsdf_sample_nested(m, wi, mode, sx, sy, epsilon, r);
    if (flipped && !is_zero(r.weight) && r.pdf != 0.f) r.wo.z = -r.wo.z;
}
