// bsdf.cuh -- device BSDF models in the local shading frame (z = shading normal).
//
// Behavioural parity targets: src/bsdfs/diffuse.cpp:109-150, dielectric.cpp:227-330,
// conductor.cpp:223-285, roughconductor.cpp:258-417, roughdielectric.cpp:270-611, plastic.cpp:240-420, roughplastic.cpp:325-470 (+ rtrans.h)
// with src/bsdfs/microfacet.h (Beckmann/GGX,
// isotropic, sampleAll and sampleVisible), wrapped by twosided.cpp:107-186 when DR_MAT_TWOSIDED.
// Branching is on the material type (a small enum) -- no virtual dispatch.
#pragma once
#include "scene.h"
#include "real.cuh"

enum { MODE_RADIANCE = 0, MODE_IMPORTANCE = 1 };
enum { MEAS_INVALID = 0, MEAS_SOLID_ANGLE = 1, MEAS_AREA = 3, MEAS_DISCRETE = 4 };
enum { BT_DIFFUSE_R = 1, BT_GLOSSY_R = 2, BT_DELTA_R = 4, BT_DELTA_T = 8, BT_GLOSSY_T = 16, BT_SMOOTH = 19, BT_DELTA = 12 };

struct Mat {   // material fetched into registers
    int type; uint32_t flags;
    R3 refl, trans, eta, k;
    Real alpha;
    const double *table;     // roughplastic: its rough-transmittance table
};

DR_D Mat load_material(const DevScene &sc, int id) {
    const float4 *p = reinterpret_cast<const float4 *>(sc.materials + id);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    Mat m;
    m.type = __float_as_int(a.x); m.flags = (uint32_t) __float_as_int(a.y);
    m.refl = r3(a.z, a.w, b.x); m.trans = r3(b.y, b.z, b.w);
    m.eta = r3(c.x, c.y, c.z); m.k = r3(c.w, d.x, d.y);
    m.alpha = d.z;
    m.table = m.type == DR_BSDF_ROUGHPLASTIC ? sc.roughTables + (size_t) __float_as_uint(d.w) * DR_ROUGH_TABLE_DOUBLES : nullptr;
    return m;
}
// Bitmap texture lookup without ray differentials: BitmapTexture::eval(uv) (bitmap.cpp:432-455) -> MIPMap::evalBilinear(0, uv) /
// evalBox(0, uv) with evalTexel's boundary conditions (mipmap.h:503-596), after Texture2D::eval's scale and offset (texture.cpp:112-121).
DR_D bool tex_wrap(uint32_t mode, int &x, int size, Real &constant) {      // false: the texel is the constant (zero / one)
    if (x >= 0 && x < size) return true;
    switch (mode) {
        case DR_WRAP_REPEAT: { const int r = x % size; x = r < 0 ? r + size : r; return true; }
        case DR_WRAP_CLAMP: x = min(max(x, 0), size - 1); return true;
        case DR_WRAP_MIRROR: { const int r = x % (2 * size); x = r < 0 ? r + 2 * size : r; if (x >= size) x = 2 * size - x - 1; return true; }
        case DR_WRAP_ZERO: constant = 0.; return false;
        default: constant = 1.; return false;
    }
}
DR_D R3 tex_texel(const DevScene &sc, const DevTexture &t, int x, int y) {
    Real c = 0.;
    if (!tex_wrap(t.wrapU, x, (int) t.w, c)) return r3(c);
    if (!tex_wrap(t.wrapV, y, (int) t.h, c)) return r3(c);
    const float4 v = __ldg(sc.texels + t.first + (size_t) y * t.w + (size_t) x);
    return r3(v.x, v.y, v.z);
}
static __device__ __noinline__ R3 tex_eval(const DevScene &sc, uint32_t id, R2 uvIn) {
    const DevTexture t = sc.textures[id];
    const Real ux = uvIn.x * t.scaleU + t.offU, uy = uvIn.y * t.scaleV + t.offV;
    if (t.nearest) return tex_texel(sc, t, (int) floor(ux * (Real) t.w), (int) floor(uy * (Real) t.h));
    if (!isfinite(ux) || !isfinite(uy)) return r3(0.);
    const Real u = ux * (Real) t.w - 0.5f, v = uy * (Real) t.h - 0.5f;
    const int xPos = (int) floor(u), yPos = (int) floor(v);
    const Real dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
    return tex_texel(sc, t, xPos, yPos) * dx2 * dy2 + tex_texel(sc, t, xPos, yPos + 1) * dx2 * dy1
         + tex_texel(sc, t, xPos + 1, yPos) * dx1 * dy2 + tex_texel(sc, t, xPos + 1, yPos + 1) * dx1 * dy1;
}
// the material of a surface vertex: textured colour parameters (DR_MAT_TEX_*) are looked up at the vertex' uv
DR_D Mat load_material(const DevScene &sc, int id, R2 uv) {
    Mat m = load_material(sc, id);
    if (m.flags >> 8) {
        const uint32_t tr = (m.flags >> 8) & 0xfffu, tt = m.flags >> 20;
        if (tr) m.refl = tex_eval(sc, tr - 1u, uv);
        if (tt) m.trans = tex_eval(sc, tt - 1u, uv);
    }
    return m;
}
DR_D bool mat_has_smooth(int type) { return type == DR_BSDF_DIFFUSE || type == DR_BSDF_ROUGHCONDUCTOR || type == DR_BSDF_ROUGHDIELECTRIC || type == DR_BSDF_PLASTIC || type == DR_BSDF_ROUGHPLASTIC; }
DR_D bool mat_non_symmetric(int type) { return type == DR_BSDF_DIELECTRIC || type == DR_BSDF_ROUGHDIELECTRIC; }
DR_D bool mat_transmissive_or_backside(const Mat &m) { return m.type == DR_BSDF_DIELECTRIC || m.type == DR_BSDF_ROUGHDIELECTRIC || (m.flags & DR_MAT_TWOSIDED); }
// BSDF::EUsesSampler: sample() draws one more number from the vertex's sampler (roughdielectric.cpp:464, 555)
DR_D bool mat_uses_sampler(int type) { return type == DR_BSDF_ROUGHDIELECTRIC; }

// ---- warps (src/libcore/warp.cpp:44-100)
DR_D R2 square_to_disk_concentric(Real sx, Real sy) {
    Real ra = 2.0 * sx - 1.0, rb = 2.0 * sy - 1.0;
    Real phi, r;
    if (ra == 0. && rb == 0.) { r = phi = 0.; }
    else if (ra * ra > rb * rb) { r = ra; phi = (R_PI / 4.0) * (rb / ra); }
    else { r = rb; phi = (R_PI / 2.0) - (ra / rb) * (R_PI / 4.0); }
    Real s, c;
    sincos(phi, &s, &c);
    return r2(r * c, r * s);
}
DR_D R3 square_to_cosine_hemisphere(Real sx, Real sy) {
    R2 p = square_to_disk_concentric(sx, sy);
    Real z = safe_sqrt(1.0 - p.x * p.x - p.y * p.y);
    if (z == 0.) z = 1e-10f;
    return r3(p.x, p.y, z);
}

// ---- Fresnel (src/libcore/util.cpp:659-693, 765-789)
DR_D Real fresnel_dielectric_ext(Real cosThetaI_, Real &cosThetaT_, Real eta) {
    if (eta == 1.) { cosThetaT_ = -cosThetaI_; return 0.0; }
    Real scale = (cosThetaI_ > 0.) ? 1.0 / eta : eta;
    Real cosThetaTSqr = 1.0 - (1.0 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0) { cosThetaT_ = 0.0; return 1.0; }
    Real cosThetaI = fabs(cosThetaI_), cosThetaT = sqrt(cosThetaTSqr);
    Real Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    Real Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0.) ? -cosThetaT : cosThetaT;
    return 0.5 * (Rs * Rs + Rp * Rp);
}
DR_D Real fresnel_conductor_1(Real cosThetaI, Real eta, Real k) {
    Real cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1. - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
    Real temp1 = eta * eta - k * k - sinThetaI2;
    Real a2pb2 = safe_sqrt(temp1 * temp1 + 4. * k * k * eta * eta);
    Real a = safe_sqrt(0.5 * (a2pb2 + temp1));
    Real term1 = a2pb2 + cosThetaI2, term2 = 2. * a * cosThetaI;
    Real Rs2 = (term1 - term2) / (term1 + term2);
    Real term3 = a2pb2 * cosThetaI2 + sinThetaI4, term4 = term2 * sinThetaI2;
    Real Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
    return 0.5 * (Rp2 + Rs2);
}
DR_D R3 fresnel_conductor(Real cosThetaI, R3 eta, R3 k) {
    return r3(fresnel_conductor_1(cosThetaI, eta.x, k.x), fresnel_conductor_1(cosThetaI, eta.y, k.y),
              fresnel_conductor_1(cosThetaI, eta.z, k.z));
}

// ---- isotropic microfacet distribution (src/bsdfs/microfacet.h)
struct Microfacet {
    bool ggx, visible;
    Real alpha;
    DR_D Microfacet(const Mat &m) : ggx(m.flags & DR_MAT_GGX), visible(m.flags & DR_MAT_SAMPLE_VISIBLE), alpha(fmax(average3((Real) m.alpha), (Real) 1e-4f)) {}
    // the reference averages the alpha texture's Spectrum with a FLOAT third (spectrum.h:481-486, roughconductor.cpp:273)
    DR_D static Real average3(Real a) { Real r = 0.; r += a; r += a; r += a; return r * (Real) (1.0f / 3); }
    DR_D void scale_alpha(Real v) { alpha *= v; }      // microfacet.h:178-183
    DR_D Real eval(R3 m) const {
        if (m.z <= 0.) return 0.0;
        Real cosTheta2 = m.z * m.z;
        Real e = ((m.x * m.x) / (alpha * alpha) + (m.y * m.y) / (alpha * alpha)) / cosTheta2;
        Real result;
        if (!ggx) result = exp(-e) / (R_PI * alpha * alpha * cosTheta2 * cosTheta2);
        else { Real root = (1.0 + e) * cosTheta2; result = 1.0 / (R_PI * alpha * alpha * root * root); }
        if (result * m.z < 1e-20f) result = 0.;
        return result;
    }
    DR_D Real smithG1(R3 v, R3 m) const {
        if (dot(v, m) * v.z <= 0.) return 0.0;
        Real temp = 1. - v.z * v.z;
        Real tanTheta = temp <= 0. ? 0. : fabs(sqrt(temp) / v.z);
        if (tanTheta == 0.0) return 1.0;
        if (!ggx) {
            Real a = 1.0 / (alpha * tanTheta);
            if (a >= 1.6f) return 1.0;
            Real aSqr = a * a;
            return (3.535f * a + 2.181f * aSqr) / (1.0 + 2.276f * a + 2.577f * aSqr);
        } else {
            Real root = alpha * tanTheta;
            return 2.0 / (1.0 + sqrt(1.0 + root * root));
        }
    }
    DR_D Real G(R3 wi, R3 wo, R3 m) const { return smithG1(wi, m) * smithG1(wo, m); }
    DR_D Real pdf_all(R3 m) const { return eval(m) * m.z; }
    DR_D Real pdf_visible(R3 wi, R3 m) const {
        if (wi.z == 0.) return 0.0;
        return smithG1(wi, m) * absdot(wi, m) * eval(m) / fabs(wi.z);
    }
    DR_D Real pdf(R3 wi, R3 m) const { return visible ? pdf_visible(wi, m) : pdf_all(m); }
    DR_D R3 sample_all(Real sx, Real sy, Real &pdf) const {
        Real sinPhiM, cosPhiM;
        sincos(2.0 * R_PI * sy, &sinPhiM, &cosPhiM);
        Real alphaSqr = alpha * alpha, cosThetaM;
        if (!ggx) {
            Real tanThetaMSqr = alphaSqr * -log(1.0 - sx);
            cosThetaM = 1.0 / sqrt(1.0 + tanThetaMSqr);
            pdf = (1.0 - sx) / (R_PI * alpha * alpha * cosThetaM * cosThetaM * cosThetaM);
        } else {
            Real tanThetaMSqr = alphaSqr * sx / (1.0 - sx);
            cosThetaM = 1.0 / sqrt(1.0 + tanThetaMSqr);
            Real temp = 1. + tanThetaMSqr / alphaSqr;
            pdf = R_INV_PI / (alpha * alpha * cosThetaM * cosThetaM * cosThetaM * temp * temp);
        }
        if (pdf < 1e-20f) pdf = 0.;
        Real sinThetaM = sqrt(fmax(0., 1. - cosThetaM * cosThetaM));
        return r3(sinThetaM * cosPhiM, sinThetaM * sinPhiM, cosThetaM);
    }
    static DR_D Real erfinv_giles(Real x) {   // src/libcore/math.cpp:25-52
        Real w = -log((1.0 - x) * (1.0 + x)), p;
        if (w < 5.0) {
            w = w - 2.5;
            p = 2.81022636e-08; p = 3.43273939e-07 + p * w; p = -3.5233877e-06 + p * w; p = -4.39150654e-06 + p * w;
            p = 0.00021858087 + p * w; p = -0.00125372503 + p * w; p = -0.00417768164 + p * w;
            p = 0.246640727 + p * w; p = 1.50140941 + p * w;
        } else {
            w = sqrt(w) - 3.0;
            p = -0.000200214257; p = 0.000100950558 + p * w; p = 0.00134934322 + p * w; p = -0.00367342844 + p * w;
            p = 0.00573950773 + p * w; p = -0.0076224613 + p * w; p = 0.00943887047 + p * w;
            p = 1.00167406 + p * w; p = 2.83297682 + p * w;
        }
        return p * x;
    }
    static DR_D Real erf_as(Real x) {   // math.cpp:54-70 (A&S 7.1.26)
        Real sign = signbit(x) ? -1. : 1.;
        x = fabs(x);
        Real t = 1.0 / (1.0 + 0.3275911 * x);
        Real y = 1.0 - (((((1.061405429 * t + -1.453152027) * t) + 1.421413741) * t + -0.284496736) * t + 0.254829592) * t * exp(-x * x);
        return sign * y;
    }
    DR_D R2 sample_visible11(Real thetaI, Real sx, Real sy, Real epsilon) const {
        const Real SQRT_PI_INV = 0.5641895835477563;
        R2 slope;
        if (!ggx) {
            if (thetaI < 1e-4f) {
                Real r = sqrt(-log(1.0 - sx)), s, c;
                sincos(2. * R_PI * sy, &s, &c);
                return r2(r * c, r * s);
            }
            Real tanThetaI = tan(thetaI), cotThetaI = 1. / tanThetaI;
            Real a = -1., c = erf_as(cotThetaI);
            Real sample_x = fmax(sx, (Real) 1e-6f);
            Real fit = 1. + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
            Real b = c - (1. + c) * pow(1. - sample_x, fit);
            Real normalization = 1. / (1. + c + SQRT_PI_INV * tanThetaI * exp(-cotThetaI * cotThetaI));
            int it = 0;
            while (++it < 10) {
                if (!(b >= a && b <= c)) b = 0.5 * (a + c);
                Real invErf = erfinv_giles(b);
                Real value = normalization * (1. + b + SQRT_PI_INV * tanThetaI * exp(-invErf * invErf)) - sample_x;
                Real derivative = normalization * (1. - invErf * tanThetaI);
                if (fabs(value) < 1e-5f) break;
                if (value > 0.) c = b; else a = b;
                b -= value / derivative;
            }
            slope.x = erfinv_giles(b);
            slope.y = erfinv_giles(2.0 * fmax(sy, (Real) 1e-6f) - 1.0);
        } else {
            if (thetaI < 1e-4f) {
                Real r = safe_sqrt(sx / (1. - sx)), s, c;
                sincos(2. * R_PI * sy, &s, &c);
                return r2(r * c, r * s);
            }
            Real tanThetaI = tan(thetaI);
            Real a = 1. / tanThetaI;
            Real G1 = 2.0 / (1.0 + safe_sqrt(1.0 + 1.0 / (a * a)));
            Real A = 2.0 * sx / G1 - 1.0;
            if (fabs(A) == 1.) A -= (signbit(A) ? -1. : 1.) * epsilon;
            Real tmp = 1.0 / (A * A - 1.0);
            Real B = tanThetaI;
            Real D = safe_sqrt(B * B * tmp * tmp - (A * A - B * B) * tmp);
            Real slope_x_1 = B * tmp - D, slope_x_2 = B * tmp + D;
            slope.x = (A < 0.0 || slope_x_2 > 1.0 / tanThetaI) ? slope_x_1 : slope_x_2;
            Real S;
            if (sy > 0.5) { S = 1.0; sy = 2.0 * (sy - 0.5); }
            else { S = -1.0; sy = 2.0 * (0.5 - sy); }
            Real z = (sy * (sy * (sy * (-0.365728915865723) + 0.790235037209296) - 0.424965825137544) + 0.000152998850436920) /
                      (sy * (sy * (sy * (sy * 0.169507819808272 - 0.397203533833404) - 0.232500544458471) + 1.) - 0.539825872510702);
            slope.y = S * z * sqrt(1.0 + slope.x * slope.x);
        }
        return slope;
    }
    DR_D R3 sample_visible(R3 _wi, Real sx, Real sy, Real epsilon) const {
        R3 wi = normalize(r3(alpha * _wi.x, alpha * _wi.y, _wi.z));
        Real theta = 0., phi = 0.;
        if (wi.z < 0.99999) { theta = acos(wi.z); phi = atan2(wi.y, wi.x); }
        Real sinPhi, cosPhi;
        sincos(phi, &sinPhi, &cosPhi);
        R2 slope = sample_visible11(theta, sx, sy, epsilon);
        slope = r2(cosPhi * slope.x - sinPhi * slope.y, sinPhi * slope.x + cosPhi * slope.y);
        slope.x *= alpha; slope.y *= alpha;
        Real normalization = 1.0 / sqrt(slope.x * slope.x + slope.y * slope.y + 1.0);
        return r3(-slope.x * normalization, -slope.y * normalization, normalization);
    }
    DR_D R3 sample(R3 wi, Real sx, Real sy, Real &pdf, Real epsilon) const {
        if (visible) { R3 m = sample_visible(wi, sx, sy, epsilon); pdf = pdf_visible(wi, m); return m; }
        return sample_all(sx, sy, pdf);
    }
};

DR_D R3 reflect_z(R3 wi) { return r3(-wi.x, -wi.y, wi.z); }

// ---- smooth plastic (plastic.cpp): m.refl = diffuseReflectance, m.trans = specularReflectance, m.eta.x = eta; the library
// derives m.k.x = m_fdrInt (fresnelDiffuseReflectance(1 / eta), util.cpp:822-867) and m.k.y = m_specularSamplingWeight
// (plastic.cpp:196-203) when the scene is created.
DR_D Real plastic_prob_specular(const Mat &m, Real Fi) {            // plastic.cpp:291-295
    const Real w = m.k.y;
    return (Fi * w) / (Fi * w + (1. - Fi) * (1. - w));
}
DR_D R3 plastic_diffuse(const Mat &m) {                             // :266-271
    R3 diff = m.refl;
    if (m.flags & DR_MAT_NONLINEAR) { diff.x /= 1. - diff.x * m.k.x; diff.y /= 1. - diff.y * m.k.x; diff.z /= 1. - diff.z * m.k.x; }
    else diff = diff / (1. - m.k.x);
    return diff;
}

// ---- rough plastic (roughplastic.cpp): m.refl = diffuseReflectance, m.trans = specularReflectance, m.eta.x = eta, m.alpha; m.table =
// the rough transmittance reduced to (eta, alpha) (include/drmlt_b200.h DR_ROUGH_TABLE_*), whose slot [102] the library fills with
// m_specularSamplingWeight (roughplastic.cpp:273-277) when the scene is created.
// RoughTransmittance::eval with eta and alpha fixed (rtrans.h:136-146) = evalCubicInterp1D (spline.cpp:23-60) of cos^(1/4)
DR_D Real rough_transmittance(const double *tab, Real cosTheta) {
    if (!(cosTheta >= 0.)) return 0.;
    const Real x = pow(fabs(cosTheta), 0.25);
    if (!(x >= 0. && x <= 1.)) return 0.;                      // (spline.cpp:25-26; the clamp below maps it to 0)
    const int size = DR_ROUGH_TABLE_THETA;
    Real t = (x - 0.) * (size - 1) / (1. - 0.);
    const int k = max(0, min((int) t, size - 2));
    const Real f0 = __ldg(tab + k), f1 = __ldg(tab + k + 1);
    const Real d0 = k > 0 ? 0.5 * (f1 - __ldg(tab + k - 1)) : f1 - f0;
    const Real d1 = k + 2 < size ? 0.5 * (__ldg(tab + k + 2) - f0) : f1 - f0;
    t = t - (Real) k;
    const Real t2 = t * t, t3 = t2 * t;
    const Real result = (2 * t3 - 3 * t2 + 1) * f0 + (-2 * t3 + 3 * t2) * f1 + (t3 - 2 * t2 + t) * d0 + (t3 - t2) * d1;
    return fmin(1., fmax(0., result));
}
DR_D Real roughplastic_prob_specular(const Mat &m, Real cosThetaI) {   // roughplastic.cpp:407-416
    const Real p = 1. - rough_transmittance(m.table, cosThetaI), w = __ldg(m.table + 102);
    return (p * w) / (p * w + (1. - p) * (1. - w));
}
DR_D R3 roughplastic_eval(const Mat &m, R3 wi, R3 wo) {               // :325-381 (both components enabled)
    Microfacet distr(m);
    const R3 H = normalize(wo + wi);
    const Real D = distr.eval(H);
    Real cosThetaT;
    const Real F = fresnel_dielectric_ext(dot(wi, H), cosThetaT, m.eta.x);
    const Real G = distr.G(wi, wo, H);
    R3 result = m.trans * (F * D * G / (4.0 * wi.z));
    R3 diff = m.refl;
    const Real T12 = rough_transmittance(m.table, wi.z), T21 = rough_transmittance(m.table, wo.z);
    const Real Fdr = 1. - __ldg(m.table + DR_ROUGH_TABLE_THETA);
    if (m.flags & DR_MAT_NONLINEAR) { diff.x /= 1. - diff.x * Fdr; diff.y /= 1. - diff.y * Fdr; diff.z /= 1. - diff.z * Fdr; }
    else diff = diff / (1. - Fdr);
    return result + diff * (R_INV_PI * wo.z * T12 * T21 * (1. / (m.eta.x * m.eta.x)));
}
DR_D Real roughplastic_pdf(const Mat &m, R3 wi, R3 wo) {              // :383-432
    Microfacet distr(m);
    const R3 H = normalize(wo + wi);
    const Real probSpecular = roughplastic_prob_specular(m, wi.z);
    const Real dwh_dwo = 1.0 / (4.0 * dot(wo, H));
    return distr.pdf(wi, H) * dwh_dwo * probSpecular + (1. - probSpecular) * (R_INV_PI * wo.z);
}

// ---- nested (one-sided) evaluation
DR_D R3 bsdf_eval_nested(const Mat &m, R3 wi, R3 wo, int mode, int measure) {
    switch (m.type) {
    case DR_BSDF_DIFFUSE:
        if (measure != MEAS_SOLID_ANGLE || wi.z <= 0. || wo.z <= 0.) return r3(0.);
        return m.refl * (R_INV_PI * wo.z);
    case DR_BSDF_CONDUCTOR:
        if (measure != MEAS_DISCRETE || wi.z <= 0. || wo.z <= 0. || fabs(dot(reflect_z(wi), wo) - 1.) > R_DELTA_EPS) return r3(0.);
        return m.refl * fresnel_conductor(wi.z, m.eta, m.k);
    case DR_BSDF_ROUGHCONDUCTOR: {
        if (measure != MEAS_SOLID_ANGLE || wi.z <= 0. || wo.z <= 0.) return r3(0.);
        R3 H = normalize(wo + wi);
        Microfacet distr(m);
        const Real D = distr.eval(H);
        if (D == 0.) return r3(0.);
        const R3 F = fresnel_conductor(dot(wi, H), m.eta, m.k) * m.refl;
        const Real G = distr.G(wi, wo, H);
        return F * (D * G / (4.0 * wi.z));
    }
    case DR_BSDF_PLASTIC: {                                          // plastic.cpp:240-277
        if (wo.z <= 0. || wi.z <= 0.) return r3(0.);
        Real cosThetaT;
        const Real Fi = fresnel_dielectric_ext(wi.z, cosThetaT, m.eta.x);
        if (measure == MEAS_DISCRETE) {
            if (fabs(dot(reflect_z(wi), wo) - 1.) < R_DELTA_EPS) return m.trans * Fi;
        } else if (measure == MEAS_SOLID_ANGLE) {
            const Real Fo = fresnel_dielectric_ext(wo.z, cosThetaT, m.eta.x);
            return plastic_diffuse(m) * (R_INV_PI * wo.z * (1. / (m.eta.x * m.eta.x)) * (1. - Fi) * (1. - Fo));
        }
        return r3(0.);
    }
    case DR_BSDF_ROUGHPLASTIC:
        if (measure != MEAS_SOLID_ANGLE || wi.z <= 0. || wo.z <= 0.) return r3(0.);
        return roughplastic_eval(m, wi, wo);
    case DR_BSDF_ROUGHDIELECTRIC: {                                  // roughdielectric.cpp:270-348
        if (measure != MEAS_SOLID_ANGLE || wi.z == 0.) return r3(0.);
        const Real mEta = m.eta.x, mInvEta = 1. / mEta;
        const bool reflect = wi.z * wo.z > 0.;
        R3 H;
        if (reflect) H = normalize(wo + wi);
        else H = normalize(wi + wo * (wi.z > 0. ? mEta : mInvEta));
        H = H * copysign(1.0, H.z);
        Microfacet distr(m);
        const Real D = distr.eval(H);
        if (D == 0.) return r3(0.);
        Real cosThetaT;
        const Real F = fresnel_dielectric_ext(dot(wi, H), cosThetaT, mEta);
        const Real G = distr.G(wi, wo, H);
        if (reflect) return m.refl * (F * D * G / (4.0 * fabs(wi.z)));
        const Real eta = wi.z > 0. ? mEta : mInvEta;
        const Real sqrtDenom = dot(wi, H) + eta * dot(wo, H);
        const Real value = ((1. - F) * D * G * eta * eta * dot(wi, H) * dot(wo, H)) / (wi.z * sqrtDenom * sqrtDenom);
        const Real factor = (mode == MODE_RADIANCE) ? (wi.z > 0. ? mInvEta : mEta) : 1.0;
        return m.trans * fabs(value * factor * factor);
    }
    case DR_BSDF_DIELECTRIC: {
        if (measure != MEAS_DISCRETE) return r3(0.);
        Real eta = m.eta.x, invEta = 1. / eta, cosThetaT;
        Real F = fresnel_dielectric_ext(wi.z, cosThetaT, eta);
        if (wi.z * wo.z >= 0.) {
            if (fabs(dot(reflect_z(wi), wo) - 1.) > R_DELTA_EPS) return r3(0.);
            return m.refl * F;
        } else {
            Real scale = -(cosThetaT < 0. ? invEta : eta);
            R3 refr = r3(scale * wi.x, scale * wi.y, cosThetaT);
            if (fabs(dot(refr, wo) - 1.) > R_DELTA_EPS) return r3(0.);
            Real factor = (mode == MODE_RADIANCE) ? (cosThetaT < 0. ? invEta : eta) : 1.0;
            return m.trans * (factor * factor * (1. - F));
        }
    }
    }
    return r3(0.);
}

DR_D Real bsdf_pdf_nested(const Mat &m, R3 wi, R3 wo, int measure) {
    switch (m.type) {
    case DR_BSDF_DIFFUSE:
        if (measure != MEAS_SOLID_ANGLE || wi.z <= 0. || wo.z <= 0.) return 0.;
        return R_INV_PI * wo.z;
    case DR_BSDF_CONDUCTOR:
        if (measure != MEAS_DISCRETE || wi.z <= 0. || wo.z <= 0. || fabs(dot(reflect_z(wi), wo) - 1.) > R_DELTA_EPS) return 0.;
        return 1.0;
    case DR_BSDF_ROUGHCONDUCTOR: {
        if (measure != MEAS_SOLID_ANGLE || wi.z <= 0. || wo.z <= 0.) return 0.;
        R3 H = normalize(wo + wi);
        Microfacet distr(m);
        if (distr.visible) return distr.eval(H) * distr.smithG1(wi, H) / (4.0 * wi.z);
        return distr.pdf(wi, H) / (4. * absdot(wo, H));
    }
    case DR_BSDF_PLASTIC: {                                          // plastic.cpp:279-307
        if (wo.z <= 0. || wi.z <= 0.) return 0.;
        Real cosThetaT;
        const Real probSpecular = plastic_prob_specular(m, fresnel_dielectric_ext(wi.z, cosThetaT, m.eta.x));
        if (measure == MEAS_DISCRETE) {
            if (fabs(dot(reflect_z(wi), wo) - 1.) < R_DELTA_EPS) return probSpecular;
        } else if (measure == MEAS_SOLID_ANGLE) return R_INV_PI * wo.z * (1. - probSpecular);
        return 0.;
    }
    case DR_BSDF_ROUGHPLASTIC:
        if (measure != MEAS_SOLID_ANGLE || wi.z <= 0. || wo.z <= 0.) return 0.;
        return roughplastic_pdf(m, wi, wo);
    case DR_BSDF_ROUGHDIELECTRIC: {                                  // roughdielectric.cpp:350-420 (both components enabled)
        if (measure != MEAS_SOLID_ANGLE) return 0.;
        const Real mEta = m.eta.x, mInvEta = 1. / mEta;
        const bool reflect = wi.z * wo.z > 0.;
        R3 H;
        Real dwh_dwo;
        if (reflect) {
            H = normalize(wo + wi);
            dwh_dwo = 1.0 / (4.0 * dot(wo, H));
        } else {
            const Real eta = wi.z > 0. ? mEta : mInvEta;
            H = normalize(wi + wo * eta);
            const Real sqrtDenom = dot(wi, H) + eta * dot(wo, H);
            dwh_dwo = (eta * eta * dot(wo, H)) / (sqrtDenom * sqrtDenom);
        }
        H = H * copysign(1.0, H.z);
        Microfacet sampleDistr(m);
        if (!sampleDistr.visible) sampleDistr.scale_alpha(1.2f - 0.2f * sqrt(fabs(wi.z)));
        Real prob = sampleDistr.pdf(wi * copysign(1.0, wi.z), H);
        Real cosThetaT;
        const Real F = fresnel_dielectric_ext(dot(wi, H), cosThetaT, mEta);
        prob *= reflect ? F : (1. - F);
        return fabs(prob * dwh_dwo);
    }
    case DR_BSDF_DIELECTRIC: {
        if (measure != MEAS_DISCRETE) return 0.;
        Real eta = m.eta.x, invEta = 1. / eta, cosThetaT;
        Real F = fresnel_dielectric_ext(wi.z, cosThetaT, eta);
        if (wi.z * wo.z >= 0.) {
            if (fabs(dot(reflect_z(wi), wo) - 1.) > R_DELTA_EPS) return 0.;
            return F;
        } else {
            Real scale = -(cosThetaT < 0. ? invEta : eta);
            R3 refr = r3(scale * wi.x, scale * wi.y, cosThetaT);
            if (fabs(dot(refr, wo) - 1.) > R_DELTA_EPS) return 0.;
            return 1. - F;
        }
    }
    }
    return 0.;
}

struct BsdfSample { R3 wo; R3 weight; Real pdf; int sampledType; Real eta; };

// `sz`: the number an EUsesSampler BSDF draws from bRec.sampler inside sample() (roughdielectric.cpp:555)
DR_D void bsdf_sample_nested(const Mat &m, R3 wi, int mode, Real sx, Real sy, Real sz, Real epsilon, BsdfSample &r) {
    r.weight = r3(0.); r.pdf = 0.; r.sampledType = 0; r.eta = 1.; r.wo = r3(0., 0., 1.);
    switch (m.type) {
    case DR_BSDF_PLASTIC: {                                          // plastic.cpp:368-412
        if (wi.z <= 0.) return;
        Real cosThetaT;
        const Real Fi = fresnel_dielectric_ext(wi.z, cosThetaT, m.eta.x);
        const Real probSpecular = plastic_prob_specular(m, Fi);
        r.eta = 1.;
        if (sx < probSpecular) {
            r.sampledType = BT_DELTA_R; r.wo = reflect_z(wi); r.pdf = probSpecular;
            r.weight = m.trans * (Fi / probSpecular);
        } else {
            r.sampledType = BT_DIFFUSE_R;
            r.wo = square_to_cosine_hemisphere((sx - probSpecular) / (1. - probSpecular), sy);
            const Real Fo = fresnel_dielectric_ext(r.wo.z, cosThetaT, m.eta.x);
            r.pdf = (1. - probSpecular) * (R_INV_PI * r.wo.z);
            r.weight = plastic_diffuse(m) * ((1. / (m.eta.x * m.eta.x)) * (1. - Fi) * (1. - Fo) / (1. - probSpecular));
        }
        return;
    }
    case DR_BSDF_ROUGHPLASTIC: {                                     // roughplastic.cpp:434-491 (both components enabled)
        if (wi.z <= 0.) return;
        Microfacet distr(m);
        const Real probSpecular = roughplastic_prob_specular(m, wi.z);
        r.eta = 1.;
        if (sy < probSpecular) {
            Real tpdf = 0.;
            const R3 mm = distr.sample(wi, sx, sy / probSpecular, tpdf, epsilon);
            r.wo = mm * (2. * dot(wi, mm)) - wi;
            r.sampledType = BT_GLOSSY_R;
            if (r.wo.z <= 0.) return;
        } else {
            r.sampledType = BT_DIFFUSE_R;
            r.wo = square_to_cosine_hemisphere(sx, (sy - probSpecular) / (1. - probSpecular));
        }
        const Real pdf = (wi.z <= 0. || r.wo.z <= 0.) ? 0. : roughplastic_pdf(m, wi, r.wo);      // "guard against numerical imprecisions" (:484-490)
        if (pdf == 0.) return;
        r.pdf = pdf;
        r.weight = roughplastic_eval(m, wi, r.wo) / pdf;
        return;
    }
    case DR_BSDF_ROUGHDIELECTRIC: {                                  // roughdielectric.cpp:514-611 (both components enabled)
        const Real mEta = m.eta.x, mInvEta = 1. / mEta;
        Microfacet distr(m);
        Microfacet sampleDistr(distr);
        if (!distr.visible) sampleDistr.scale_alpha(1.2f - 0.2f * sqrt(fabs(wi.z)));
        Real microfacetPDF = 0.;
        const R3 mm = sampleDistr.sample(wi * copysign(1.0, wi.z), sx, sy, microfacetPDF, epsilon);
        if (microfacetPDF == 0.) return;
        float temporaryPdf = (float) microfacetPDF;              // sic: single precision in the reference (:543)
        Real cosThetaT;
        const Real F = fresnel_dielectric_ext(dot(wi, mm), cosThetaT, mEta);
        R3 weight = r3(1.);
        bool sampleReflection = true;
        if (sz > F) { sampleReflection = false; temporaryPdf *= (float) (1. - F); }
        else temporaryPdf *= (float) F;
        Real dwh_dwo;
        if (sampleReflection) {
            r.wo = mm * (2. * dot(wi, mm)) - wi;
            r.eta = 1.; r.sampledType = BT_GLOSSY_R;
            if (wi.z * r.wo.z <= 0.) return;
            weight = weight * m.refl;
            dwh_dwo = 1.0 / (4.0 * dot(r.wo, mm));
        } else {
            if (cosThetaT == 0.) return;
            const Real e = cosThetaT < 0. ? mInvEta : mEta;      // refract(): util.cpp:775-780
            r.wo = mm * (dot(wi, mm) * e + cosThetaT) - wi * e;
            r.eta = cosThetaT < 0. ? mEta : mInvEta;
            r.sampledType = BT_GLOSSY_T;
            if (wi.z * r.wo.z >= 0.) return;
            const Real factor = (mode == MODE_RADIANCE) ? (cosThetaT < 0. ? mInvEta : mEta) : 1.0;
            weight = weight * m.trans * (factor * factor);
            const Real sqrtDenom = dot(wi, mm) + r.eta * dot(r.wo, mm);
            dwh_dwo = (r.eta * r.eta * dot(r.wo, mm)) / (sqrtDenom * sqrtDenom);
        }
        if (distr.visible) weight = weight * distr.smithG1(r.wo, mm);
        else weight = weight * fabs(distr.eval(mm) * distr.G(wi, r.wo, mm) * dot(wi, mm) / (microfacetPDF * wi.z));
        temporaryPdf *= (float) fabs(dwh_dwo);
        r.pdf = (Real) temporaryPdf;
        r.weight = weight;
        return;
    }
    case DR_BSDF_DIFFUSE:
        if (wi.z <= 0.) return;
        r.wo = square_to_cosine_hemisphere(sx, sy);
        r.sampledType = BT_DIFFUSE_R; r.pdf = R_INV_PI * r.wo.z; r.weight = m.refl;
        return;
    case DR_BSDF_CONDUCTOR:
        if (wi.z <= 0.) return;
        r.sampledType = BT_DELTA_R; r.wo = reflect_z(wi); r.pdf = 1.;
        r.weight = m.refl * fresnel_conductor(wi.z, m.eta, m.k);
        return;
    case DR_BSDF_ROUGHCONDUCTOR: {
        if (wi.z < 0.) return;
        Microfacet distr(m);
        Real tpdf = 0.;
        R3 mm = distr.sample(wi, sx, sy, tpdf, epsilon);
        if (tpdf == 0.) return;
        r.wo = mm * (2. * dot(wi, mm)) - wi;
        r.sampledType = BT_GLOSSY_R;
        if (r.wo.z <= 0.) return;
        R3 F = fresnel_conductor(dot(wi, mm), m.eta, m.k) * m.refl;
        Real weight;
        if (distr.visible) weight = distr.smithG1(r.wo, mm);
        else weight = distr.eval(mm) * distr.G(wi, r.wo, mm) * dot(wi, mm) / (tpdf * wi.z);
        if (weight > 0.) { r.pdf = tpdf / (4.0 * dot(r.wo, mm)); r.weight = F * weight; }
        return;
    }
    case DR_BSDF_DIELECTRIC: {
        Real eta = m.eta.x, invEta = 1. / eta, cosThetaT;
        Real F = fresnel_dielectric_ext(wi.z, cosThetaT, eta);
        if (sx <= F) {
            r.sampledType = BT_DELTA_R; r.wo = reflect_z(wi); r.pdf = F; r.weight = m.refl;
        } else {
            r.sampledType = BT_DELTA_T;
            Real scale = -(cosThetaT < 0. ? invEta : eta);
            r.wo = r3(scale * wi.x, scale * wi.y, cosThetaT);
            r.eta = cosThetaT < 0. ? eta : invEta;
            r.pdf = 1. - F;
            Real factor = (mode == MODE_RADIANCE) ? (cosThetaT < 0. ? invEta : eta) : 1.0;
            r.weight = m.trans * (factor * factor);
        }
        return;
    }
    }
}

// ---- public: twosided adapter (twosided.cpp:107-186)
DR_D R3 bsdf_eval(const Mat &m, R3 wi, R3 wo, int mode, int measure) {
    if ((m.flags & DR_MAT_TWOSIDED) && !(wi.z > 0.)) { wi.z = -wi.z; wo.z = -wo.z; }
    return bsdf_eval_nested(m, wi, wo, mode, measure);
}
DR_D Real bsdf_pdf(const Mat &m, R3 wi, R3 wo, int measure) {
    if ((m.flags & DR_MAT_TWOSIDED) && !(wi.z > 0.)) { wi.z = -wi.z; wo.z = -wo.z; }
    return bsdf_pdf_nested(m, wi, wo, measure);
}
DR_D void bsdf_sample(const Mat &m, R3 wi, int mode, Real sx, Real sy, Real sz, Real epsilon, BsdfSample &r) {
    bool flipped = false;
    if ((m.flags & DR_MAT_TWOSIDED) && wi.z < 0.) { wi.z = -wi.z; flipped = true; }
    bsdf_sample_nested(m, wi, mode, sx, sy, sz, epsilon, r);
    if (flipped && !is_zero(r.weight) && r.pdf != 0.) r.wo.z = -r.wo.z;
}
