// pss.cuh -- primary-sample-space state of one Markov chain and its lazily evaluated proposals.
//
// The reference keeps three std::vector<Float> per sampler (current, stage-1 proposal, stage-2
// proposal; src/integrators/drmlt/drmlt_sampler.h:207-210) and fills a whole proposal vector at the
// first query of a stage (drmlt_sampler.cpp:313-394, pssmlt_sampler.cpp:124-166).  Here only the
// CURRENT vector lives in HBM (SoA: coordinate-major, chain-minor, so a warp of chains reads one
// coordinate with one coalesced 128-byte transaction).  Proposal coordinates are pure functions of
// (current value, keyed uniforms) and are recomputed in registers when the path sampler asks for
// them; nothing but the accepted vector is ever written back.
//
// Transition kernels: src/integrators/drmlt/tools/transition.h:23-190 (Kelemen, Gaussian,
// WrappedCauchy, Identity), PSSMLTSampler::mutate (pssmlt_sampler.h:117-147).
#pragma once
#include "common.cuh"
#include "../../include/drmlt_b200.h"

enum { PSS_ARRAY = 0, PSS_BOOT = 1, PSS_STAGE1 = 2, PSS_STAGE2 = 3, PSS_REVERSE = 4 };
enum { SMP_SENSOR = 0, SMP_EMITTER = 1, SMP_DIRECT = 2 };

struct PssParams {          // per-launch constants
    uint64_t seed;
    int integrator, type;
    float kel_s2, kel_logRatio;        // DRMLT stage 1 (x1.9 for orbital), drmlt_sampler.h:201-205
    float sigma2;                      // scaleSecond * sigma
    float cauchy_disp;                 // 2 rho / (1 + rho^2), rho = exp(-1/4)
    int pss_kelemen; float pss_s2, pss_logRatio, pss_sigma;   // PSSMLT
    uint32_t identity1, identity2;     // bit s: sampler s uses the identity kernel in stage 1 / stage 2
};

DR_D float wrap_reflect(float y) { return y > 1.f ? 2.f - y : (y <= 0.f ? fabsf(y) : y); }   // drmlt_sampler.h:140-144

DR_D float kelemen_sample(float xi, float s2, float logRatio) {   // transition.h:96-110
    float sign;
    if (xi < 0.5f) { sign = 1.f; xi *= 2.0f; } else { sign = -1.f; xi = 2.0f * (xi - 0.5f); }
    return sign * s2 * expf((1.f - xi) * logRatio);
}
DR_D float kelemen_logpdf(float du, float s1, float s2, float logRatio) {   // transition.h:112-121
    float d = fabsf(du);
    if (d < s1 || d > s2) return -INFINITY;
    return logf(1.0f / (2.0f * d * (-logRatio)));
}
DR_D float gaussian_sample(float xi1, float xi2, float sigma) {   // transition.h:63-68
    return sqrtf(-2.0f * logf(1.f - xi1)) * cospif(2.0f * xi2) * sigma;
}
DR_D float cauchy_sample(float xi, float disp) {   // transition.h:162-178
    float sign;
    if (xi < 0.5f) { sign = 1.f; xi *= 2.0f; } else { sign = -1.f; xi = 2.0f * (xi - 0.5f); }
    float V = cospif(2.0f * xi);
    return sign * safe_acosf((V + disp) / (1.0f + disp * V));
}
DR_D float pssmlt_mutate(float value, float xi1, float xi2, const PssParams &pp) {   // pssmlt_sampler.h:117-147
    if (pp.pss_kelemen) {
        bool add;
        if (xi1 < 0.5f) { add = true; xi1 *= 2.0f; } else { add = false; xi1 = 2.0f * (xi1 - 0.5f); }
        float dv = pp.pss_s2 * expf(xi1 * pp.pss_logRatio);
        if (add) { value += dv; if (value > 1.f) value -= 1.f; }
        else { value -= dv; if (value < 0.f) value += 1.f; }
    } else {
        float dv = sqrtf(-2.f * logf(1.f - xi1)) * cospif(2.f * xi2);
        float r = fmodf(value + pp.pss_sigma * dv, 1.0f);
        value = (r < 0.0f) ? r + 1.0f : r;
    }
    return value;
}

struct Pss {
    const PssParams *pp;
    const float *xs[3];       // coordinate (s,k) of this chain lives at xs[s][k * stride]
    size_t stride;
    int dim[3];
    uint64_t chain;           // chain id (stages) or bootstrap sample index (PSS_BOOT)
    uint32_t mut;
    int mode;
    bool largeStep;
    bool lightTracing;        // nextStage(current->t == 1): emitter sampler keeps its real stage-2 kernel
    int pos[3];
    int maxIdx[3];            // largest index touched in this stage (m_dimStage*, drmlt_sampler.cpp:237-238)
    int cacheKey; float2 cacheVal;

    DR_D void begin(int mode_) {
        mode = mode_;
        pos[0] = pos[1] = pos[2] = 0;
        maxIdx[0] = maxIdx[1] = maxIdx[2] = 0;
        cacheKey = -1;
    }
    DR_D float xat(int s, int k) const { return k < dim[s] ? xs[s][(size_t) k * stride] : 0.f; }
    DR_D bool identity1(int s) const { return (pp->identity1 >> s) & 1u; }
    DR_D bool identity2(int s) const {
        if ((pp->identity1 >> s) & 1u) return true;               // setStagesToIdentity
        if ((pp->identity2 >> s) & 1u) return !lightTracing;      // handleLightTracing: stage2 = Identity, stageLT = real kernel
        return false;
    }
    // un-wrapped stage-1 proposal of the coordinate pair (2p, 2p+1)
    DR_D float2 prop1(int s, int p) const {
        const float x0 = xat(s, 2 * p), x1 = xat(s, 2 * p + 1);
        if (!largeStep && identity1(s)) return make_float2(x0, x1);
        const float4 u = keyed_uniform4(pp->seed, S_STAGE1 + s, chain, mut, (uint32_t) p);
        if (largeStep) return make_float2(u.x, u.z);
        if (pp->integrator == DR_INTEGRATOR_PSSMLT)
            return make_float2(pssmlt_mutate(x0, u.x, u.y, *pp), pssmlt_mutate(x1, u.z, u.w, *pp));
        if (pp->type != DR_TYPE_ORBITAL)
            return make_float2(x0 + kelemen_sample(u.x, pp->kel_s2, pp->kel_logRatio), x1 + kelemen_sample(u.z, pp->kel_s2, pp->kel_logRatio));
        const float d = kelemen_sample(u.x, pp->kel_s2, pp->kel_logRatio);   // drmlt_sampler.cpp:351-359
        float sa, ca;
        sincospif(2.0f * u.y, &sa, &ca);
        return make_float2(x0 + d * ca, x1 + d * sa);
    }
    // un-wrapped stage-2 proposal
    DR_D float2 prop2(int s, int p) const {
        const float x0 = xat(s, 2 * p), x1 = xat(s, 2 * p + 1);
        if (!largeStep && identity2(s)) return make_float2(x0, x1);
        const float4 u = keyed_uniform4(pp->seed, S_STAGE2 + s, chain, mut, (uint32_t) p);
        if (largeStep) return make_float2(u.x, u.z);   // release-build behaviour of fillSpace with m_largeStep still set
        if (pp->type != DR_TYPE_ORBITAL)
            return make_float2(x0 + gaussian_sample(u.x, u.y, pp->sigma2), x1 + gaussian_sample(u.z, u.w, pp->sigma2));
        const float2 y = prop1(s, p);                  // drmlt_sampler.cpp:361-392
        const float theta = cauchy_sample(u.x, pp->cauchy_disp);
        const float du1 = y.x - x0, du2 = y.y - x1;
        const float norm = sqrtf(du1 * du1 + du2 * du2);
        float mu = safe_acosf(-du1 / norm);
        if (-du2 < 0.f) mu = 2.f * DR_PI - mu;
        float sa, ca;
        sincosf(theta + mu, &sa, &ca);
        return make_float2(y.x + ca * norm, y.y + sa * norm);
    }
    DR_D float2 pair_value(int s, int p) const {
        switch (mode) {
        case PSS_ARRAY: return make_float2(xat(s, 2 * p), xat(s, 2 * p + 1));
        case PSS_BOOT: {
            const float4 u = keyed_uniform4(pp->seed, S_BOOT, chain, (uint32_t) s, (uint32_t) (p >> 1));
            return (p & 1) ? make_float2(u.z, u.w) : make_float2(u.x, u.y);
        }
        case PSS_STAGE1: return prop1(s, p);
        case PSS_STAGE2: return prop2(s, p);
        default: {   // PSS_REVERSE: y* = z - (y - x) (drmlt_sampler.cpp:293-296)
            const float2 y = prop1(s, p), z = prop2(s, p);
            return make_float2(z.x - (y.x - xat(s, 2 * p)), z.y - (y.y - xat(s, 2 * p + 1)));
        }
        }
    }
    DR_D float next1D(int s) {
        const int k = pos[s]++;
        maxIdx[s] = max(maxIdx[s], k);
        if (k >= dim[s] + (dim[s] & 1)) return 0.5f;    // the reference raises EError here
        const int key = (s << 16) | (k >> 1);
        if (key != cacheKey) { cacheVal = pair_value(s, k >> 1); cacheKey = key; }
        const float v = (k & 1) ? cacheVal.y : cacheVal.x;
        // DRMLT stores un-wrapped values and reflects on read; PSSMLT / bootstrap values are in [0,1)
        return (pp->integrator == DR_INTEGRATOR_DRMLT) ? wrap_reflect(v) : v;
    }
    DR_D float2 next2D(int s) { float a = next1D(s); float b = next1D(s); return make_float2(a, b); }
};
