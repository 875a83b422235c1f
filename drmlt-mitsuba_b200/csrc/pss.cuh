// pss.cuh -- primary-sample-space state of one Markov chain and its lazily evaluated proposals.
//
// The reference keeps three std::vector<Float> per sampler (current, stage-1 proposal, stage-2
// proposal; src/integrators/drmlt/drmlt_sampler.h:207-210) and fills a whole proposal vector at the
// first query of a stage (drmlt_sampler.cpp:313-394, pssmlt_sampler.cpp:124-166).  Here only the
// CURRENT vector lives in HBM (SoA of doubles: coordinate-major, chain-minor, so a warp of chains
// reads one coordinate with two coalesced 128-byte transactions).  Proposal coordinates are pure functions of
// (current value, keyed uniforms) and are recomputed in registers when the path sampler asks for
// them; nothing but the accepted vector is ever written back.
//
// Transition kernels: src/integrators/drmlt/tools/transition.h:23-190 (Kelemen, Gaussian,
// WrappedCauchy, Identity), PSSMLTSampler::mutate (pssmlt_sampler.h:117-147).
#pragma once
#include "real.cuh"
#include "../../include/drmlt_b200.h"

enum { PSS_ARRAY = 0, PSS_BOOT = 1, PSS_STAGE1 = 2, PSS_STAGE2 = 3, PSS_REVERSE = 4 };
enum { SMP_SENSOR = 0, SMP_EMITTER = 1, SMP_DIRECT = 2 };

struct PssParams {          // per-launch constants
    uint64_t seed;
    int integrator, type;
    Real kel_s2, kel_logRatio;        // DRMLT stage 1 (x1.9 for orbital), drmlt_sampler.h:201-205
    Real sigma2;                      // scaleSecond * sigma
    Real cauchy_disp;                 // 2 rho / (1 + rho^2), rho = exp(-1/4)
    int pss_kelemen; Real pss_s2, pss_logRatio, pss_sigma;   // PSSMLT
    uint32_t identity1, identity2;     // bit s: sampler s uses the identity kernel in stage 1 / stage 2
};

DR_D Real wrap_reflect(Real y) { return y > 1. ? 2. - y : (y <= 0. ? fabs(y) : y); }   // drmlt_sampler.h:140-144

DR_D Real kelemen_sample(Real xi, Real s2, Real logRatio) {   // transition.h:96-110
    Real sign;
    if (xi < 0.5) { sign = 1.; xi *= 2.0; } else { sign = -1.; xi = 2.0 * (xi - 0.5); }
    return sign * s2 * exp((1. - xi) * logRatio);
}
DR_D Real kelemen_logpdf(Real du, Real s1, Real s2, Real logRatio) {   // transition.h:112-121
    Real d = fabs(du);
    if (d < s1 || d > s2) return -INFINITY;
    return log(1.0 / (2.0 * d * (-logRatio)));
}
DR_D Real gaussian_sample(Real xi1, Real xi2, Real sigma) {   // transition.h:63-68
    return sqrt(-2.0 * log(1. - xi1)) * cospi(2.0 * xi2) * sigma;
}
DR_D Real cauchy_sample(Real xi, Real disp) {   // transition.h:162-178
    Real sign;
    if (xi < 0.5) { sign = 1.; xi *= 2.0; } else { sign = -1.; xi = 2.0 * (xi - 0.5); }
    Real V = cospi(2.0 * xi);
    return sign * safe_acos((V + disp) / (1.0 + disp * V));
}
DR_D Real pssmlt_mutate(Real value, Real xi1, Real xi2, const PssParams &pp) {   // pssmlt_sampler.h:117-147
    if (pp.pss_kelemen) {
        bool add;
        if (xi1 < 0.5) { add = true; xi1 *= 2.0; } else { add = false; xi1 = 2.0 * (xi1 - 0.5); }
        Real dv = pp.pss_s2 * exp(xi1 * pp.pss_logRatio);
        if (add) { value += dv; if (value > 1.) value -= 1.; }
        else { value -= dv; if (value < 0.) value += 1.; }
    } else {
        Real dv = sqrt(-2. * log(1. - xi1)) * cospi(2. * xi2);
        Real r = fmod(value + pp.pss_sigma * dv, 1.0);
        value = (r < 0.0) ? r + 1.0 : r;
    }
    return value;
}

struct Pss {
    const PssParams *pp;
    const void *xs[3];        // coordinate (s,k) of this chain lives at xs[s][k * stride] (double state, or float replay input)
    bool f32;                 // xs point to float arrays (dr_eval_paths replays host vectors)
    size_t stride;
    int dim[3];
    uint64_t chain;           // chain id (stages) or bootstrap sample index (PSS_BOOT)
    uint32_t mut;
    int mode;
    bool largeStep;
    bool lightTracing;        // nextStage(current->t == 1): emitter sampler keeps its real stage-2 kernel
    int pos[3];
    int maxIdx[3];            // largest index touched in this stage (m_dimStage*, drmlt_sampler.cpp:237-238)
    int cacheKey; R2 cacheVal;

    DR_D void begin(int mode_) {
        mode = mode_;
        pos[0] = pos[1] = pos[2] = 0;
        maxIdx[0] = maxIdx[1] = maxIdx[2] = 0;
        cacheKey = -1;
    }
    DR_D Real xat(int s, int k) const {
        if (k >= dim[s]) return 0.0;
        return f32 ? (Real) static_cast<const float *>(xs[s])[(size_t) k * stride] : static_cast<const double *>(xs[s])[(size_t) k * stride];
    }
    DR_D bool identity1(int s) const { return (pp->identity1 >> s) & 1u; }
    DR_D bool identity2(int s) const {
        if ((pp->identity1 >> s) & 1u) return true;               // setStagesToIdentity
        if ((pp->identity2 >> s) & 1u) return !lightTracing;      // handleLightTracing: stage2 = Identity, stageLT = real kernel
        return false;
    }
    // un-wrapped stage-1 proposal of the coordinate pair (2p, 2p+1)
    DR_D R2 prop1(int s, int p) const {
        const Real x0 = xat(s, 2 * p), x1 = xat(s, 2 * p + 1);
        if (!largeStep && identity1(s)) return r2(x0, x1);
        const float4 u = keyed_uniform4(pp->seed, S_STAGE1 + s, chain, mut, (uint32_t) p);
        if (largeStep) return r2(u.x, u.z);
        if (pp->integrator == DR_INTEGRATOR_PSSMLT)
            return r2(pssmlt_mutate(x0, u.x, u.y, *pp), pssmlt_mutate(x1, u.z, u.w, *pp));
        if (pp->type != DR_TYPE_ORBITAL)
            return r2(x0 + kelemen_sample(u.x, pp->kel_s2, pp->kel_logRatio), x1 + kelemen_sample(u.z, pp->kel_s2, pp->kel_logRatio));
        const Real d = kelemen_sample(u.x, pp->kel_s2, pp->kel_logRatio);   // drmlt_sampler.cpp:351-359
        Real sa, ca;
        sincospi(2.0 * u.y, &sa, &ca);
        return r2(x0 + d * ca, x1 + d * sa);
    }
    // un-wrapped stage-2 proposal
    DR_D R2 prop2(int s, int p) const {
        const Real x0 = xat(s, 2 * p), x1 = xat(s, 2 * p + 1);
        if (!largeStep && identity2(s)) return r2(x0, x1);
        const float4 u = keyed_uniform4(pp->seed, S_STAGE2 + s, chain, mut, (uint32_t) p);
        if (largeStep) return r2(u.x, u.z);   // release-build behaviour of fillSpace with m_largeStep still set
        if (pp->type != DR_TYPE_ORBITAL)
            return r2(x0 + gaussian_sample(u.x, u.y, pp->sigma2), x1 + gaussian_sample(u.z, u.w, pp->sigma2));
        const R2 y = prop1(s, p);                  // drmlt_sampler.cpp:361-392
        const Real theta = cauchy_sample(u.x, pp->cauchy_disp);
        const Real du1 = y.x - x0, du2 = y.y - x1;
        const Real norm = sqrt(du1 * du1 + du2 * du2);
        Real mu = safe_acos(-du1 / norm);
        if (-du2 < 0.) mu = 2. * R_PI - mu;
        Real sa, ca;
        sincos(theta + mu, &sa, &ca);
        return r2(y.x + ca * norm, y.y + sa * norm);
    }
    DR_D R2 pair_value(int s, int p) const {
        switch (mode) {
        case PSS_ARRAY: return r2(xat(s, 2 * p), xat(s, 2 * p + 1));
        case PSS_BOOT: {
            const float4 u = keyed_uniform4(pp->seed, S_BOOT, chain, (uint32_t) s, (uint32_t) (p >> 1));
            return (p & 1) ? r2(u.z, u.w) : r2(u.x, u.y);
        }
        case PSS_STAGE1: return prop1(s, p);
        case PSS_STAGE2: return prop2(s, p);
        default: {   // PSS_REVERSE: y* = z - (y - x) (drmlt_sampler.cpp:293-296)
            const R2 y = prop1(s, p), z = prop2(s, p);
            return r2(z.x - (y.x - xat(s, 2 * p)), z.y - (y.y - xat(s, 2 * p + 1)));
        }
        }
    }
    DR_D Real next1D(int s) {
        const int k = pos[s]++;
        maxIdx[s] = max(maxIdx[s], k);
        if (k >= dim[s] + (dim[s] & 1)) return 0.5;    // the reference raises EError here
        const int key = (s << 16) | (k >> 1);
        if (key != cacheKey) { cacheVal = pair_value(s, k >> 1); cacheKey = key; }
        const Real v = (k & 1) ? cacheVal.y : cacheVal.x;
        // DRMLT stores un-wrapped values and reflects on read; PSSMLT / bootstrap values are in [0,1)
        return (pp->integrator == DR_INTEGRATOR_DRMLT) ? wrap_reflect(v) : v;
    }
    DR_D R2 next2D(int s) { Real a = next1D(s); Real b = next1D(s); return r2(a, b); }
};
