// pss.cuh -- primary-sample-space state of one Markov chain: transition kernels, proposal
// vectors and the coordinate reader of the path samplers.
//
// The reference keeps three std::vector<Float> per sampler (current, stage-1 proposal, stage-2
// proposal; src/integrators/drmlt/drmlt_sampler.h:207-210) and fills a whole proposal vector at
// the first query of a stage (drmlt_sampler.cpp:313-394, pssmlt_sampler.cpp:124-166).  Here every
// chain owns four coordinate buffers in HBM -- X (current), Y (stage 1), Z (stage 2), R (Green's
// reverse state) -- each `nU` doubles, contiguous per chain, read and written as 16-byte pairs.
// A proposal is filled ONCE, by the chain kernel, when a path starts; the wavefront stages that
// walk the path only read pairs back (no transcendental math in the ray-bound kernels), and an
// accepted proposal is committed by copying its pairs into X.
//
// drmlt + mmlt: the strategy coordinate uses the identity kernel (drmlt_proc.cpp:133-136), so (s,t)
// only changes in a large step, which redraws every coordinate.  A path of strategy (s,t) can only
// consume sensor coordinates [0, 2t) and emitter coordinates [0, 2s); all other coordinates can never
// influence anything before they are redrawn, so only that subset is filled and committed
// ("subset mode"; exact, not an approximation).  Everything else (pssmlt, technique=path/bdpt,
// green + timidAfterLarge whose reverse state reads x after a strategy change) uses full vectors.
//
// Transition kernels: src/integrators/drmlt/tools/transition.h:23-190 (Kelemen, Gaussian,
// WrappedCauchy, Identity), PSSMLTSampler::mutate (pssmlt_sampler.h:117-147).
#pragma once
#include "real.cuh"
#include "../../include/drmlt_b200.h"

enum { SMP_SENSOR = 0, SMP_EMITTER = 1, SMP_DIRECT = 2 };
enum { UB_X = 0, UB_Y = 1, UB_Z = 2, UB_R = 3, UB_COUNT = 4 };

struct PssParams {          // per-launch constants
    uint64_t seed;
    int integrator, type;
    Real kel_s2, kel_logRatio;        // DRMLT stage 1 (x1.9 for orbital), drmlt_sampler.h:201-205
    Real sigma2;                      // scaleSecond * sigma
    Real cauchy_disp;                 // 2 rho / (1 + rho^2), rho = exp(-1/4)
    int pss_kelemen; Real pss_s2, pss_logRatio, pss_sigma;   // PSSMLT
    uint32_t identity1, identity2;     // bit s: sampler s uses the identity kernel in stage 1 / stage 2
    int subset;                        // subset mode (see above)
    int off[3];                        // first slot of each sampler inside a coordinate buffer (even)
    int nU;                            // doubles per buffer (even)
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

// 16-byte pair access to a coordinate buffer
DR_D R2 ub_load(const double *buf, int slot) { const double2 v = *reinterpret_cast<const double2 *>(buf + slot); return r2(v.x, v.y); }
DR_D void ub_store(double *buf, int slot, R2 v) { *reinterpret_cast<double2 *>(buf + slot) = make_double2(v.x, v.y); }

// Four uniforms of one coordinate pair: (coordinate 2p, draw 0), (2p, draw 1), (2p + 1, draw 0), (2p + 1, draw 1).
struct U4 { Real x, y, z, w; };

// Replay table of one chain (dr_chain_replay, include/drmlt_b200.h): the uniforms of a recorded chain of the reference in the
// keyed address space.  D = dim coordinates per sampler; doubles; NaN = never drawn by the recorded chain:
//   [3][D] seed state per sampler | per mutation: [4] coins (large step, accept 1, accept 2, mixture), [3][2 D] stage-1 draws
//   (sampler; 2 * coordinate + draw), [3][2 D] stage-2 draws.
struct ReplayTable {
    const double *base;      // this chain's table (null: keyed Philox uniforms)
    int dim;
    DR_D size_t mut_stride() const { return 4 + 12 * (size_t) dim; }
    DR_D const double *mut_block(uint32_t mut) const { return base + 3 * (size_t) dim + mut * mut_stride(); }
    DR_D Real coin(uint32_t mut, int which) const { return mut_block(mut)[which]; }
    DR_D U4 stage(uint32_t mut, int stage, int sampler, int pair) const {
        U4 u = { NAN, NAN, NAN, NAN };
        if (2 * pair + 1 < dim) {
            const double *q = mut_block(mut) + 4 + (size_t) (stage * 3 + sampler) * 2 * dim + 4 * pair;
            u.x = q[0]; u.y = q[1]; u.z = q[2]; u.w = q[3];
        }
        return u;
    }
};

// What a proposal needs to know about the mutation it belongs to.
struct MutCtx {
    const PssParams *pp;
    uint64_t chain;          // chain id
    uint32_t mut;            // mutation counter of the chain
    ReplayTable table;       // replayed chain: uniforms come from the table
    bool largeStep;
    bool lightTracing;       // nextStage(current->t == 1): the emitter sampler keeps its real stage-2 kernel
    DR_D bool identity1(int s) const { return (pp->identity1 >> s) & 1u; }
    DR_D bool identity2(int s) const {
        if ((pp->identity1 >> s) & 1u) return true;               // setStagesToIdentity (drmlt_proc.cpp:133-136)
        if ((pp->identity2 >> s) & 1u) return !lightTracing;      // handleLightTracing: stage2 = Identity, stageLT = real kernel
        return false;
    }
    // the uniforms of coordinate pair p of sampler s in stage `stage` (0 / 1)
    DR_D U4 draws(int stage, int s, int p) const {
        if (table.base) return table.stage(mut, stage, s, p);
        const float4 f = keyed_uniform4(pp->seed, (stage == 0 ? S_STAGE1 : S_STAGE2) + s, chain, mut, (uint32_t) p);
        U4 u = { (Real) f.x, (Real) f.y, (Real) f.z, (Real) f.w };
        return u;
    }
};
// A replayed chain did not draw what the reference's lazy, touch-driven fill never asked for (a sampler the path did not touch;
// a dimension PSSMLTSampler created in that very step): NaN in the table = the coordinate keeps its current value.
DR_D Real drawn_or(Real candidate, Real draw, Real current) { return isnan(draw) ? current : candidate; }

// un-wrapped stage-1 proposal of the coordinate pair p of sampler s (drmlt_sampler.cpp:313-359, pssmlt_sampler.cpp:124-166)
DR_D R2 propose_stage1(const MutCtx &m, int s, int p, R2 x) {
    const PssParams &pp = *m.pp;
    if (!m.largeStep && m.identity1(s)) return x;
    const U4 u = m.draws(0, s, p);
    if (m.largeStep) return r2(drawn_or(u.x, u.x, x.x), drawn_or(u.z, u.z, x.y));
    if (pp.integrator == DR_INTEGRATOR_PSSMLT)
        return r2(drawn_or(pssmlt_mutate(x.x, u.x, u.y, pp), u.x, x.x), drawn_or(pssmlt_mutate(x.y, u.z, u.w, pp), u.z, x.y));
    if (pp.type != DR_TYPE_ORBITAL)
        return r2(drawn_or(x.x + kelemen_sample(u.x, pp.kel_s2, pp.kel_logRatio), u.x, x.x),
                  drawn_or(x.y + kelemen_sample(u.z, pp.kel_s2, pp.kel_logRatio), u.z, x.y));
    if (isnan(u.x)) return x;
    const Real d = kelemen_sample(u.x, pp.kel_s2, pp.kel_logRatio);   // drmlt_sampler.cpp:351-359
    Real sa, ca;
    sincospi(2.0 * u.y, &sa, &ca);
    return r2(x.x + d * ca, x.y + d * sa);
}
// un-wrapped stage-2 proposal (drmlt_sampler.cpp:313-332, 361-392); y = the stage-1 proposal of the same pair
DR_D R2 propose_stage2(const MutCtx &m, int s, int p, R2 x, R2 y) {
    const PssParams &pp = *m.pp;
    if (!m.largeStep && m.identity2(s)) return x;
    const U4 u = m.draws(1, s, p);
    if (m.largeStep) return r2(drawn_or(u.x, u.x, x.x), drawn_or(u.z, u.z, x.y));   // fillSpace with m_largeStep still set (assertions aside)
    if (pp.type != DR_TYPE_ORBITAL)
        return r2(drawn_or(x.x + gaussian_sample(u.x, u.y, pp.sigma2), u.x, x.x), drawn_or(x.y + gaussian_sample(u.z, u.w, pp.sigma2), u.z, x.y));
    if (isnan(u.x)) return x;
    const Real theta = cauchy_sample(u.x, pp.cauchy_disp);
    const Real du1 = y.x - x.x, du2 = y.y - x.y;
    const Real norm = sqrt(du1 * du1 + du2 * du2);
    Real mu = safe_acos(-du1 / norm);
    if (-du2 < 0.) mu = 2. * R_PI - mu;
    Real sa, ca;
    sincos(theta + mu, &sa, &ca);
    return r2(y.x + ca * norm, y.y + sa * norm);
}

// Sequential reader of the active coordinate buffer of one chain (Sampler::next1D/next2D of the three
// DRMLTSampler / PSSMLTSampler instances).  pos[] survives between wavefront stages in the lane record.
struct UReader {
    const double *buf;        // active buffer of this chain
    int off[3], lim[3];       // first slot / number of readable coordinates per sampler
    int pos[3];
    bool reflect;             // DRMLT stores un-wrapped values and reflects on read
    DR_D Real next1D(int s) {
        const int k = pos[s]++;
        if (k >= lim[s]) return 0.5;                  // the reference raises EError here
        const Real v = buf[off[s] + k];
        return reflect ? wrap_reflect(v) : v;
    }
    DR_D R2 next2D(int s) {
        const int k = pos[s];
        pos[s] = k + 2;
        if (k + 1 >= lim[s] || (k & 1)) {             // odd position or running off the end: scalar path
            pos[s] = k;
            const Real a = next1D(s), b = next1D(s);
            return r2(a, b);
        }
        R2 v = ub_load(buf, off[s] + k);
        if (reflect) { v.x = wrap_reflect(v.x); v.y = wrap_reflect(v.y); }
        return v;
    }
};
