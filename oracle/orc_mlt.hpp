// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
//
// orc_mlt.hpp: primary-sample-space samplers, transition kernels, the PSSMLT and DRMLT chain
// loops, bootstrap, splatting and develop.  Restates
//   src/integrators/drmlt/tools/transition.h:23-190         (Gaussian/Kelemen/Identity/WrappedCauchy)
//   src/integrators/drmlt/drmlt_sampler.{h,cpp}             (DRMLTSampler + Green/Mira/Orbital)
//   src/integrators/drmlt/drmlt_proc.cpp:161-380, 386-771   (processMixture, process), :813-854 (develop)
//   src/integrators/pssmlt/pssmlt_sampler.{h,cpp}, pssmlt_proc.cpp:110-285, :326-365
//   src/integrators/pssmlt_utils.h:27-77                    (findMaxDimensions)
//   src/libbidir/pathsampler.cpp:859-960                    (generateSeeds)
//   include/mitsuba/render/imageblock.h:149-196, src/libcore/rfilter.cpp:37-55 (film splat)
//
// Uniform addressing: every uniform is addressed by KEY (orc_rng.hpp), so a sampler's whole
// proposal vector is a pure function of (chain, mutation, stage, sampler, coordinate).  The
// reference fills a sampler's vector lazily, at its first primarySample(0) of a stage
// (drmlt_sampler.cpp:252-254), from one sequential SFMT stream; with keyed uniforms lazy and
// eager fills give the same vectors, and a sampler the path never touches cannot influence
// the chain, so the restatement fills eagerly.
#pragma once
#include "orc_sampler.hpp"
#include "orc_rng.hpp"
#include <memory>
#include <functional>

namespace orc {

// ------------------------------------------------------------------ findMaxDimensions (pssmlt_utils.h:27-77)
// No media in scope: offsetMedium = 0.  hasRoughDielectric: some shape's BSDF is a RoughDielectric (:35-45).
struct MaxDim { int sensor, emitter, direct; };
inline MaxDim findMaxDimensions(int maxDepth, int rrDepth, int depth, int technique, bool useDirectSampling, bool hasRoughDielectric) {
    int offsetRR = (rrDepth < maxDepth ? 1 : 0) + (hasRoughDielectric ? 1 : 0);
    if (technique == DR_TECH_MMLT) {
        int maxDim = (depth + 2) * 3;
        if (maxDim % 2 == 1) maxDim++;
        return MaxDim{ maxDim, maxDim, 1 };
    } else if (technique == DR_TECH_PATH) {
        int maxDim = (maxDepth + 2) * (4 + offsetRR);
        if (maxDim % 2 == 1) maxDim++;
        return MaxDim{ maxDim, 0, 0 };
    } else {
        int maxDim = (maxDepth + 2) * (2 + offsetRR);
        if (maxDim % 2 == 1) maxDim++;
        return MaxDim{ maxDim, maxDim, useDirectSampling ? maxDepth : 0 };
    }
}

// ------------------------------------------------------------------ transition kernels (transition.h)
struct KelemenKernel {
    Float s1, s2, logRatio;
    KelemenKernel(Float a, Float b) : s1(a), s2(b) { logRatio = -std::log(b / a); }
    Float sample(Float xi) const {   // transition.h:96-110
        int sign;
        if (xi < 0.5) { sign = 1; xi *= 2.0; } else { sign = -1; xi = 2.0 * (xi - 0.5); }
        Float dv = s2 * std::exp((1 - xi) * logRatio);
        return dv * sign;
    }
    Float pdf(Float du) const {      // :112-117
        Float d = std::abs(du);
        if (d < s1 || d > s2) return 0.0;
        return 1.0 / (2.0 * d * (-logRatio));
    }
    Float logPdf(Float du) const { return std::log(pdf(du)); }
};
struct GaussianKernel {
    Float sigma;
    Float sample(Float xi1, Float xi2) const {   // :63-68 Box-Muller
        Float tmp = std::sqrt(-2.0 * std::log(1 - xi1));
        return tmp * std::cos(2.0 * PI * xi2) * sigma;
    }
};
struct WrappedCauchyKernel {
    Float rho, dispersion;
    explicit WrappedCauchyKernel(Float r) : rho(r), dispersion(2.0 * r / (1.0 + r * r)) {}
    Float sample(Float xi) const {   // :162-178
        int sign = 1;
        if (xi < 0.5) { sign = 1; xi *= 2.0; } else { sign = -1; xi = 2.0 * (xi - 0.5); }
        Float V = std::cos(2.0 * PI * xi);
        Float angle = (V + dispersion) / (1.0 + dispersion * V);
        return sign * safe_acos(angle);
    }
};

// Float-precision uniform source addressed by key.  `mut` is the mutation index of the chain.
//
// Replay table (whole chains of the reference on the CUDA path, dr_chain_replay): a sequential stream says nothing about WHICH
// uniform a value is; the keyed address does.  While a chain runs in sequential mode every value it consumes is also written to
// `table` at its keyed address, and a chain can run FROM such a table instead of Philox (tableIn).  Layout (doubles, NaN = never
// drawn), D = tableDim:  [3][D] the replayed seed state per sampler | per mutation: [4] coins, [3][2 D] stage-1 draws (sampler,
// 2 * coordinate + draw), [3][2 D] stage-2 draws.
struct KeyedSource {
    uint64_t seed = 0, chain = 0;
    uint32_t mut = 0;
    // sequential mode (tests/test_ref_pins.py): the uniforms come, in call order, from a recorded stream of the
    // reference's own generator, so that the arithmetic can be held against the reference draw by draw
    const double *seq = nullptr;
    mutable size_t seqPos = 0;
    double *table = nullptr;             // written in sequential mode
    const double *tableIn = nullptr;     // read instead of Philox
    int tableDim = 0;
    size_t tableMutStride() const { return 4 + 12 * (size_t) tableDim; }
    size_t coinAddr(int which) const { return 3 * (size_t) tableDim + mut * tableMutStride() + (size_t) which; }
    size_t stageAddr(int stageIdx, int sampler, int coord, int draw) const {
        return 3 * (size_t) tableDim + mut * tableMutStride() + 4 + ((size_t) (stageIdx * 3 + sampler) * 2 * tableDim) + (size_t) (2 * coord + draw);
    }
    Float coin(int which) const {
        if (seq) { const double v = seq[seqPos++]; if (table) table[coinAddr(which)] = v; return (Float) v; }
        if (tableIn) return (Float) tableIn[coinAddr(which)];
        return (Float) keyedUniform(seed, S_COIN, chain, mut, (uint32_t) which);
    }
    Float stage(int stageIdx, int sampler, int coord, int draw) const {
        if (seq) { const double v = seq[seqPos++]; if (table && coord < tableDim) table[stageAddr(stageIdx, sampler, coord, draw)] = v; return (Float) v; }
        if (tableIn) return (Float) tableIn[stageAddr(stageIdx, sampler, coord, draw)];
        return keyedUniform(seed, (stageIdx == 0 ? S_STAGE1 : S_STAGE2) + sampler, chain, mut, (uint32_t) (2 * coord + draw));
    }
    // the reference fills a proposal lazily, at the sampler's first query of a stage: sequential and table runs must not draw
    // (or read) anything for a sampler the path never touches
    bool lazy() const { return seq || tableIn; }
    // a value that becomes part of the chain's STATE without being a mutation draw (fillReplay, drmlt_sampler.h:127-131; a new
    // dimension of PSSMLTSampler, pssmlt_sampler.cpp:137-139): recorded in the table's seed-state section
    Float fresh(int sampler, int coord) const {
        const double v = seq[seqPos++];
        if (table && coord < tableDim) table[(size_t) sampler * tableDim + coord] = v;
        return (Float) v;
    }
    // table mode: a NaN draw = this coordinate is not mutated in this step
    static bool missing(Float v) { return std::isnan(v); }
    static Float boot(uint64_t seed, uint64_t index, int sampler, int coord) {
        return keyedUniform(seed, S_BOOT, index, (uint32_t) sampler, (uint32_t) coord);
    }
};

// drmlt_sampler.h:140-144
inline Float wrapReflect(Float y) { return y > 1 ? 2.0 - y : (y <= 0 ? std::abs(y) : y); }

enum EKernel { K_KELEMEN, K_GAUSSIAN, K_ORBITAL, K_IDENTITY };

// DRMLTSampler and subclasses (drmlt_sampler.{h,cpp}), eager fill, keyed uniforms.
struct DRMLTSampler : Sampler {
    int type;                 // dr_type
    int samplerId;            // 0 sensor, 1 emitter, 2 direct
    const KeyedSource *src = nullptr;
    size_t maxDim = 0;
    bool largeStep = false;
    bool isFirst = true, isLightTracing = false, isReverse = false;
    bool identityAll = false;        // setStagesToIdentity(): MMLT direct sampler (drmlt_proc.cpp:133-136)
    bool stage2Identity = false;     // handleLightTracing(): fixEmitterPath on the emitter sampler (:137-140)
    Float sigma, scaleSecond;
    const Float s1 = 1.0 / 1024.0, s2 = 1.0 / 64.0, kelemenScale = 1.9f;   // drmlt_sampler.h:201-205
    std::vector<Float> uCurrent, uProp1, uProp2;
    size_t sampleIndex = 0, dimStage1 = 0, dimStage2 = 0;
    bool filled1 = false, filled2 = false;
    bool arrayMode = false;          // replay of an explicit vector (seed replay): read uCurrent directly
    // stream mode (whole chains of the reference replayed, tests/test_ref_pins.py): seed replay pulls the three samplers'
    // values, in call order, from ONE interleaved stream (drmlt_sampler.cpp:245-249, drmlt_proc.cpp:470-485)
    const double *replaySeq = nullptr;
    size_t *replayPos = nullptr;

    bool kernelIsIdentity() const {
        if (identityAll) return true;
        if (isFirst) return false;
        if (stage2Identity) return !isLightTracing;   // stage2 = Identity, stageLT = real kernel
        return false;
    }
    void setLargeStep(bool v) { largeStep = v; }
    void nextStage(bool lightTracing = false) { sampleIndex = 0; isFirst = false; isLightTracing = lightTracing; }
    void setReverse(bool v) { sampleIndex = 0; isReverse = v; }
    void resetStage() {
        sampleIndex = 0; isFirst = true; isReverse = false; isLightTracing = false;
        uProp1.clear(); uProp2.clear(); filled1 = filled2 = false; dimStage1 = dimStage2 = 0;
    }
    void accept(bool acceptFirst) {   // drmlt_sampler.cpp:189-199
        // a sampler the path never touched holds an EMPTY proposal, which the reference copies over its current state (SURVEY
        // C.14) -- values that nothing reads before the next large step refills them; a sequential stream must not be advanced
        // for it, so the (unobservable) current state is simply kept
        if (src && src->lazy() && !(acceptFirst ? filled1 : filled2)) { resetStage(); return; }
        ensureFilled(acceptFirst);
        uCurrent = acceptFirst ? uProp1 : uProp2;
        uCurrent.resize(maxDim);      // orbital with odd maxDim pushes one extra coordinate
        for (Float &v : uCurrent) v = wrapReflect(v);
        resetStage();
    }
    void reject() { resetStage(); }

    // fillSpace (drmlt_sampler.cpp:313-394)
    void fill(bool first) {
        std::vector<Float> &uProposed = first ? uProp1 : uProp2;
        uProposed.clear();
        const int st = first ? 0 : 1;
        KelemenKernel kel(type == DR_TYPE_ORBITAL ? s1 * kelemenScale : s1, type == DR_TYPE_ORBITAL ? s2 * kelemenScale : s2);
        GaussianKernel gauss{ scaleSecond * sigma };
        WrappedCauchyKernel cauchy(std::exp(-0.25f));   // a FLOAT exponential: m_rho = std::exp(-0.25f) (drmlt_sampler.h:204)
        bool savedFirst = isFirst; isFirst = first;
        const bool identity = kernelIsIdentity();
        isFirst = savedFirst;
        for (size_t i = 0; i < maxDim; i++) {
            if (largeStep) {
                uProposed.push_back(src->stage(st, samplerId, (int) i, 0));
            } else if (identity) {
                uProposed.push_back(uCurrent[i]);
            } else if (type != DR_TYPE_ORBITAL) {
                if (first) uProposed.push_back(uCurrent[i] + kel.sample(src->stage(st, samplerId, (int) i, 0)));
                else {
                    const Float xi1 = src->stage(st, samplerId, (int) i, 0), xi2 = src->stage(st, samplerId, (int) i, 1);   // in this order (transition.h:63-66)
                    uProposed.push_back(uCurrent[i] + gauss.sample(xi1, xi2));
                }
            } else if (first) {   // orbital first stage: 2-D radial Kelemen (:351-359)
                Float d = kel.sample(src->stage(st, samplerId, (int) i, 0));
                Float a = src->stage(st, samplerId, (int) i, 1) * 2.0 * PI;
                Float x1 = i + 1 < uCurrent.size() ? uCurrent[i + 1] : 0.0;
                uProposed.push_back(uCurrent[i] + d * std::cos(a));
                uProposed.push_back(x1 + d * std::sin(a));
                i++;
            } else {              // orbital second stage (:361-392)
                Float theta_i = cauchy.sample(src->stage(st, samplerId, (int) i, 0));
                Float x1 = i + 1 < uCurrent.size() ? uCurrent[i + 1] : 0.0;
                Float du1 = uProp1[i] - uCurrent[i];
                Float du2 = uProp1[i + 1] - x1;
                Float norm = std::sqrt(du1 * du1 + du2 * du2);
                Float mu_i = safe_acos(-du1 / norm);
                if (-du2 < 0) mu_i = 2.0 * PI - mu_i;
                Float c1 = uProp1[i] + std::cos(theta_i + mu_i) * norm;
                Float c2 = uProp1[i + 1] + std::sin(theta_i + mu_i) * norm;
                uProposed.push_back(c1);
                uProposed.push_back(c2);
                i++;
            }
        }
        (first ? filled1 : filled2) = true;
    }
    void ensureFilled(bool first) {
        if (first ? !filled1 : !filled2) {
            // keyed uniforms: a second stage can always look at the first-stage proposal.  The reference's lazy fill never draws a
            // first-stage vector on behalf of a second stage (a sampler first touched in stage 2 -- a strategy change by a large-step
            // second stage, timidAfterLarge -- is filled with uniforms that need none)
            if (!first && !filled1 && !(src && src->lazy())) fill(true);
            fill(first);
        }
    }
    // primarySample (drmlt_sampler.cpp:231-307)
    Float primarySample(size_t k) {
        if (arrayMode) return k < uCurrent.size() ? wrapReflect(uCurrent[k]) : 0.5;
        if (replaySeq) {                 // m_replay: one draw per call, kept as the state (drmlt_sampler.cpp:245-249)
            if (k == 0) uCurrent.clear();
            uCurrent.push_back((Float) replaySeq[(*replayPos)++]);
            return wrapReflect(uCurrent[k]);
        }
        (isFirst ? dimStage1 : dimStage2) = std::max(k, isFirst ? dimStage1 : dimStage2);
        if (type == DR_TYPE_GREEN && isReverse) {
            ensureFilled(true); ensureFilled(false);
            Float du = uProp1[k] - uCurrent[k];
            return wrapReflect(uProp2[k] - du);
        }
        ensureFilled(isFirst);
        const std::vector<Float> &uProposed = isFirst ? uProp1 : uProp2;
        if (k >= uProposed.size()) return 0.5;   // the reference logs EError here (:257-259)
        return wrapReflect(uProposed[k]);
    }
    Float next1D() override { return primarySample(sampleIndex++); }

    // MiraDRMLTSampler::getTransitionRatio (drmlt_sampler.cpp:400-414).  dimStage* hold the largest
    // INDEX touched, so the last used coordinate is skipped exactly as in the reference.
    Float getTransitionRatio() {
        if (identityAll) return 1.0;   // stage1->isIdentity()
        if (src && src->lazy() && std::max(dimStage1, dimStage2) == 0) return 1.0;   // untouched: empty sums, no draws
        ensureFilled(true); ensureFilled(false);
        KelemenKernel kel(s1, s2);
        size_t dimStage = std::max(dimStage1, dimStage2);
        Float num = 0.0, denum = 0.0;
        for (size_t i = 0; i < dimStage; i++) {
            num += kel.logPdf(uProp2[i] - uProp1[i]);
            denum += kel.logPdf(uCurrent[i] - uProp1[i]);
        }
        return std::exp(num - denum);
    }
};

// PSSMLTSampler (pssmlt_sampler.{h,cpp}), eager fill, keyed uniforms.
struct PSSMLTSampler : Sampler {
    int samplerId;
    const KeyedSource *src = nullptr;
    size_t maxDim = 0;
    bool largeStep = false, useKelemen = true;
    Float s1, s2, logRatio, sigma;
    std::vector<Float> u, backup;
    size_t sampleIndex = 0;
    bool filled = false, arrayMode = false;
    void configure(Float s1_, Float s2_, Float sigma_) { s1 = s1_; s2 = s2_; sigma = sigma_; logRatio = -std::log(s2 / s1); }
    Float mutate(Float value, int coord) const {   // pssmlt_sampler.h:117-147
        if (src->tableIn && KeyedSource::missing(src->stage(0, samplerId, coord, 0))) return value;   // not mutated in the recorded step
        if (useKelemen) {
            Float sample = src->stage(0, samplerId, coord, 0);
            bool add;
            if (sample < 0.5) { add = true; sample *= 2.0; } else { add = false; sample = 2.0 * (sample - 0.5); }
            Float dv = s2 * std::exp(sample * logRatio);
            if (add) { value += dv; if (value > 1) value -= 1; }
            else { value -= dv; if (value < 0) value += 1; }
        } else {
            Float tmp1 = std::sqrt(-2 * std::log(1 - src->stage(0, samplerId, coord, 0)));
            Float dv = tmp1 * std::cos(2 * PI * src->stage(0, samplerId, coord, 1));
            Float r = std::fmod(value + sigma * dv, 1.0);
            value = (r < 0.0) ? r + 1.0 : r;
        }
        return value;
    }
    void setLargeStep(bool v) { largeStep = v; }
    const double *replaySeq = nullptr;   // stream mode: seed replay from ONE interleaved stream (pssmlt_sampler.cpp:126-129)
    size_t *replayPos = nullptr;
    void ensureFilled() {   // primarySample(0) (pssmlt_sampler.cpp:124-166)
        if (filled) return;
        backup = u;
        const size_t have = u.size();    // keyed uniforms: the state is padded up front, have == maxDim
        for (size_t k = 0; k < maxDim; k++) {
            if (k >= have) u.push_back(src->fresh(samplerId, (int) k));      // a NEW dimension: a fresh value, not a mutation (:137-139)
            else if (largeStep) { const Float v = src->stage(0, samplerId, (int) k, 0); if (!(src->tableIn && KeyedSource::missing(v))) u[k] = v; }
            else u[k] = mutate(u[k], (int) k);
        }
        filled = true;
    }
    Float primarySample(size_t i) {
        if (arrayMode) return i < u.size() ? u[i] : 0.5;
        if (replaySeq) { u.push_back((Float) replaySeq[(*replayPos)++]); return u[i]; }
        ensureFilled();
        return i < u.size() ? u[i] : 0.5;
    }
    Float next1D() override { return primarySample(sampleIndex++); }
    void accept() { if (!(src && src->lazy())) ensureFilled(); backup.clear(); sampleIndex = 0; filled = false; }
    void reject() {          // dimensions created by this step were never backed up: they stay (pssmlt_sampler.cpp:108-113)
        if (filled) for (size_t k = 0; k < backup.size(); ++k) u[k] = backup[k];
        backup.clear(); sampleIndex = 0; filled = false;
    }
};

// ------------------------------------------------------------------ film (imageblock.h:149-196)
struct Film {
    int w, h;
    Float radius, scaleFactor;
    Float values[32];
    std::vector<Float> data;   // w*h*3
    // the filter plugins' eval functions with their default parameters (src/rfilters/*.cpp) -> radius + 32-entry table (rfilter.cpp:37-55)
    static Float evalFilter(int rfilter, Float x, Float radius) {
        auto cubic = [](Float x, Float B, Float C) {   // mitchell.cpp:55-69, catmullrom.cpp:43-58
            x = std::abs(x);
            Float x2 = x * x, x3 = x2 * x;
            if (x < 1) return 1.0f / 6.0f * ((12 - 9 * B - 6 * C) * x3 + (-18 + 12 * B + 6 * C) * x2 + (6 - 2 * B));
            else if (x < 2) return 1.0f / 6.0f * ((-B - 6 * C) * x3 + (6 * B + 30 * C) * x2 + (-12 * B - 48 * C) * x + (8 * B + 24 * C));
            return (Float) 0.0;
        };
        switch (rfilter) {
        case DR_FILTER_BOX: return std::abs(x) <= radius ? 1.0 : 0.0;                              // box.cpp:43-45
        case DR_FILTER_TENT: return std::max((Float) 0.0, 1.0f - std::abs(x / radius));               // tent.cpp:43-45
        case DR_FILTER_MITCHELL: return cubic(x, 1.0f / 3.0f, 1.0f / 3.0f);
        case DR_FILTER_CATMULLROM: return cubic(x, 0.0f, 0.5f);
        case DR_FILTER_LANCZOS: {                                                                  // lanczos.cpp:44-57
            x = std::abs(x);
            if (x < 1e-7) return 1.0;
            else if (x > radius) return 0.0;
            Float x1 = M_PI * x, x2 = x1 / radius;
            return (std::sin(x1) * std::sin(x2)) / (x1 * x2);
        }
        default: {                                                                                 // gaussian.cpp:52-57, stddev 0.5
            const Float stddev = 0.5, alpha = -1.0 / (2.0 * stddev * stddev);
            return std::max((Float) 0.0, std::exp(alpha * x * x) - std::exp(alpha * radius * radius));
        }
        }
    }
    void init(int w_, int h_, int rfilter) {
        w = w_; h = h_; data.assign((size_t) w * h * 3, 0.0);
        switch (rfilter) {
        case DR_FILTER_BOX: radius = 0.5 + (Float) 1e-5f; break;      // a FLOAT literal added to Float 0.5 (box.cpp:38)
        case DR_FILTER_TENT: radius = 1.0; break;
        case DR_FILTER_LANCZOS: radius = 3.0; break;                  // lobes = 3
        default: radius = 2.0; break;                                 // gaussian: 4 stddev; mitchell, catmullrom
        }
        Float sum = 0.0;
        for (int i = 0; i < 31; ++i) {
            Float v = evalFilter(rfilter, (radius * i) / 31, radius);
            values[i] = v; sum += v;
        }
        values[31] = 0.0;
        scaleFactor = 31 / radius;
        sum *= 2 * radius / 31;
        const Float normalization = 1.0 / sum;          // multiplied in, as rfilter.cpp:52-54
        for (int i = 0; i < 31; ++i) values[i] *= normalization;
    }
    // a job's film: the configuration's filter, or its explicit table (DR_FILTER_TABLE)
    void init(int w_, int h_, const dr_config &c) {
        if (c.rfilter != DR_FILTER_TABLE) { init(w_, h_, c.rfilter); return; }
        w = w_; h = h_; data.assign((size_t) w * h * 3, 0.0);
        radius = c.filter_radius; scaleFactor = 31 / radius;
        for (int i = 0; i < 32; ++i) values[i] = c.filter_table[i];
    }
    Float evalDiscretized(Float x) const { return values[std::min((int) std::abs(x * scaleFactor), 31)]; }
    bool put(const Vec2 &_pos, const RGB &value) {
        if (!value.isValid()) return false;
        const Float px = _pos.x - 0.5, py = _pos.y - 0.5;
        const int minx = std::max((int) std::ceil(px - radius), 0), miny = std::max((int) std::ceil(py - radius), 0);
        const int maxx = std::min((int) std::floor(px + radius), w - 1), maxy = std::min((int) std::floor(py + radius), h - 1);
        for (int y = miny; y <= maxy; ++y) {
            const Float wy = evalDiscretized(y - py);
            for (int x = minx; x <= maxx; ++x) {
                const Float wgt = evalDiscretized(x - px) * wy;
                Float *dest = &data[((size_t) y * w + x) * 3];
                dest[0] += wgt * value.r; dest[1] += wgt * value.g; dest[2] += wgt * value.b;
            }
        }
        return true;
    }
};

struct ChainStats {
    uint64_t mutations = 0, first_accept = 0, first_base = 0, large_accept = 0, large_base = 0, bold_accept = 0, bold_base = 0,
             second_accept = 0, second_base = 0, second_large_accept = 0, second_large_base = 0,
             second_bold_accept = 0, second_bold_base = 0, accept = 0, accept_base = 0, paths = 0, rays = 0;
    void add(const ChainStats &o) {
        mutations += o.mutations; first_accept += o.first_accept; first_base += o.first_base;
        large_accept += o.large_accept; large_base += o.large_base; bold_accept += o.bold_accept; bold_base += o.bold_base;
        second_accept += o.second_accept; second_base += o.second_base;
        second_large_accept += o.second_large_accept; second_large_base += o.second_large_base;
        second_bold_accept += o.second_bold_accept; second_bold_base += o.second_bold_base;
        accept += o.accept; accept_base += o.accept_base; paths += o.paths; rays += o.rays;
    }
};

inline PathSamplerConfig pathConfigOf(const dr_config &c) {
    PathSamplerConfig p;
    p.technique = c.technique; p.maxDepth = c.max_depth; p.rrDepth = c.rr_depth;
    p.excludeDirectIllum = c.direct_samples >= 0;   // separateDirect (drmlt.cpp:233)
    p.lightImage = c.light_image != 0;
    return p;
}

// One bootstrap path sample `index` (generateSeeds body, pathsampler.cpp:879-920): all three
// samplers replay the BOOT stream; MMLT depth = (index % maxDepth) + 1.
struct BootSampler : Sampler {
    uint64_t seed, index; int samplerId; int pos = 0;
    Float next1D() override { return KeyedSource::boot(seed, index, samplerId, pos++); }
};
inline int bootstrapDepth(const dr_config &c, uint64_t index) {
    return c.technique == DR_TECH_MMLT ? (int) (index % (uint64_t) c.max_depth) + 1 : -1;
}
inline void bootstrapSample(const Scene &sc, const dr_config &c, uint64_t index, SplatList &list, uint64_t *rays = nullptr) {
    BootSampler se, em, di;
    se.seed = em.seed = di.seed = c.seed; se.index = em.index = di.index = index;
    se.samplerId = 0; em.samplerId = 1; di.samplerId = 2;
    PathSampler ps(&sc, pathConfigOf(c), &em, &se, &di);
    ps.sampleSplats(list, bootstrapDepth(c, index));
    if (rays) *rays += ps.ctx.rays;
}

struct StepRecord { Float L_x, L_y, L_z, a1, a2; bool large, acc1, did2, acc2; };

// ------------------------------------------------------------------ DRMLTRenderer::process (+ processMixture)
struct ChainRunner {
    const Scene &sc;
    dr_config cfg;
    Float b;                    // m_config.luminance
    Film *film;                 // may be null
    ChainStats stats;
    const float *imp = nullptr; // m_config.importanceMap (two-stage MLT), at the job's image size
    int impW = 0, impH = 0;
    ChainRunner(const Scene &s, const dr_config &c, Float b_, Film *f) : sc(s), cfg(c), b(b_), film(f) {
        if (c.importance_map && !c.first_stage) { imp = c.importance_map; impW = (int) s.cam.resX; impH = (int) s.cam.resY; }
    }
    void norm(SplatList &l) const { l.normalize(imp, impW, impH); }

    static bool invalidStrict(Float x) { return std::isnan(x) || std::isinf(x) || x <= 0; }   // drmlt_proc.cpp:428
    static bool invalidLoose(Float x) { return std::isnan(x) || std::isinf(x) || x < 0; }     // :181
    static Float metropolisClamp(Float x) { return std::min((Float) 1.0, x); }                // :425 (NaN -> 1)

    void splat(const SplatList &l, Float weight) {
        if (!film || cfg.acceptance_map || !(weight > 0)) return;
        for (size_t k = 0; k < l.size(); ++k) {
            RGB value = l.splats[k].second * weight;
            if (value.isValid()) film->put(l.splats[k].first, value);
        }
    }
    void splatAcceptanceOnly(const SplatList &l, int stage) {   // drmlt_proc.cpp:443-450
        if (!film || !cfg.acceptance_map) return;
        for (size_t k = 0; k < l.size(); ++k)
            film->put(l.splats[k].first, stage == 0 ? RGB(1, 0, 0) : RGB(0, 1, 0));
    }

    // chain `chainId`, seeded from bootstrap sample `seedIndex`; records (optional) has nMutations entries
    // Stream mode: the seed's replay stream and the worker's stream of a recorded chain of the reference (oracle/ref/ref_sampler.cpp);
    // every uniform is then consumed in the reference's call order.  streamUsed reports the worker uniforms consumed.
    const double *bootStream = nullptr, *workerStream = nullptr;
    size_t streamUsed = 0;
    // replay table (KeyedSource): written while a stream is replayed / read instead of Philox
    double *tableOut = nullptr;
    const double *tableIn = nullptr;
    int tableDim = 0;

    void runDRMLT(uint64_t chainId, uint64_t seedIndex, int depth, uint64_t nMutations, StepRecord *records) {
        KeyedSource src; src.seed = cfg.seed; src.chain = chainId;
        src.seq = workerStream; src.table = tableOut; src.tableIn = tableIn; src.tableDim = tableDim;
        MaxDim md = findMaxDimensions(cfg.max_depth, cfg.rr_depth, depth, cfg.technique, cfg.direct_sampling != 0, sc.hasRoughDielectric);
        DRMLTSampler sensorS, emitterS, directS;
        DRMLTSampler *all[3] = { &sensorS, &emitterS, &directS };
        size_t dims[3] = { (size_t) md.sensor, (size_t) md.emitter, (size_t) md.direct };
        for (int i = 0; i < 3; ++i) {
            DRMLTSampler &s = *all[i];
            s.type = cfg.type; s.samplerId = i; s.src = &src; s.maxDim = dims[i];
            s.sigma = cfg.sigma; s.scaleSecond = cfg.scale_second;
            // seed replay + fillReplay (drmlt_proc.cpp:467-504): current = BOOT vector of the seed
            s.uCurrent.resize(dims[i]);
            for (size_t k = 0; k < dims[i]; ++k)
                s.uCurrent[k] = tableIn ? (Float) tableIn[(size_t) i * tableDim + k] : KeyedSource::boot(cfg.seed, seedIndex, i, (int) k);
        }
        if (cfg.technique == DR_TECH_MMLT) {
            directS.identityAll = true;
            if (cfg.fix_emitter_path) emitterS.stage2Identity = true;
        }
        PathSampler ps(&sc, pathConfigOf(cfg), &emitterS, &sensorS, &directS);
        SplatList current, prop1, prop2, reverse;
        size_t bootPos = 0;
        if (bootStream) for (auto s : all) { s->uCurrent.clear(); s->replaySeq = bootStream; s->replayPos = &bootPos; s->sampleIndex = 0; }
        else for (auto s : all) { s->arrayMode = true; s->sampleIndex = 0; }
        ps.sampleSplats(current, depth);
        for (auto s : all) { s->arrayMode = false; s->replaySeq = nullptr; s->resetStage(); }
        if (bootStream)                  // accept(true) wraps the replayed values, fillReplay pads with the worker's uniforms (drmlt_proc.cpp:493-504)
            for (auto s : all) {
                for (Float &v : s->uCurrent) v = wrapReflect(v);
                while (s->uCurrent.size() < s->maxDim) s->uCurrent.push_back(src.fresh(s->samplerId, (int) s->uCurrent.size()));
                if (src.table) for (size_t k = 0; k < s->uCurrent.size() && (int) k < src.tableDim; ++k) src.table[(size_t) s->samplerId * src.tableDim + k] = s->uCurrent[k];
            }
        ++stats.paths;
        if (cfg.acceptance_map) {}   // luminance override only affects develop
        norm(current);

        const bool mixture = cfg.use_mixture != 0;
        for (uint64_t m = 0; m < nMutations; ++m) {
            src.mut = (uint32_t) m;
            StepRecord rec = { current.luminance, 0, 0, 0, 0, false, false, false, false };
            Float a1 = 0, a2 = 0; bool acc1 = false, acc2 = false;
            bool largeStep = src.coin(0) < cfg.p_large;
            for (auto s : all) s->setLargeStep(largeStep);
            ps.sampleSplats(prop1, depth);
            norm(prop1);
            ++stats.mutations; ++stats.paths;
            auto flipCoin = [&](Float x, int which) { return (x >= 1) || (src.coin(which) < x); };

            if (mixture) {   // processMixture (drmlt_proc.cpp:161-380)
                Float a = 0; bool accept = false;
                if (!invalidLoose(prop1.luminance)) { a = metropolisClamp(prop1.luminance / current.luminance); accept = flipCoin(a, 1); }
                bool doSecond = false;
                if (!largeStep) doSecond = flipCoin(0.5, 3);
                SplatList *proposed = &prop1;
                if (doSecond) {
                    sensorS.nextStage(); directS.nextStage();
                    if (cfg.fix_emitter_path) emitterS.nextStage(current.t == 1); else emitterS.nextStage();
                    ps.sampleSplats(prop2, depth); norm(prop2); ++stats.paths;
                    proposed = &prop2;
                    if (invalidLoose(prop2.luminance)) { a = 0; accept = false; }
                    else { a = metropolisClamp(prop2.luminance / current.luminance); accept = flipCoin(a, 2); }
                }
                splat(current, 1.0 - a);
                splat(*proposed, a);
                rec.large = largeStep; rec.L_y = prop1.luminance; rec.L_z = doSecond ? prop2.luminance : 0;
                rec.did2 = doSecond; rec.a1 = doSecond ? 0 : a; rec.a2 = doSecond ? a : 0;
                rec.acc1 = accept && !doSecond; rec.acc2 = accept && doSecond;
                stats.accept_base++;
                if (!doSecond) { stats.first_base++; if (largeStep) stats.large_base++; else stats.bold_base++; }
                else stats.second_base++;
                if (accept) {
                    stats.accept++;
                    if (!doSecond) { stats.first_accept++; if (largeStep) stats.large_accept++; else stats.bold_accept++; }
                    else stats.second_accept++;
                    std::swap(current, *proposed);
                    for (auto s : all) s->accept(!doSecond);
                } else {
                    for (auto s : all) s->reject();
                }
                if (records) records[m] = rec;
                continue;
            }

            // first stage (drmlt_proc.cpp:543-550)
            if (!invalidStrict(prop1.luminance)) { a1 = metropolisClamp(prop1.luminance / current.luminance); acc1 = flipCoin(a1, 1); }
            bool doSecond = !acc1;
            if (!cfg.timid_after_large) doSecond = doSecond && !largeStep;
            prop2.clear();
            if (doSecond) {
                sensorS.nextStage(); directS.nextStage();
                if (cfg.fix_emitter_path) emitterS.nextStage(current.t == 1); else emitterS.nextStage();
                ps.sampleSplats(prop2, depth); norm(prop2); ++stats.paths;
                if (!invalidStrict(prop2.luminance)) {
                    if (cfg.type == DR_TYPE_GREEN) {   // :588-621
                        for (auto s : all) s->setReverse(true);
                        ps.sampleSplats(reverse, depth); norm(reverse); ++stats.paths;
                        Float aReverse = invalidStrict(reverse.luminance) ? 0.0 : metropolisClamp(reverse.luminance / prop2.luminance);
                        if (aReverse == 1) { a2 = 0; acc2 = false; }
                        else {
                            Float lumRatio = prop2.luminance / current.luminance;
                            a2 = metropolisClamp(lumRatio * (1.0 - aReverse) / (1.0 - a1));
                            acc2 = flipCoin(a2, 2);
                        }
                        for (auto s : all) s->setReverse(false);
                    } else if (cfg.type == DR_TYPE_MIRA) {   // :625-650
                        Float aReverse = metropolisClamp(prop1.luminance / prop2.luminance);
                        if (aReverse >= 1) { a2 = 0; acc2 = false; }
                        else {
                            Float transitionRatio = largeStep ? 1.0 :
                                sensorS.getTransitionRatio() * emitterS.getTransitionRatio() * directS.getTransitionRatio();
                            if (invalidStrict(transitionRatio)) { a2 = 0; acc2 = false; }
                            else {
                                Float lumRatio = prop2.luminance / current.luminance;
                                a2 = metropolisClamp(lumRatio * transitionRatio * (1.0 - aReverse) / (1.0 - a1));
                                acc2 = flipCoin(a2, 2);
                            }
                        }
                    } else {   // orbital :655-669
                        if (prop2.luminance < prop1.luminance) { a2 = 0; acc2 = false; }
                        else if (prop2.luminance >= current.luminance) { a2 = 1.0; acc2 = true; }
                        else {
                            a2 = (prop2.luminance - prop1.luminance) / (current.luminance - prop1.luminance);
                            acc2 = flipCoin(a2, 2);
                        }
                    }
                }
            }
            // expectation weights (:676-688)
            Float w1 = a1, w2 = (1.0 - a1) * a2, wc = 1.0 - w1 - w2;
            splat(current, wc); splat(prop1, w1); splat(prop2, w2);
            rec.large = largeStep; rec.L_y = prop1.luminance; rec.L_z = doSecond ? prop2.luminance : 0;
            rec.a1 = a1; rec.a2 = a2; rec.acc1 = acc1; rec.did2 = doSecond; rec.acc2 = acc2;
            if (records) records[m] = rec;

            if (acc1 || acc2) {   // :691-742
                // NB: after the swap proposed.* holds the OLD current state, and that is what the reference
                // hands to splatAcceptanceOnly (drmlt_proc.cpp:695-708): the map is binned at the state left.
                if (acc1) { std::swap(prop1, current); if (!largeStep) splatAcceptanceOnly(prop1, 0); }
                else { std::swap(prop2, current); splatAcceptanceOnly(prop2, 1); }
                for (auto s : all) s->accept(acc1);
                stats.accept_base++; stats.accept++;
                if (acc1) {
                    stats.first_base++; stats.first_accept++;
                    if (largeStep) { stats.large_base++; stats.large_accept++; } else { stats.bold_base++; stats.bold_accept++; }
                } else {
                    stats.accept_base++; stats.first_base++; stats.second_base++; stats.second_accept++;
                    if (largeStep) { stats.large_base++; stats.second_large_base++; stats.second_large_accept++; }
                    else { stats.bold_base++; stats.second_bold_base++; stats.second_bold_accept++; }
                }
            } else {              // :746-769
                for (auto s : all) s->reject();
                stats.accept_base++; stats.first_base++;
                if (largeStep) { stats.large_base++; if (doSecond) { stats.second_base++; stats.second_large_base++; stats.accept_base++; } }
                else { stats.bold_base++; if (doSecond) { stats.second_base++; stats.second_bold_base++; stats.accept_base++; } }
            }
        }
        stats.rays += ps.ctx.rays;
        streamUsed = src.seqPos;
    }

    // PSSMLTRenderer::process (pssmlt_proc.cpp:110-285)
    void runPSSMLT(uint64_t chainId, uint64_t seedIndex, int depth, uint64_t nMutations, StepRecord *records) {
        KeyedSource src; src.seed = cfg.seed; src.chain = chainId;
        src.seq = workerStream; src.table = tableOut; src.tableIn = tableIn; src.tableDim = tableDim; src.table = tableOut; src.tableIn = tableIn; src.tableDim = tableDim;
        MaxDim md = findMaxDimensions(cfg.max_depth, cfg.rr_depth, depth, cfg.technique, cfg.direct_sampling != 0, sc.hasRoughDielectric);
        PSSMLTSampler sensorS, emitterS, directS;
        PSSMLTSampler *all[3] = { &sensorS, &emitterS, &directS };
        size_t dims[3] = { (size_t) md.sensor, (size_t) md.emitter, (size_t) md.direct };
        for (int i = 0; i < 3; ++i) {
            PSSMLTSampler &s = *all[i];
            s.samplerId = i; s.src = &src; s.maxDim = dims[i]; s.useKelemen = cfg.kelemen_style_mutation != 0;
            s.configure(cfg.mutation_size_low, cfg.mutation_size_high, cfg.sigma);
            s.u.resize(dims[i]);
            for (size_t k = 0; k < dims[i]; ++k)
                s.u[k] = tableIn ? (Float) tableIn[(size_t) i * tableDim + k] : KeyedSource::boot(cfg.seed, seedIndex, i, (int) k);
        }
        PathSampler ps(&sc, pathConfigOf(cfg), &emitterS, &sensorS, &directS);
        SplatList current, proposed;
        size_t bootPos = 0;
        if (bootStream) for (auto s : all) { s->u.clear(); s->replaySeq = bootStream; s->replayPos = &bootPos; s->sampleIndex = 0; }
        else for (auto s : all) { s->arrayMode = true; s->sampleIndex = 0; }
        ps.sampleSplats(current, depth);
        for (auto s : all) { s->arrayMode = false; s->replaySeq = nullptr; s->sampleIndex = 0; }
        if (bootStream && tableOut)      // the replayed part of the seed state (new dimensions are added by KeyedSource::fresh)
            for (auto s : all) for (size_t k = 0; k < s->u.size() && (int) k < tableDim; ++k) tableOut[(size_t) s->samplerId * tableDim + k] = s->u[k];
        ++stats.paths;
        norm(current);
        Float cumulativeWeight = 0;
        const bool kelemenW = cfg.kelemen_style_weights != 0 && !imp;   // pssmlt_proc.cpp:205
        for (uint64_t m = 0; m < nMutations; ++m) {
            src.mut = (uint32_t) m;
            bool largeStep = src.coin(0) < cfg.p_large;
            for (auto s : all) s->setLargeStep(largeStep);
            ps.sampleSplats(proposed, depth);
            norm(proposed);
            ++stats.mutations; ++stats.paths;
            Float a = std::min((Float) 1.0, proposed.luminance / current.luminance);
            if (std::isnan(proposed.luminance) || proposed.luminance < 0) a = 0;
            bool accept; Float currentWeight, proposedWeight;
            if (a > 0) {
                if (kelemenW) {
                    currentWeight = (1 - a) * current.luminance / (current.luminance / b + cfg.p_large);
                    proposedWeight = (a + (largeStep ? 1 : 0)) * proposed.luminance / (proposed.luminance / b + cfg.p_large);
                } else { currentWeight = 1 - a; proposedWeight = a; }
                accept = (a == 1) || (src.coin(1) < a);
            } else {
                currentWeight = kelemenW ? current.luminance / (current.luminance / b + cfg.p_large) : 1;
                proposedWeight = 0; accept = false;
            }
            cumulativeWeight += currentWeight;
            StepRecord rec = { current.luminance, proposed.luminance, 0, a, 0, largeStep, accept, false, false };
            if (records) records[m] = rec;
            stats.accept_base++;
            if (largeStep) stats.large_base++; else stats.bold_base++;
            if (accept) {
                splatNonZero(current, cumulativeWeight);
                cumulativeWeight = proposedWeight;
                std::swap(proposed, current);
                for (auto s : all) s->accept();
                stats.accept++;
                if (largeStep) stats.large_accept++; else stats.bold_accept++;
            } else {
                splatNonZero(proposed, proposedWeight);
                for (auto s : all) s->reject();
            }
        }
        splatNonZero(current, cumulativeWeight);
        stats.rays += ps.ctx.rays;
        streamUsed = src.seqPos;
    }
    void splatNonZero(const SplatList &l, Float weight) {   // pssmlt_proc.cpp:229-233
        if (!film) return;
        for (size_t k = 0; k < l.size(); ++k) {
            RGB value = l.splats[k].second * weight;
            if (!value.isZero()) film->put(l.splats[k].first, value);
        }
    }
    void run(uint64_t chainId, uint64_t seedIndex, int depth, uint64_t nMutations, StepRecord *records) {
        if (cfg.integrator == DR_INTEGRATOR_PSSMLT) runPSSMLT(chainId, seedIndex, depth, nMutations, records);
        else runDRMLT(chainId, seedIndex, depth, nMutations, records);
    }
};

// develop (drmlt_proc.cpp:813-854): image = accum * (b / mean pixel luminance)
inline void develop(const Film &film, Float b, bool acceptanceMap, float *out, const float *importanceMap = nullptr) {
    size_t n = (size_t) film.w * film.h;
    Float avg = 0;
    for (size_t i = 0; i < n; ++i) {
        Float l = RGB(film.data[3 * i], film.data[3 * i + 1], film.data[3 * i + 2]).luminance();
        avg += importanceMap ? l * importanceMap[i] : l;      // :825-830
    }
    avg /= (Float) n;
    Float factor = acceptanceMap ? 1.0 : b / avg;
    for (size_t i = 0; i < n; ++i) {
        Float correction = importanceMap ? factor * importanceMap[i] : factor;   // :841-844
        for (int c = 0; c < 3; ++c) out[3 * i + c] = (float) (film.data[3 * i + c] * correction);
    }
}

// ------------------------------------------------------------------ two-stage MLT: importance map
// Resampler (include/mitsuba/core/rfilter.h:107-324) in resampling mode with the gaussian filter (gaussian.cpp:30-60),
// EClamp boundary (:437-458) and resampleAndClamp to [0, inf) (:232-280).
struct Resampler {
    int sourceRes, targetRes, taps;
    std::vector<int> start;
    std::vector<Float> weights;
    static Float gaussianEval(Float x) {
        const Float stddev = 0.5, radius = 4 * stddev, alpha = -1.0 / (2.0 * stddev * stddev);
        return std::max((Float) 0.0, std::exp(alpha * x * x) - std::exp(alpha * radius * radius));
    }
    Resampler(int src, int tgt) : sourceRes(src), targetRes(tgt) {
        Float filterRadius = 2.0, scale = 1.0, invScale = 1.0;
        if (targetRes < sourceRes) { scale = (Float) sourceRes / (Float) targetRes; invScale = 1 / scale; filterRadius *= scale; }
        taps = (int) std::ceil(filterRadius * 2);
        start.resize(targetRes); weights.resize((size_t) taps * targetRes);
        for (int i = 0; i < targetRes; i++) {
            Float center = (i + 0.5) / targetRes * sourceRes;
            start[i] = (int) std::floor(center - filterRadius + 0.5);
            Float sum = 0;
            for (int j = 0; j < taps; j++) {
                Float pos = start[i] + j + 0.5 - center;
                Float weight = gaussianEval(pos * invScale);
                weights[(size_t) i * taps + j] = weight;
                sum += weight;
            }
            Float normalization = 1.0 / sum;
            for (int j = 0; j < taps; j++) weights[(size_t) i * taps + j] *= normalization;
        }
    }
    // one line: source[stride * k] -> target[tstride * i]
    void resampleAndClamp(const Float *source, size_t stride, Float *target, size_t tstride) const {
        for (int i = 0; i < targetRes; ++i) {
            Float result = 0;
            for (int j = 0; j < taps; ++j) {
                int pos = std::min(std::max(start[i] + j, 0), sourceRes - 1);
                result += source[stride * pos] * weights[(size_t) i * taps + j];
            }
            target[tstride * i] = std::max((Float) 0.0, result);
        }
    }
};
// mltLuminancePass, last part (util.cpp:180-196): developed RGB -> luminance -> Bitmap::resample (bitmap.cpp:2230-2329)
// Bitmap::resample of a luminance bitmap (bitmap.cpp:2230-2329): along x first (into a [h][W] temporary), then along y
inline std::vector<Float> resampleMap(const std::vector<Float> &lum, int w, int h, int W, int H) {
    std::vector<Float> tmp, out;
    const Float *cur = lum.data();
    int curW = w;
    if (w != W) {
        Resampler r(w, W);
        tmp.resize((size_t) W * h);
        for (int y = 0; y < h; ++y) r.resampleAndClamp(cur + (size_t) y * w, 1, tmp.data() + (size_t) y * W, 1);
        cur = tmp.data(); curW = W;
    }
    if (h != H) {
        Resampler r(h, H);
        out.resize((size_t) W * H);
        for (int x = 0; x < curW; ++x) r.resampleAndClamp(cur + x, curW, out.data() + x, curW);
        cur = out.data();
    }
    return std::vector<Float>(cur, cur + (size_t) W * H);
}
inline void resampleLuminance(const float *rgb, int w, int h, int W, int H, float *map) {
    std::vector<Float> lum((size_t) w * h);
    for (size_t i = 0; i < lum.size(); ++i) lum[i] = RGB(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]).luminance();
    const std::vector<Float> out = resampleMap(lum, w, h, W, H);
    for (size_t i = 0; i < (size_t) W * H; ++i) map[i] = (float) out[i];
}

} // namespace orc
