/* oracle/_ref sampler driver -- TEST INFRASTRUCTURE ONLY.
 * The reference's own DRMLT samplers (src/integrators/drmlt/drmlt_sampler.{h,cpp}: Green / Mira / Orbital) compiled into this
 * translation unit from where they lie under /root/reference, driven through their public interface.  The generator is the
 * reference's Random seeded explicitly; a twin generator with the same seed yields the very uniforms the sampler consumes, so
 * that tests can feed them to the oracle restatement in the same order. */
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/random.h>
#include "src/integrators/drmlt/drmlt_sampler.h"
#include "src/integrators/drmlt/drmlt_sampler.cpp"

using namespace mitsuba;

extern "C" void ref_init();     // ref_path.cpp: the start-up sequence of mitsuba.cpp (class table, threads, logger)

extern "C" int ref_drmlt_sampler(int type /* dr_type: 0 green, 1 mira, 2 orbital */, int maxDim, double sigma, double scaleSecond,
                                 int largeStep, uint64_t seed, double *uCurrent, double *stream, int nStream,
                                 double *prop1, double *prop2, double *reverse, double *ratio) {
    try {
        ref_init();
        DRMLTConfiguration conf;
        conf.type = type == 0 ? DRMLTConfiguration::EGreen : type == 1 ? DRMLTConfiguration::EMira : DRMLTConfiguration::EOrbital;
        conf.sigma = sigma;
        conf.scaleSecond = scaleSecond;
        ref<DRMLTSampler> s;
        if (type == 0) s = new GreenDRMLTSampler(conf);
        else if (type == 1) s = new MiraDRMLTSampler(conf);
        else s = new OrbitalDRMLTSampler(conf);
        ref<Random> rA = new Random(seed), rB = new Random(seed);
        s->setRandom(rA);
        s->setMaxDim((size_t) maxDim);
        /* the current state: the seed-replay protocol of DRMLTRenderer::process (drmlt_proc.cpp:467-504) */
        s->setReplay(true);
        for (int k = 0; k < maxDim; ++k) uCurrent[k] = s->primarySample((size_t) k);
        s->accept(true);
        s->setReplay(false);
        for (int k = 0; k < maxDim; ++k) rB->nextFloat();
        for (int j = 0; j < nStream; ++j) stream[j] = rB->nextFloat();
        s->setLargeStep(largeStep != 0);
        for (int k = 0; k < maxDim; ++k) prop1[k] = s->primarySample((size_t) k);
        s->nextStage();
        s->setLargeStep(false);                       // what timidAfterLarge does before a second stage
        for (int k = 0; k < maxDim; ++k) prop2[k] = s->primarySample((size_t) k);
        if (type == 0) { s->setReverse(true); for (int k = 0; k < maxDim; ++k) reverse[k] = s->primarySample((size_t) k); }
        *ratio = s->getTransitionRatio(0.3);
        return 0;
    } catch (const std::exception &e) { fprintf(stderr, "oracle/_ref: %s\n", e.what()); return 1; }
}

/* A SEQUENCE of mutations through the same samplers, driven the way DRMLTRenderer::process drives them (drmlt_proc.cpp:541-760):
 * per mutation setLargeStep -> stage-1 proposal; outcome 0: accept(true); otherwise nextStage(lightTracing) -> setLargeStep(false)
 * -> stage-2 proposal [-> Green's reverse state] [-> Mira's ratio]; outcome 1: accept(false); outcome 2: reject().
 * mode 1 = handleLightTracing() (the emitter sampler under fixEmitterPath: stage 2 is the identity unless the stage is a
 * light-tracing one), mode 2 = setStagesToIdentity() (the MMLT direct sampler).  NaN marks what a mutation did not produce. */
extern "C" int ref_drmlt_sampler_seq(int type, int mode, int maxDim, double sigma, double scaleSecond, uint64_t seed, int nMut,
                                     const int *large, const int *outcome, const int *lightTracing,
                                     double *uCurrent, double *stream, int nStream,
                                     double *prop1, double *prop2, double *reverse, double *ratio /* [nMut](*maxDim) */) {
    try {
        ref_init();
        DRMLTConfiguration conf;
        conf.type = type == 0 ? DRMLTConfiguration::EGreen : type == 1 ? DRMLTConfiguration::EMira : DRMLTConfiguration::EOrbital;
        conf.sigma = sigma;
        conf.scaleSecond = scaleSecond;
        ref<DRMLTSampler> s;
        if (type == 0) s = new GreenDRMLTSampler(conf);
        else if (type == 1) s = new MiraDRMLTSampler(conf);
        else s = new OrbitalDRMLTSampler(conf);
        if (mode == 1) s->handleLightTracing();
        if (mode == 2) s->setStagesToIdentity();
        ref<Random> rA = new Random(seed), rB = new Random(seed);
        s->setRandom(rA);
        s->setMaxDim((size_t) maxDim);
        s->setReplay(true);
        for (int k = 0; k < maxDim; ++k) uCurrent[k] = s->primarySample((size_t) k);
        s->accept(true);
        s->setReplay(false);
        for (int k = 0; k < maxDim; ++k) rB->nextFloat();
        for (int j = 0; j < nStream; ++j) stream[j] = rB->nextFloat();
        const double nan = std::numeric_limits<double>::quiet_NaN();
        for (int m = 0; m < nMut; ++m) {
            double *p1 = prop1 + (size_t) m * maxDim, *p2 = prop2 + (size_t) m * maxDim, *rv = reverse + (size_t) m * maxDim;
            for (int k = 0; k < maxDim; ++k) p1[k] = p2[k] = rv[k] = nan;
            ratio[m] = nan;
            s->setLargeStep(large[m] != 0);
            for (int k = 0; k < maxDim; ++k) p1[k] = s->primarySample((size_t) k);
            if (outcome[m] == 0) { s->accept(true); continue; }
            s->nextStage(lightTracing[m] != 0);
            s->setLargeStep(false);
            for (int k = 0; k < maxDim; ++k) p2[k] = s->primarySample((size_t) k);
            if (type == 0) {
                s->setReverse(true);
                for (int k = 0; k < maxDim; ++k) rv[k] = s->primarySample((size_t) k);
                s->setReverse(false);
            }
            ratio[m] = s->getTransitionRatio(0.3);
            if (outcome[m] == 1) s->accept(false); else s->reject();
        }
        return 0;
    } catch (const std::exception &e) { fprintf(stderr, "oracle/_ref: %s\n", e.what()); return 1; }
}
