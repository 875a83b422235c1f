/* oracle/_ref sampler driver -- TEST INFRASTRUCTURE ONLY.
 * The reference's own DRMLT samplers (src/integrators/drmlt/drmlt_sampler.{h,cpp}: Green / Mira / Orbital) compiled into this
 * translation unit from where they lie under /root/reference, driven through their public interface.  The generator is the
 * reference's Random seeded explicitly; a twin generator with the same seed yields the very uniforms the sampler consumes, so
 * that tests can feed them to the oracle restatement in the same order. */
#ifdef REF_NDEBUG   /* the assertion-free twin (see below) needs classes -- and vtables -- of its own */
#define DRMLTSampler DRMLTSamplerN
#define GreenDRMLTSampler GreenDRMLTSamplerN
#define MiraDRMLTSampler MiraDRMLTSamplerN
#define OrbitalDRMLTSampler OrbitalDRMLTSamplerN
#endif
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/random.h>
#include "src/integrators/drmlt/drmlt_sampler.h"
#include "src/integrators/drmlt/drmlt_sampler.cpp"

using namespace mitsuba;

extern "C" void ref_init();     // ref_path.cpp: the start-up sequence of mitsuba.cpp (class table, threads, logger)

#ifndef REF_NDEBUG
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

#endif  /* REF_NDEBUG */

/* ------------------------------------------------------------------------------------------------------------------------
 * WHOLE CHAINS of the reference's DRMLTRenderer::process (drmlt_proc.cpp:386-771; processMixture :161-380), replayable.
 * drmlt_proc.cpp is compiled into this translation unit from where it lies, so the renderer class is visible here; it is
 * driven exactly as a Scheduler worker drives it (prepare -> process on one SeedWorkUnit), but with explicitly seeded
 * generators: seedBoot for the ReplayableSampler the seeds are drawn from (generateSeeds, pathsampler.cpp:859-960), seedWorker
 * for the worker's Random.  Twin generators with the same seeds yield the very uniforms process() consumes -- the seed's replay
 * stream and the worker's stream -- so that a test can feed them, in call order, to the oracle's chain step
 * (orc_mlt.hpp, stream mode) and to the CUDA path (dr_chain_steps `uniforms`).  process() is run for every prefix
 * 0 .. nMut of the chain: the per-prefix ImageBlock (un-normalised doubles) and statistics counters pin every single mutation
 * -- its splat weights a1, (1 - a1) a2, which state was accepted, whether a second stage ran. */
/* The renderer's member functions are inline (vague linkage): under their own names the linker would fold them with the copies
 * in the drmlt plugin object, whose statistics counters are not the ones of this translation unit. */
/* This file is compiled twice: as is, and with -DMTS_NDEBUG -DREF_NDEBUG (ref_drmlt_chain_ndebug).  The reference's build files never
 * define MTS_NDEBUG, so its assertions are live -- and timidAfterLarge=true trips them at the first rejected large step
 * (drmlt_sampler.cpp:320 / :346 SAssert(isFirst); drmlt_proc.cpp:636 SAssert(!largeStep)): as shipped, that flag ends the job with
 * an exception.  The arithmetic behind the assertions (a second-stage fill with m_largeStep still set draws uniforms again) is
 * what the product mirrors; the assertion-free twin makes that arithmetic replayable. */
#ifdef REF_NDEBUG
#define DRMLTRenderer DRMLTRendererReplayN
#define DRMLTProcess DRMLTProcessReplayN
#define ref_drmlt_chain ref_drmlt_chain_ndebug
#else
#define DRMLTRenderer DRMLTRendererReplay
#define DRMLTProcess DRMLTProcessReplay
#endif
#include "src/integrators/drmlt/drmlt_proc.cpp"
#include "drmlt_b200.h"

using namespace mitsuba;

extern "C" void *ref_scene_create(const dr_scene_desc *d, int rfilter);
extern "C" void *ref_scene_ptr(void *h);
extern "C" void ref_scene_destroy(void *h);

namespace {
struct ChainDriver : public DRMLTRenderer {
    ChainDriver(const DRMLTConfiguration &c, const ref_vector<ReplayableSampler> &r) : DRMLTRenderer(c, r) {}
    void bind(const std::string &name, SerializableObject *o) { m_resources[name] = o; }
};
void fillConfig(DRMLTConfiguration &conf, const dr_config *c) {
    conf.technique = c->technique == DR_TECH_MMLT ? PathSampler::EMMLT : c->technique == DR_TECH_BDPT ? PathSampler::EBidirectional : PathSampler::EUnidirectional;
    conf.maxDepth = c->max_depth; conf.rrDepth = c->rr_depth;
    conf.directSampling = c->direct_sampling != 0; conf.directSamples = c->direct_samples; conf.separateDirect = c->direct_samples >= 0;
    conf.luminance = 1.0; conf.luminanceSamples = c->luminance_samples; conf.pLarge = (Float) c->p_large; conf.workUnits = 1;
    conf.nMutations = 0; conf.kelemenStyleWeights = c->kelemen_style_weights != 0; conf.twoStage = false; conf.firstStage = false;
    conf.firstStageSizeReduction = 16; conf.timeout = 0; conf.importanceMap = NULL; conf.averageLuminance = -1.f; conf.lightImage = c->light_image != 0;
    conf.type = c->type == DR_TYPE_GREEN ? DRMLTConfiguration::EGreen : c->type == DR_TYPE_MIRA ? DRMLTConfiguration::EMira : DRMLTConfiguration::EOrbital;
    conf.acceptanceMap = c->acceptance_map != 0; conf.timidAfterLarge = c->timid_after_large != 0; conf.fixEmitterPath = c->fix_emitter_path != 0;
    conf.useMixture = c->use_mixture != 0; conf.sigma = (Float) c->sigma; conf.scaleSecond = (Float) c->scale_second;
}
}

/* counters[k][14]: value / base of firstLevelRatio, largeStepRatio, boldStepRatio, secondLevelRatio, secondLevelLargeRatio,
 * secondLevelBoldRatio, acceptanceRate after k mutations.  films[k] = the work unit's ImageBlock after k mutations (W*H*3). */
extern "C" int ref_drmlt_chain(const dr_scene_desc *d, const dr_config *c, uint64_t seedBoot, uint64_t seedWorker,
                               int nBootSamples, int nSeeds, int pick, int nMut,
                               int32_t *seedDepth, uint64_t *seedSampleIndex, double *seedLuminance,
                               double *bootStream, int nBootStream, double *workerStream, int nWorkerStream,
                               double *films, uint64_t *counters, double *workerNext) {
    try {
        void *h = ref_scene_create(d, c->rfilter);
        if (!h) return 1;
        Scene *scene = (Scene *) ref_scene_ptr(h);
        DRMLTConfiguration conf;
        fillConfig(conf, c);
        PathSeed seed;
        {   /* the seeds, as DRMLT::render draws them (drmlt.cpp:498-546), from an explicitly seeded generator */
            ref<Random> base = new Random(seedBoot);
            ref<ReplayableSampler> rpl = new ReplayableSampler(base);
            ref<PathSampler> ps = new PathSampler(conf.technique, scene, rpl, rpl, rpl, conf.maxDepth, conf.rrDepth, conf.separateDirect,
                                                  conf.directSampling, conf.lightImage);
            std::vector<PathSeed> seeds;
            ps->generateSeeds((size_t) nBootSamples, (size_t) nSeeds, false, NULL, seeds);
            if (seeds.empty()) { ref_scene_destroy(h); return 2; }
            seed = seeds[(size_t) pick % seeds.size()];
            seed.sampler_id = 0;
        }
        *seedDepth = seed.depth; *seedSampleIndex = (uint64_t) seed.sampleIndex; *seedLuminance = seed.luminance;
        {   /* twins: the replay stream from the seed's sample index on, and the worker's stream */
            ref<Random> base = new Random(seedBoot);
            ref<Random> tw = new Random(base);                 // = ReplayableSampler::m_initial (rsampler.cpp:31-33)
            for (size_t i = 0; i < seed.sampleIndex; ++i) tw->nextFloat();
            for (int i = 0; i < nBootStream; ++i) bootStream[i] = tw->nextFloat();
            ref<Random> w = new Random(seedWorker);
            for (int i = 0; i < nWorkerStream; ++i) workerStream[i] = w->nextFloat();
        }
        const Vector2i size = scene->getSensor()->getFilm()->getCropSize();
        const size_t nPix = (size_t) size.x * size.y;
        StatsCounter *ctr[7] = { &firstLevelRatio, &largeStepRatio, &boldStepRatio, &secondLevelRatio, &secondLevelLargeRatio, &secondLevelBoldRatio, &acceptanceRate };
        for (int k = 0; k <= nMut; ++k) {
            conf.nMutations = (size_t) k;
            ref<Random> base = new Random(seedBoot);
            ref_vector<ReplayableSampler> rpls;
            rpls.push_back(new ReplayableSampler(base));
            ref<DRMLTSampler> mlt;
            if (conf.type == DRMLTConfiguration::EGreen) mlt = new GreenDRMLTSampler(conf);
            else if (conf.type == DRMLTConfiguration::EMira) mlt = new MiraDRMLTSampler(conf);
            else mlt = new OrbitalDRMLTSampler(conf);
            ref<Random> worker = new Random(seedWorker);
            mlt->setRandom(worker);
            ref<ChainDriver> wp = new ChainDriver(conf, rpls);
            wp->bind("scene", scene); wp->bind("sensor", scene->getSensor()); wp->bind("sampler", mlt);
            wp->prepare();
            ref<WorkUnit> wu = wp->createWorkUnit();
            ref<WorkResult> wr = wp->createWorkResult();
            static_cast<SeedWorkUnit *>(wu.get())->setSeed(seed);
            static_cast<SeedWorkUnit *>(wu.get())->setTimeout(0);
            for (int i = 0; i < 7; ++i) ctr[i]->reset();
            const bool stop = false;
            wp->process(wu, wr, stop);
            const ImageBlock *block = static_cast<const ImageBlock *>(wr.get());
            const Bitmap *bmp = block->getBitmap();
            const int border = block->getBorderSize();
            const Float *data = bmp->getFloatData();
            const int ch = bmp->getChannelCount(), bw = bmp->getWidth();
            double *out = films + (size_t) k * nPix * 3;
            for (int y = 0; y < size.y; ++y)
                for (int x = 0; x < size.x; ++x) {
                    const Float *p = data + ((size_t) (y + border) * bw + (x + border)) * ch;
                    Spectrum s; for (int q = 0; q < SPECTRUM_SAMPLES; ++q) s[q] = p[q];
                    Float R, G, B; s.toLinearRGB(R, G, B);
                    double *o = out + ((size_t) y * size.x + x) * 3;
                    o[0] = R; o[1] = G; o[2] = B;
                }
            for (int i = 0; i < 7; ++i) { counters[(size_t) k * 14 + 2 * i] = ctr[i]->getValue(); counters[(size_t) k * 14 + 2 * i + 1] = ctr[i]->getBase(); }
            if (k == nMut) *workerNext = worker->nextFloat();     // how many uniforms the whole chain consumed: the next one tells
        }
        ref_scene_destroy(h);
        return 0;
    } catch (const std::exception &e) { fprintf(stderr, "oracle/_ref: %s\n", e.what()); return 1; }
}
