/* oracle/_ref PSSMLT sampler driver -- TEST INFRASTRUCTURE ONLY.
 * The reference's own PSSMLTSampler (src/integrators/pssmlt/pssmlt_sampler.{h,cpp}) compiled into this translation unit from
 * where it lies under /root/reference and driven through its public interface over a SEQUENCE of mutations: seed replay
 * (pssmlt_proc.cpp:134-150), then per mutation setLargeStep -> primarySample(0..maxDim-1) -> accept / reject, so that the
 * eager fill, the Kelemen / Gaussian mutation and the backup / restore logic are all exercised.  The generator is the
 * reference's Random seeded explicitly; a twin generator with the same seed yields the uniforms the sampler consumes. */
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/random.h>
#include "src/integrators/pssmlt/pssmlt_sampler.h"
#include "src/integrators/pssmlt/pssmlt_sampler.cpp"

using namespace mitsuba;

extern "C" void ref_init();     // ref_path.cpp

extern "C" int ref_pssmlt_sampler(int kelemen, int maxDim, double s1, double s2, double sigma, uint64_t seed, int nMut,
                                  const int *large, const int *accepted, double *uCurrent, double *stream, int nStream,
                                  double *proposals /* [nMut][maxDim] */) {
    try {
        ref_init();
        PSSMLTConfiguration conf;
        conf.mutationSizeLow = s1;
        conf.mutationSizeHigh = s2;
        conf.sigma = sigma;
        ref<PSSMLTSampler> s = new PSSMLTSampler(conf);
        s->setMutationType(kelemen != 0);
        ref<Random> rA = new Random(seed), rB = new Random(seed);
        s->setRandom(rA);
        s->setMaxDim((size_t) maxDim);
        s->setReplay(true);
        for (int k = 0; k < maxDim; ++k) uCurrent[k] = s->primarySample((size_t) k);
        s->accept();
        s->setReplay(false);
        for (int k = 0; k < maxDim; ++k) rB->nextFloat();
        for (int j = 0; j < nStream; ++j) stream[j] = rB->nextFloat();
        for (int m = 0; m < nMut; ++m) {
            s->setLargeStep(large[m] != 0);
            for (int k = 0; k < maxDim; ++k) proposals[(size_t) m * maxDim + k] = s->primarySample((size_t) k);
            if (accepted[m]) s->accept(); else s->reject();
        }
        return 0;
    } catch (const std::exception &e) { fprintf(stderr, "oracle/_ref: %s\n", e.what()); return 1; }
}

/* ------------------------------------------------------------------------------------------------------------------------
 * WHOLE CHAINS of the reference's PSSMLTRenderer::process (pssmlt_proc.cpp:110-285), replayable: the same protocol as
 * ref_drmlt_chain (ref_sampler.cpp) -- pssmlt_proc.cpp compiled into this translation unit, prepare -> process on one
 * SeedWorkUnit with explicitly seeded generators, twin generators for the two uniform streams, one run per prefix 0 .. nMut. */
#define PSSMLTRenderer PSSMLTRendererReplay
#define PSSMLTProcess PSSMLTProcessReplay
#include "src/integrators/pssmlt/pssmlt_proc.cpp"
#include "drmlt_b200.h"

using namespace mitsuba;

extern "C" void *ref_scene_create(const dr_scene_desc *d, int rfilter);
extern "C" void *ref_scene_ptr(void *h);
extern "C" void ref_scene_destroy(void *h);

namespace {
struct PssChainDriver : public PSSMLTRenderer {
    PssChainDriver(const PSSMLTConfiguration &c, const ref_vector<ReplayableSampler> &r) : PSSMLTRenderer(c, r) {}
    void bind(const std::string &name, SerializableObject *o) { m_resources[name] = o; }
};
}

/* counters[k][6]: value / base of largeStepRatio, smallStepRatio, acceptanceRate after k mutations; b = m_config.luminance */
extern "C" int ref_pssmlt_chain(const dr_scene_desc *d, const dr_config *c, double b, uint64_t seedBoot, uint64_t seedWorker,
                                int nBootSamples, int nSeeds, int pick, int nMut,
                                int32_t *seedDepth, uint64_t *seedSampleIndex, double *seedLuminance,
                                double *bootStream, int nBootStream, double *workerStream, int nWorkerStream,
                                double *films, uint64_t *counters, double *workerNext) {
    try {
        void *h = ref_scene_create(d, c->rfilter);
        if (!h) return 1;
        Scene *scene = (Scene *) ref_scene_ptr(h);
        PSSMLTConfiguration conf;
        conf.technique = c->technique == DR_TECH_MMLT ? PathSampler::EMMLT : c->technique == DR_TECH_BDPT ? PathSampler::EBidirectional : PathSampler::EUnidirectional;
        conf.maxDepth = c->max_depth; conf.rrDepth = c->rr_depth;
        conf.directSampling = c->direct_sampling != 0; conf.directSamples = c->direct_samples; conf.separateDirect = c->direct_samples >= 0;
        conf.luminance = (Float) b; conf.luminanceSamples = c->luminance_samples; conf.pLarge = (Float) c->p_large; conf.workUnits = 1;
        conf.nMutations = 0; conf.kelemenStyleWeights = c->kelemen_style_weights != 0; conf.twoStage = false; conf.firstStage = false;
        conf.firstStageSizeReduction = 16; conf.timeout = 0; conf.importanceMap = NULL; conf.averageLuminance = -1.f; conf.lightImage = c->light_image != 0;
        conf.kelemenStyleMutation = c->kelemen_style_mutation != 0;
        conf.mutationSizeLow = (Float) c->mutation_size_low; conf.mutationSizeHigh = (Float) c->mutation_size_high; conf.sigma = (Float) c->sigma;
        PathSeed seed;
        {
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
        {
            ref<Random> base = new Random(seedBoot);
            ref<Random> tw = new Random(base);
            for (size_t i = 0; i < seed.sampleIndex; ++i) tw->nextFloat();
            for (int i = 0; i < nBootStream; ++i) bootStream[i] = tw->nextFloat();
            ref<Random> w = new Random(seedWorker);
            for (int i = 0; i < nWorkerStream; ++i) workerStream[i] = w->nextFloat();
        }
        const Vector2i size = scene->getSensor()->getFilm()->getCropSize();
        const size_t nPix = (size_t) size.x * size.y;
        StatsCounter *ctr[3] = { &largeStepRatio, &smallStepRatio, &acceptanceRate };
        for (int k = 0; k <= nMut; ++k) {
            conf.nMutations = (size_t) k;
            ref<Random> base = new Random(seedBoot);
            ref_vector<ReplayableSampler> rpls;
            rpls.push_back(new ReplayableSampler(base));
            ref<PSSMLTSampler> mlt = new PSSMLTSampler(conf);
            ref<Random> worker = new Random(seedWorker);
            mlt->setRandom(worker);
            ref<PssChainDriver> wp = new PssChainDriver(conf, rpls);
            wp->bind("scene", scene); wp->bind("sensor", scene->getSensor()); wp->bind("sampler", mlt);
            wp->prepare();
            ref<WorkUnit> wu = wp->createWorkUnit();
            ref<WorkResult> wr = wp->createWorkResult();
            static_cast<SeedWorkUnit *>(wu.get())->setSeed(seed);
            static_cast<SeedWorkUnit *>(wu.get())->setTimeout(0);
            for (int i = 0; i < 3; ++i) ctr[i]->reset();
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
            for (int i = 0; i < 3; ++i) { counters[(size_t) k * 6 + 2 * i] = ctr[i]->getValue(); counters[(size_t) k * 6 + 2 * i + 1] = ctr[i]->getBase(); }
            if (k == nMut) *workerNext = worker->nextFloat();
        }
        ref_scene_destroy(h);
        return 0;
    } catch (const std::exception &e) { fprintf(stderr, "oracle/_ref: %s\n", e.what()); return 1; }
}
