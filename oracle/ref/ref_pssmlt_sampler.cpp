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
