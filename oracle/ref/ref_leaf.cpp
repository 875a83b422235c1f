/* oracle/_ref leaf driver -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * The reference (drmlt-mitsuba) as a whole cannot be built in this image (Boost, Eigen, Xerces-C, OpenEXR are
 * absent).  Its numerical LEAVES on the hot path, however, compile from where they lie
 * under /root/reference once the handful of Boost headers their include closure names are replaced by the
 * small stand-ins in stubs/boost/ (version, static_assert, scoped_ptr, function, bind, algorithm/string: no
 * arithmetic in any of them).  This file is the C entry-point layer around those leaves; the Makefile next
 * to it compiles the reference's own
 *     src/libcore/warp.cpp, util.cpp, triangle.cpp, quad.cpp, math.cpp
 * and the header-only
 *     src/bsdfs/microfacet.h, include/mitsuba/core/pmf.h, include/mitsuba/render/triaccel.h, src/integrators/drmlt/tools/transition.h
 * into oracle/_ref/libref_leaf.so.  No reference source is copied into this repository.
 *
 * The few libcore run-time symbols those translation units reference but never reach on these calls
 * (logging, thread-local lookup) are defined at the bottom as aborting stand-ins.
 */
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/frame.h>
#include <mitsuba/core/warp.h>
#include <mitsuba/core/pmf.h>
#include <mitsuba/core/triangle.h>
#include <mitsuba/core/quad.h>
#include <mitsuba/render/triaccel.h>
#include <deque>

/* transition.h draws from a mitsuba::Random; the product replaces that generator by keyed uniforms (DESIGN.md
 * deviation 1), so the kernels are driven here with a replayed stream: a Random whose nextFloat() pops a queue.
 * (mitsuba/core/random.h is deliberately not included; fwd.h only forward-declares the class.) */
MTS_NAMESPACE_BEGIN
class Random {
public:
    std::deque<Float> q;
    Float nextFloat();
    size_t nextSize(size_t n);
};
Float Random::nextFloat() { if (q.empty()) abort(); Float v = q.front(); q.pop_front(); return v; }
size_t Random::nextSize(size_t) { abort(); }
MTS_NAMESPACE_END

#include "src/bsdfs/microfacet.h"
#include "src/integrators/drmlt/tools/transition.h"

using namespace mitsuba;

extern "C" {

/* libcore/warp.cpp */
void ref_squareToCosineHemisphere(double u, double v, double *out) { Vector d = warp::squareToCosineHemisphere(Point2(u, v)); out[0] = d.x; out[1] = d.y; out[2] = d.z; }
void ref_squareToUniformDiskConcentric(double u, double v, double *out) { Point2 p = warp::squareToUniformDiskConcentric(Point2(u, v)); out[0] = p.x; out[1] = p.y; }
void ref_squareToUniformTriangle(double u, double v, double *out) { Point2 p = warp::squareToUniformTriangle(Point2(u, v)); out[0] = p.x; out[1] = p.y; }
void ref_squareToUniformSphere(double u, double v, double *out) { Vector d = warp::squareToUniformSphere(Point2(u, v)); out[0] = d.x; out[1] = d.y; out[2] = d.z; }

/* libcore/util.cpp */
double ref_fresnelDielectricExt(double cosThetaI, double eta, double *cosThetaT) { Float ct; Float r = fresnelDielectricExt(cosThetaI, ct, eta); *cosThetaT = ct; return r; }
void ref_fresnelConductorExact(double cosThetaI, const double *eta, const double *k, double *out) {
    Spectrum e, kk; for (int i = 0; i < 3; ++i) { e[i] = eta[i]; kk[i] = k[i]; }
    Spectrum r = fresnelConductorExact(cosThetaI, e, kk);
    for (int i = 0; i < 3; ++i) out[i] = r[i];
}
double ref_fresnelDiffuseReflectance(double eta, int fast) { return fresnelDiffuseReflectance(eta, fast != 0); }
void ref_coordinateSystem(const double *a, double *b, double *c) {
    Vector B, Cv; coordinateSystem(Vector(a[0], a[1], a[2]), B, Cv);
    b[0] = B.x; b[1] = B.y; b[2] = B.z; c[0] = Cv.x; c[1] = Cv.y; c[2] = Cv.z;
}
void ref_refract(const double *wi, const double *n, double eta, double *out) {
    Vector r = refract(Vector(wi[0], wi[1], wi[2]), Normal(n[0], n[1], n[2]), eta);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

/* include/mitsuba/core/spectrum.h: luminance of an RGB triple */
double ref_luminance(const double *rgb) { Spectrum s; for (int i = 0; i < 3; ++i) s[i] = rgb[i]; return s.getLuminance(); }

/* include/mitsuba/core/triangle.h:  Triangle::rayIntersect (static, Moeller-Trumbore) and libcore/triangle.cpp: Triangle::sample */
int ref_triangleRayIntersect(const double *p0, const double *p1, const double *p2, const double *o, const double *d, double *uvt) {
    Ray ray(Point(o[0], o[1], o[2]), Vector(d[0], d[1], d[2]), 0);
    Float u, v, t;
    bool hit = Triangle::rayIntersect(Point(p0[0], p0[1], p0[2]), Point(p1[0], p1[1], p1[2]), Point(p2[0], p2[1], p2[2]), ray, u, v, t);
    uvt[0] = u; uvt[1] = v; uvt[2] = t;
    return hit ? 1 : 0;
}
void ref_triangleSample(const double *p0, const double *p1, const double *p2, double u, double v, double *p, double *n) {
    Point pos[3] = { Point(p0[0], p0[1], p0[2]), Point(p1[0], p1[1], p1[2]), Point(p2[0], p2[1], p2[2]) };
    Triangle tri; tri.idx[0] = 0; tri.idx[1] = 1; tri.idx[2] = 2;
    Normal nn;
    Point2 uv;
    Point r = tri.sample(pos, NULL, NULL, nn, uv, Point2(u, v));
    p[0] = r.x; p[1] = r.y; p[2] = r.z; n[0] = nn.x; n[1] = nn.y; n[2] = nn.z;
}

/* include/mitsuba/render/triaccel.h: the kd-tree's projection triangle test (TriAccel::load + rayIntersect) */
int ref_triAccel(const double *p0, const double *p1, const double *p2, const double *o, const double *d, double mint, double maxt, double *uvt) {
    TriAccel ta;
    if (ta.load(Point(p0[0], p0[1], p0[2]), Point(p1[0], p1[1], p1[2]), Point(p2[0], p2[1], p2[2])) != 0) return -1;
    Ray ray(Point(o[0], o[1], o[2]), Vector(d[0], d[1], d[2]), 0);
    Float u = 0, v = 0, t = 0;
    bool hit = ta.rayIntersect(ray, mint, maxt, u, v, t);
    uvt[0] = u; uvt[1] = v; uvt[2] = t;
    return hit ? 1 : 0;
}

/* include/mitsuba/core/pmf.h: DiscreteDistribution append / normalize / sample / sampleReuse */
double ref_pmf(const double *weights, int n, const double *xi, int m, int32_t *index, double *reused, double *pmf_out) {
    DiscreteDistribution dist;
    for (int i = 0; i < n; ++i) dist.append(weights[i]);
    Float sum = dist.normalize();
    for (int i = 0; i < n; ++i) pmf_out[i] = dist[i];
    for (int j = 0; j < m; ++j) {
        Float s = xi[j];
        index[j] = (int32_t) dist.sampleReuse(s);
        reused[j] = s;
    }
    return sum;
}

/* src/bsdfs/microfacet.h: MicrofacetDistribution (type 0 = Beckmann, 1 = GGX) */
void ref_microfacet(int type, double alpha, int sampleVisible, const double *wi, const double *m_in, double u, double v,
                    double *out /* eval(m), pdf(wi,m), G(wi,wo=reflect(wi,m),m), smithG1(wi,m), sample: m.xyz, pdf */) {
    MicrofacetDistribution distr(type == 0 ? MicrofacetDistribution::EBeckmann : MicrofacetDistribution::EGGX, alpha, sampleVisible != 0);
    Vector Wi(wi[0], wi[1], wi[2]);
    Normal M(m_in[0], m_in[1], m_in[2]);
    out[0] = distr.eval(M);
    out[1] = distr.pdf(Wi, M);
    Vector Wo = 2 * dot(Wi, M) * Vector(M) - Wi;
    out[2] = distr.G(Wi, Wo, M);
    out[3] = distr.smithG1(Wi, M);
    Float pdf;
    Normal s = distr.sample(Wi, Point2(u, v), pdf);
    out[4] = s.x; out[5] = s.y; out[6] = s.z; out[7] = pdf;
}

/* src/integrators/drmlt/tools/transition.h */
double ref_kelemen_sample(double s1, double s2, double xi) { Random r; r.q.push_back(xi); return KelemenKernel(s1, s2).sample(&r); }
double ref_kelemen_pdf(double s1, double s2, double du) { return KelemenKernel(s1, s2).pdf(du); }
double ref_kelemen_logpdf(double s1, double s2, double du) { return KelemenKernel(s1, s2).logPdf(du); }
double ref_gaussian_sample(double sigma, double xi1, double xi2) { Random r; r.q.push_back(xi1); r.q.push_back(xi2); return GaussianKernel(sigma).sample(&r); }
double ref_gaussian_logpdf(double sigma, double du) { return GaussianKernel(sigma).logPdf(du); }
double ref_cauchy_sample(double rho, double xi) { Random r; r.q.push_back(xi); return WrappedCauchyKernel(rho).sample(&r); }
double ref_cauchy_pdf(double rho, double du) { return WrappedCauchyKernel(rho).pdf(du); }

}  // extern "C"

/* libcore run-time symbols the leaf translation units name on their error paths only (SLog / Log). */
MTS_NAMESPACE_BEGIN
Thread *Thread::getThread() { return NULL; }
Logger *Thread::getLogger() { return NULL; }
void Logger::log(ELogLevel level, const Class *, const char *file, int line, const char *fmt, ...) {
    fprintf(stderr, "oracle/_ref: the reference logged at level %d from %s:%d: %s\n", (int) level, file, line, fmt);
    if (level >= EError) abort();
}
MTS_NAMESPACE_END
