// k_direct.cu -- the separate direct-illumination image.
//
// With directSamples > 0 (the reference's default is 16) the MLT part excludes paths of depth <= 2 and
// BidirectionalUtils::renderDirectComponent (src/libbidir/util.cpp:30-94) renders them with the `direct` integrator
// (MIDirectIntegrator::Li, src/integrators/direct/direct.cpp:144-305: emitter sampling + BSDF sampling, power
// heuristic) through SamplingIntegrator::renderBlock and the film's reconstruction filter; develop adds the image
// (drmlt_proc.cpp:846-847).  directSamples is split into pixelSamples (<= 8) x shadingSamples (util.cpp:44-54).
// One thread per pixel sample; the traversal is inlined (1 camera ray + shadingSamples shadow rays + shadingSamples
// BSDF rays).  The reference draws from an `ldsampler` seeded from /dev/urandom; sample j of pixel p uses the keyed
// uniforms (S_DIRECT, p, j, .) instead -- the same estimator, reproducible, and identical to the oracle's.
#include "machine.cuh"

namespace {

DR_D bool trace_closest(const DevScene &sc, R3 o, R3 d, Real tmin, Real tmax, Hit &hit) {
    if (tmin == (Real) sc.epsilon) tmin *= fmax(fmax(fmax(fabs(o.x), fabs(o.y)), fabs(o.z)), (Real) sc.epsilon);
    const double rd[8] = { o.x, o.y, o.z, d.x, d.y, d.z, tmin, tmax };
    return traverse<false>(sc, to_f3(o), to_f3(d), (float) tmin, (float) tmax, rd, hit);
}
DR_D bool trace_shadow(const DevScene &sc, R3 o, R3 d, Real tmin, Real tmax) {
    if (tmin == (Real) sc.epsilon) tmin *= fmax(fmax(fmax(fabs(o.x), fabs(o.y)), fabs(o.z)), (Real) sc.epsilon);
    const double rd[8] = { o.x, o.y, o.z, d.x, d.y, d.z, tmin, tmax };
    Hit hit;
    return traverse<true>(sc, to_f3(o), to_f3(d), (float) tmin, (float) tmax, rd, hit);
}
DR_D Real mi(Real a, Real b) { a *= a; b *= b; return a / (a + b); }

} // namespace

// film: (w * rgb, w) per pixel; li (optional): un-filtered radiance of every pixel sample
__global__ void __launch_bounds__(128)
k_direct(const __grid_constant__ DevScene sc, const __grid_constant__ FilmParams fp, unsigned long long seed, int pixelSamples, int shadingSamples,
         float4 *film, double *li) {
    const long long idx = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    const long long nPix = (long long) fp.w * fp.h;
    if (idx >= nPix * pixelSamples) return;
    const unsigned long long p = (unsigned long long) (idx / pixelSamples);
    const uint32_t j = (uint32_t) (idx % pixelSamples);
    const int x = (int) (p % fp.w), y = (int) (p / fp.w);
    auto U = [&](uint32_t k) { return (Real) keyed_uniform(seed, S_DIRECT, p, j, k); };
    const R2 samplePos = r2(x + U(0), y + U(1));
    // PerspectiveCamera::sampleRayDifferential (perspective.cpp:271-298)
    const R3 dl = cam_sample_to_dir(sc.cam, samplePos.x / sc.cam.resX, samplePos.y / sc.cam.resY);
    const Real invZ = 1.0 / dl.z;
    const R3 o = cam_pos(sc.cam), d = cam_xform_dir(sc.cam, dl);
    R3 Li = r3(0.);
    Hit hit;
    if (trace_closest(sc, o, d, sc.cam.nearClip * invZ, sc.cam.farClip * invZ, hit)) {
        Vtx v; Real tHit;
        fill_vertex(sc, hit, o, d, v, tHit);
        if (v.emitter >= 0 && dot(v.ns, -d) > 0.) Li += emitter_radiance(sc, v.emitter);      // its.Le(-ray.d)
        const Mat m = load_material(sc, v.mat, v.uv);
        const R3 wi = to_local(v, -d);
        const R3 refN = mat_transmissive_or_backside(m) ? r3(0.) : v.ns;                      // records.inl:160-164
        const int nE = shadingSamples, nB = shadingSamples;
        const Real fracLum = nE / (Real) (nE + nB), fracBSDF = nB / (Real) (nE + nB), weightLum = 1.0 / nE, weightBSDF = 1.0 / nB;
        if (mat_has_smooth(m.type) && sc.nEmitters > 0) {
            for (int i = 0; i < nE; ++i) {                  // emitter sampling (scene.cpp:879-904, area.cpp:156-170, shape.cpp:102-114)
                EmitterPoint ep;
                sample_emitter_point(sc, U(2 + 2 * i), U(3 + 2 * i), ep);
                R3 dd = ep.p - v.p;
                const Real distSq = dot(dd, dd), dist = sqrt(distSq);
                dd = dd / dist;
                const Real dp = absdot(dd, ep.n);
                Real pdf = sc.emitters[ep.emitter].invArea * (dp != 0. ? distSq / dp : 0.);
                if (!(dot(dd, refN) >= 0. && dot(dd, ep.n) < 0. && pdf != 0.)) continue;
                if (trace_shadow(sc, v.p, dd, sc.epsilon, dist * (1. - sc.shadowEpsilon))) continue;
                const R3 value = emitter_radiance(sc, ep.emitter) / pdf / ep.emPdf;
                pdf *= ep.emPdf;
                const R3 wo = to_local(v, dd);
                const R3 bsdfVal = bsdf_eval(m, wi, wo, MODE_RADIANCE, MEAS_SOLID_ANGLE);
                if (is_zero(bsdfVal)) continue;
                const Real bp = bsdf_pdf(m, wi, wo, MEAS_SOLID_ANGLE);
                Li += value * bsdfVal * (mi(pdf * fracLum, bp * fracBSDF) * weightLum);
            }
        }
        for (int i = 0; i < nB; ++i) {                      // BSDF sampling (direct.cpp:245-300)
            BsdfSample bs;
            bsdf_sample(m, wi, MODE_RADIANCE, U(2 + 2 * (nE + i)), U(3 + 2 * (nE + i)), mat_uses_sampler(m.type) ? U(2 + 4 * nE + i) : 0.5, sc.epsilon, bs);
            if (is_zero(bs.weight)) continue;
            const R3 wo = to_world(v, bs.wo);
            Hit h2;
            if (!trace_closest(sc, v.p, wo, sc.epsilon, INFINITY, h2)) continue;
            Vtx v2; Real t2;
            fill_vertex(sc, h2, v.p, wo, v2, t2);
            if (v2.emitter < 0) continue;
            const R3 value = dot(v2.ns, -wo) > 0. ? emitter_radiance(sc, v2.emitter) : r3(0.);
            Real lumPdf = 0.;
            if (!(bs.sampledType & BT_DELTA) && dot(wo, refN) >= 0. && dot(wo, v2.ns) < 0.) {
                const DevEmitter &em = sc.emitters[v2.emitter];
                lumPdf = em.invArea * (t2 * t2) / absdot(wo, v2.ns) * em.pdfDiscrete;
            }
            Li += value * bs.weight * (mi(bs.pdf * fracBSDF, lumPdf * fracLum) * weightBSDF);
        }
    }
    if (li) { double *out = li + idx * 3; out[0] = Li.x; out[1] = Li.y; out[2] = Li.z; }
    // ImageBlock::put(pos, spec, alpha) (imageblock.h:149-196): weighted value + weight channel
    if (!rgb_valid(Li)) return;
    const float px = (float) samplePos.x - 0.5f, py = (float) samplePos.y - 0.5f;
    const int minx = max((int) ceilf(px - fp.radius), 0), miny = max((int) ceilf(py - fp.radius), 0);
    const int maxx = min((int) floorf(px + fp.radius), fp.w - 1), maxy = min((int) floorf(py + fp.radius), fp.h - 1);
    for (int yy = miny; yy <= maxy; ++yy) {
        const float wy = fp.values[min((int) fabsf((yy - py) * fp.scaleFactor), 31)];
        for (int xx = minx; xx <= maxx; ++xx) {
            const float w = fp.values[min((int) fabsf((xx - px) * fp.scaleFactor), 31)] * wy;
            if (w == 0.f) continue;
            atomicAdd(film + (size_t) yy * fp.w + xx, make_float4(w * (float) Li.x, w * (float) Li.y, w * (float) Li.z, w));
        }
    }
}

// HDRFilm::develop of the weighted film: rgb / weight
__global__ void k_direct_normalize(const float4 *film, long long n, float *rgb) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = film[i];
    const float inv = p.w > 0.f ? 1.0f / p.w : 0.f;
    rgb[3 * i] = p.x * inv; rgb[3 * i + 1] = p.y * inv; rgb[3 * i + 2] = p.z * inv;
}

void launch_direct(const DevScene &sc, const FilmParams &fp, unsigned long long seed, int pixelSamples, int shadingSamples, float4 *film, float *rgb,
                   double *li, cudaStream_t stream) {
    const long long n = (long long) fp.w * fp.h, total = n * pixelSamples;
    k_direct<<<(unsigned) ((total + 127) / 128), 128, 0, stream>>>(sc, fp, seed, pixelSamples, shadingSamples, film, li);
    k_direct_normalize<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(film, n, rgb);
}

// ---------------------------------------------------------------- replay kernel of dr_texture_eval: the BSDF stage's texture lookup on its own
__global__ void k_texture_eval(const __grid_constant__ DevScene sc, uint32_t texture, const double *uv, long long n, double *rgb) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const R3 v = tex_eval(sc, texture, r2(uv[2 * i], uv[2 * i + 1]));
    rgb[3 * i] = v.x; rgb[3 * i + 1] = v.y; rgb[3 * i + 2] = v.z;
}
void launch_texture_eval(const DevScene &sc, uint32_t texture, const double *uv, long long n, double *rgb, cudaStream_t stream) {
    k_texture_eval<<<(unsigned) ((n + 127) / 128), 128, 0, stream>>>(sc, texture, uv, n, rgb);
}
