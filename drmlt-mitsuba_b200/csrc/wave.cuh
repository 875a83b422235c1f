// wave.cuh -- wavefront execution of the Markov chains.
//
// Every chain ("lane") always has at most ONE ray in flight.  A round is two kernels:
//   k_logic : one thread per lane.  Consumes the lane's last hit, advances the resumable path
//             evaluator (MMLT connection strategy or unidirectional path tracer) until it needs the
//             next ray, and -- whenever a path completes -- runs the chain-level step fused in the same
//             thread: delayed-rejection acceptance (mira / green / orbital / mixture, or PSSMLT),
//             expectation-weighted film splats, commit of the accepted primary-sample vector,
//             statistics, and the mutation of the next proposal.  Double-precision arithmetic.
//   k_trace : one thread per lane, float32 BVH traversal only (closest hit or any hit); small
//             register footprint, so many warps are resident to hide the node/triangle gather latency.
// Lanes progress at their own pace (no lock-step over mutations): a lane whose first stage was
// accepted simply starts its next mutation while its neighbour traces a second-stage path.
// The same machine runs three kinds of jobs: Markov chains (JOB_CHAIN), the bootstrap
// (JOB_BOOT, PathSampler::generateSeeds) and replayed primary-sample vectors (JOB_EVAL, parity).
//
// Behavioural parity targets:
//   PathSampler::sampleSplats                src/libbidir/pathsampler.cpp:79-571 (MMLT :84-320, PT :529-567)
//   MIPathTracer::Li                         src/integrators/path/path.cpp:123-312
//   DRMLTRenderer::process / processMixture  src/integrators/drmlt/drmlt_proc.cpp:161-380, 386-771
//   PSSMLTRenderer::process                  src/integrators/pssmlt/pssmlt_proc.cpp:110-285
#pragma once
#include "path.cuh"

// ------------------------------------------------------------------ lane memory (SoA, 32-bit words)
enum { RAY_NONE = 0, RAY_CLOSEST = 1, RAY_SHADOW = 2 };
enum { PS_IDLE = 0, PS_START, PS_SENSOR_HIT, PS_EMITTER_START, PS_EMITTER_HIT, PS_CONNECT, PS_CONNECT_SHADOW, PS_FINISH,
       PS_PT_HIT, PS_PT_SHADE, PS_PT_NEE, PS_PT_BSDF };
enum { PH_STAGE1 = 0, PH_STAGE2 = 1, PH_REVERSE = 2, PH_INIT = 3 };
enum { F_DELTA = 1u, F_ANYCONN = 2u, F_SPOS_FAIL = 4u, F_PT_FIRST = 8u, F_PT_EMITTED = 16u, F_PT_DIRECT = 32u, F_PT_NONSPEC = 64u };

struct PathCore {             // 26 words
    int pstate, s, t, j;
    uint32_t flags, connectable, pssPos, pssMax;
    int nrays, pad;           // rays cast for the path in flight
    R3 weight;                // MMLT: product of walk weights, then the connection value | PT: throughput
    R3 d;                     // direction of the ray in flight (double; the traversal gets its float cast)
    Real pdfFwd, pdfBwd;      // densities of the step in flight (solid angle or discrete)
};
struct PredRec { R3 p, ng; };                  // 12 words
struct PtExtra { R3 Li, pending, refN; Real eta, bsPdf; };   // 22 words (aliases the emitter-side vertex)
struct LaneCtl { uint32_t mut; int phase; int tx; uint32_t large; };   // 4 words
struct ChainCore {            // 36 words
    Real Lx, a1, cumW;
    Real yL, zL;
    float2 posx, ypos, zpos, spos;              // spos: pixel of the path in flight (sensor vertex 2)
    float3 valx, yval, zval;
    int acc1, yn, yt, zn, zt;
    uint32_t maxIdx1;
    int pad[2];
};

enum { W_CORE = 0, W_CTL = W_CORE + sizeof(PathCore) / 4, W_VT = W_CTL + sizeof(LaneCtl) / 4, W_VTP = W_VT + sizeof(Vtx) / 4,
       W_VS = W_VTP + sizeof(PredRec) / 4, W_VSP = W_VS + sizeof(Vtx) / 4, W_CHAIN = W_VSP + sizeof(PredRec) / 4,
       W_COUNT = W_CHAIN + sizeof(ChainCore) / 4 };
static_assert(sizeof(PathCore) % 4 == 0 && sizeof(Vtx) % 4 == 0 && sizeof(ChainCore) % 4 == 0 && sizeof(PtExtra) <= sizeof(Vtx), "lane layout");

struct LaneMem {
    uint32_t *w;              // [W_COUNT][n]
    double *mis;              // [3 * (DR_MAXK + 1)][n]: pdfImp, pdfRad, conv
    float4 *rayO, *rayD;      // [n] (o, tmin), (d, tmax): the ray in flight, float32 for the traversal
    float4 *hit;              // [n] (t, u, v, leaf-order triangle or -1)
    int *rayKind;             // [n] RAY_*
    int n;
};

template <class T> DR_D void lane_load(const LaneMem &lm, int lane, int base, T &out) {
    constexpr int N = sizeof(T) / 4;
    uint32_t tmp[N];
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = lm.w[(size_t) (base + i) * lm.n + lane];
    memcpy(&out, tmp, sizeof(T));
}
template <class T> DR_D void lane_store(const LaneMem &lm, int lane, int base, const T &in) {
    constexpr int N = sizeof(T) / 4;
    uint32_t tmp[N];
    memcpy(tmp, &in, sizeof(T));
#pragma unroll
    for (int i = 0; i < N; ++i) lm.w[(size_t) (base + i) * lm.n + lane] = tmp[i];
}
enum { MIS_IMP = 0, MIS_RAD = DR_MAXK + 1, MIS_CONV = 2 * (DR_MAXK + 1) };
DR_D void mis_put(const LaneMem &lm, int lane, int arr, int i, Real v) { lm.mis[(size_t) (arr + i) * lm.n + lane] = v; }
DR_D Real mis_get(const LaneMem &lm, int lane, int arr, int i) { return lm.mis[(size_t) (arr + i) * lm.n + lane]; }

DR_D void spos_put(const LaneMem &lm, int lane, R2 sp) {     // pixel of the path in flight (ChainCore::spos)
    lm.w[(size_t) (W_CHAIN + offsetof(ChainCore, spos) / 4) * lm.n + lane] = __float_as_uint((float) sp.x);
    lm.w[(size_t) (W_CHAIN + offsetof(ChainCore, spos) / 4 + 1) * lm.n + lane] = __float_as_uint((float) sp.y);
}
DR_D R2 spos_get(const LaneMem &lm, int lane) {
    return r2(__uint_as_float(lm.w[(size_t) (W_CHAIN + offsetof(ChainCore, spos) / 4) * lm.n + lane]),
              __uint_as_float(lm.w[(size_t) (W_CHAIN + offsetof(ChainCore, spos) / 4 + 1) * lm.n + lane]));
}

// ------------------------------------------------------------------ jobs
enum { JOB_CHAIN = 0, JOB_BOOT = 1, JOB_EVAL = 2 };
struct JobParams {
    int type;
    uint32_t mutTarget;                 // JOB_CHAIN: lanes run until they have done this many mutations
    long long nItems;                   // JOB_BOOT / JOB_EVAL: lane l evaluates items l, l + n, l + 2n, ...
    unsigned long long first;           // JOB_BOOT: bootstrap sample index of item 0
    float *lumOut;                      // JOB_BOOT: luminance of every item
    const float *us, *ue, *ud;          // JOB_EVAL: replayed primary-sample vectors [nItems][d*]
    int ds, de, dd;
    const int *depthIn;                 // JOB_EVAL: MMLT depth per item
    dr_path_result *out;                // JOB_EVAL
    dr_step_record *records;            // JOB_CHAIN (parity): one record per mutation
    int recordStride;
    uint32_t mut0;
};

struct FilmParams {
    int w, h;
    float radius, scaleFactor;
    float values[32];          // rfilter.cpp:37-55 discretised filter (MTS_FILTER_RESOLUTION = 31)
};

struct ChainParams {
    Real pLarge;
    Real b;                    // m_config.luminance
    int acceptanceMap, timidAfterLarge, fixEmitterPath, useMixture, kelemenWeights;
    Real kel_s1, kel_s2, kel_logRatio;     // un-scaled Kelemen bounds for Mira's transition ratio
};

struct ChainArrays {
    double *X;                 // [dimS + dimE + dimD][n] current primary-sample vectors (SoA)
    int *depth;                // [n] MMLT depth (or -1)
    unsigned long long *chainId;   // [n] RNG key of the chain
    unsigned long long *seedIdx;   // [n] bootstrap sample the chain starts from
    int n;
    int dimS, dimE, dimD;      // allocated coordinates per sampler (maxDepth worst case)
};

enum { ST_MUT = 0, ST_FIRST_A, ST_FIRST_B, ST_LARGE_A, ST_LARGE_B, ST_BOLD_A, ST_BOLD_B, ST_SECOND_A, ST_SECOND_B,
       ST_SECOND_LARGE_A, ST_SECOND_LARGE_B, ST_SECOND_BOLD_A, ST_SECOND_BOLD_B, ST_ACC_A, ST_ACC_B, ST_PATHS, ST_RAYS, ST_COUNT };

// ------------------------------------------------------------------ film
// Splat of one (position, RGB) pair through the tabulated reconstruction filter
// (ImageBlock::put, include/mitsuba/render/imageblock.h:149-196): one 16-byte vector atomic per
// touched pixel.
DR_D void film_put(float4 *film, const FilmParams &fp, float2 pos, float3 value) {
    if (!rgb_valid(value)) return;
    const float px = pos.x - 0.5f, py = pos.y - 0.5f;
    const int minx = max((int) ceilf(px - fp.radius), 0), miny = max((int) ceilf(py - fp.radius), 0);
    const int maxx = min((int) floorf(px + fp.radius), fp.w - 1), maxy = min((int) floorf(py + fp.radius), fp.h - 1);
    for (int y = miny; y <= maxy; ++y) {
        const float wy = fp.values[min((int) fabsf((y - py) * fp.scaleFactor), 31)];
        for (int x = minx; x <= maxx; ++x) {
            const float w = fp.values[min((int) fabsf((x - px) * fp.scaleFactor), 31)] * wy;
            if (w == 0.f) continue;
            atomicAdd(film + (size_t) y * fp.w + x, make_float4(w * value.x, w * value.y, w * value.z, 0.f));
        }
    }
}

// ------------------------------------------------------------------ per-thread context of one logic step
struct Lane {
    const DevScene &sc;
    const PathCfg &pc;
    const LaneMem &lm;
    int lane;
    Pss pss;
    PathCore core;
    int depth;                // MMLT depth of the path being evaluated
    uint32_t rays;            // rays emitted in this step
    DR_D Lane(const DevScene &sc_, const PathCfg &pc_, const LaneMem &lm_, int lane_) : sc(sc_), pc(pc_), lm(lm_), lane(lane_), rays(0) {}

    DR_D void pss_restore() {
        pss.pos[0] = core.pssPos & 0xff; pss.pos[1] = (core.pssPos >> 8) & 0xff; pss.pos[2] = (core.pssPos >> 16) & 0xff;
        pss.maxIdx[0] = core.pssMax & 0xff; pss.maxIdx[1] = (core.pssMax >> 8) & 0xff; pss.maxIdx[2] = (core.pssMax >> 16) & 0xff;
        pss.cacheKey = -1;
    }
    DR_D void pss_save() {
        core.pssPos = (uint32_t) pss.pos[0] | ((uint32_t) pss.pos[1] << 8) | ((uint32_t) pss.pos[2] << 16);
        core.pssMax = (uint32_t) pss.maxIdx[0] | ((uint32_t) pss.maxIdx[1] << 8) | ((uint32_t) pss.maxIdx[2] << 16);
    }
    // queue the next ray of this lane; mint == epsilon gets the adaptive scaling of skdtree.cpp:126-129
    DR_D void emit(int kind, R3 o, R3 d, Real tmin, Real tmax) {
        const float3 of = to_f3(o);
        float mint = (float) tmin;
        if (mint == sc.epsilon) mint *= fmaxf(fmaxf(fmaxf(fabsf(of.x), fabsf(of.y)), fabsf(of.z)), sc.epsilon);
        lm.rayO[lane] = make_float4(of.x, of.y, of.z, mint);
        lm.rayD[lane] = make_float4((float) d.x, (float) d.y, (float) d.z, (float) tmax);
        lm.rayKind[lane] = kind;
        core.d = d;
        ++core.nrays;
        ++rays;
    }
};

DR_D void result_clear(PathResult &r) { r.lum = 0.; r.n = 0; r.val = r3(0.); r.pos = r2(0., 0.); r.mis = 0.; }

// ------------------------------------------------------------------ MMLT (pathsampler.cpp:84-320), resumable
// returns true when a ray was emitted (the lane waits for k_trace), false when the path is complete
DR_D bool mmlt_advance(Lane &L, const Hit &hit, PathResult &out) {
    const DevScene &sc = L.sc;
    const PathCfg &pc = L.pc;
    PathCore &core = L.core;
    const LaneMem &lm = L.lm;
    const int lane = L.lane, depth = L.depth;
    const int k = depth + 2;                                 // s + t + 1
    for (;;) {
        switch (core.pstate) {
        case PS_START: {
            result_clear(out);
            int s, t, nStrats;
            const Real decision = L.pss.next1D(SMP_DIRECT);
            if (pc.lightImage) { nStrats = depth + 1; s = min((int) (nStrats * decision), nStrats - 1); t = nStrats - s; }
            else { nStrats = depth; s = min((int) (nStrats * decision), nStrats - 1); t = 1 + (nStrats - s); }
            core.s = s; core.t = t; out.s = s; out.t = t;
            if (depth == 1) return false;
            core.connectable = 0; core.flags = 0; core.weight = r3(1.);
            (void) L.pss.next2D(SMP_SENSOR);                 // sampleSensorPosition consumes 2 (vertex.cpp:79)
            mis_put(lm, lane, MIS_RAD, k, 1.0);
            mis_put(lm, lane, MIS_RAD, k - 1, 1.0);          // supernode pdf[ERadiance] (perspective.cpp:305)
            Vtx vt;
            vt.p = cam_pos(sc.cam); vt.ng = vt.ns = cam_dir(sc.cam); vt.ss = r3(0.); vt.type = V_SENSOR_SAMPLE; vt.degenerate = 0; vt.mat = -1; vt.emitter = -1;
            core.connectable |= 1u << (k - 1);               // sensor sample: never discrete, not degenerate
            lane_store(lm, lane, W_VT, vt);
            core.j = 1;
            if (t >= 2) {                                    // vertex.cpp:126-151, perspective.cpp:318-345
                const R2 u = L.pss.next2D(SMP_SENSOR);
                const R3 dl = cam_sample_to_dir(sc.cam, u.x, u.y);
                core.pdfFwd = sc.cam.normalization / (dl.z * dl.z * dl.z);
                core.pdfBwd = 1.0;
                core.pstate = PS_SENSOR_HIT;
                L.emit(RAY_CLOSEST, vt.p, cam_xform_dir(sc.cam, dl), sc.epsilon, INFINITY);
                return true;
            }
            core.pstate = PS_EMITTER_START;
            break;
        }
        case PS_SENSOR_HIT: {                                // the ray from sensor vertex j arrived: create vertex j + 1
            if (hit.tri < 0) { result_clear(out); return false; }
            Vtx vt; PredRec vp;
            lane_load(lm, lane, W_VT, vt);
            int j = core.j;
            if (j >= 2) lane_load(lm, lane, W_VTP, vp);
            const int g = k - j;
            const R3 d = core.d;
            Vtx nv; Real tHit;
            fill_vertex(sc, hit, vt.p, d, nv, tHit);
            if (tHit == 0.) { result_clear(out); return false; }
            const Mat nm = load_material(sc, nv.mat);
            nv.degenerate = !(mat_has_smooth(nm.type) || nv.emitter >= 0);
            // solid angle -> area (vertex.cpp:334-347); delta interactions keep their discrete pdfs
            const Real cosNext = absdot(d, nv.ng);
            Real pdfFwd = core.pdfFwd, pdfBwd = core.pdfBwd;
            if (!(core.flags & F_DELTA)) {
                pdfFwd = pdfFwd / (tHit * tHit) * cosNext;
                if (j >= 2) {
                    R3 pd = vt.p - vp.p;
                    const Real plen = length(pd);
                    pd = pd / plen;
                    pdfBwd = pdfBwd / (plen * plen) * absdot(pd, vp.ng);
                }
            }
            mis_put(lm, lane, MIS_RAD, g - 1, pdfFwd);       // density of vertex j + 1
            mis_put(lm, lane, MIS_IMP, g + 1, pdfBwd);       // density of vertex j - 1
            mis_put(lm, lane, MIS_CONV, g - 1, tHit * tHit / fabs(absdot(d, vt.ng) * cosNext));   // edge (g-1, g)
            if (j == 1) {                                    // pixel of the path (pathsampler.cpp:309-312)
                R2 sp = r2(0., 0.);
                cam_sample_position(sc.cam, nv.p - cam_pos(sc.cam), sp);
                spos_put(lm, lane, sp);
            }
            vp.p = vt.p; vp.ng = vt.ng;
            vt = nv;
            ++j;
            core.j = j;
            lane_store(lm, lane, W_VT, vt);
            lane_store(lm, lane, W_VTP, vp);
            if (j < core.t) {                                // BSDF sampling step at vertex j (vertex.cpp:153-271)
                const Mat m = nm;
                WalkStep ws;
                if (!surface_sample_next(sc, vt, m, normalize(vp.p - vt.p), MODE_RADIANCE, L.pss.next2D(SMP_SENSOR), ws)) { result_clear(out); return false; }
                if (!ws.delta && !vt.degenerate) { core.connectable |= 1u << (k - j); core.flags |= F_ANYCONN; }
                core.flags = ws.delta ? (core.flags | F_DELTA) : (core.flags & ~F_DELTA);
                core.weight *= ws.weightFwd;
                core.pdfFwd = ws.pdfFwd; core.pdfBwd = ws.pdfBwd;
                L.emit(RAY_CLOSEST, vt.p, ws.wo, sc.epsilon, INFINITY);
                return true;
            }
            // last vertex: its measure stays invalid => connectable iff not degenerate
            if (!vt.degenerate) { core.connectable |= 1u << (k - core.t); core.flags |= F_ANYCONN; }
            core.pstate = PS_EMITTER_START;
            break;
        }
        case PS_EMITTER_START: {                             // emitter subpath: 0 supernode, 1 emitter sample, 2.. surfaces
            core.flags &= ~F_DELTA;
            mis_put(lm, lane, MIS_IMP, 0, 1.0);
            core.connectable |= 1u;                          // area lights: supernode not degenerate, never discrete
            if (core.s >= 1) {
                EmitterPoint ep;
                const R2 u0 = L.pss.next2D(SMP_EMITTER);
                sample_emitter_point(sc, u0.x, u0.y, ep);
                const DevEmitter &em = sc.emitters[ep.emitter];
                core.weight *= emitter_radiance(sc, ep.emitter) * (R_PI * em.area / ep.emPdf);   // m_power / emPdf
                mis_put(lm, lane, MIS_IMP, 1, ep.pdfArea);
                Vtx vs;
                vs.p = ep.p; vs.ng = vs.ns = ep.n; vs.ss = r3(0.); vs.type = V_EMITTER_SAMPLE; vs.degenerate = 0; vs.emitter = ep.emitter; vs.mat = -1;
                core.connectable |= 1u << 1;
                lane_store(lm, lane, W_VS, vs);
                core.j = 1;
                if (core.s >= 2) {                           // vertex.cpp:99-124, area.cpp:130-138
                    const R2 u = L.pss.next2D(SMP_EMITTER);
                    const R3 local = square_to_cosine_hemisphere(u.x, u.y);
                    R3 fs, ft;
                    coordinate_system(vs.ns, fs, ft);
                    core.pdfFwd = R_INV_PI * local.z; core.pdfBwd = 1.0;
                    core.pstate = PS_EMITTER_HIT;
                    L.emit(RAY_CLOSEST, vs.p, fs * local.x + ft * local.y + vs.ns * local.z, sc.epsilon, INFINITY);
                    return true;
                }
            }
            core.pstate = PS_CONNECT;
            break;
        }
        case PS_EMITTER_HIT: {
            if (hit.tri < 0) { result_clear(out); return false; }
            Vtx vs; PredRec vp;
            lane_load(lm, lane, W_VS, vs);
            int i = core.j;
            if (i >= 2) lane_load(lm, lane, W_VSP, vp);
            const R3 d = core.d;
            Vtx nv; Real tHit;
            fill_vertex(sc, hit, vs.p, d, nv, tHit);
            if (tHit == 0.) { result_clear(out); return false; }
            const Mat nm = load_material(sc, nv.mat);
            nv.degenerate = !(mat_has_smooth(nm.type) || nv.emitter >= 0);
            const Real cosNext = absdot(d, nv.ng);
            Real pdfFwd = core.pdfFwd, pdfBwd = core.pdfBwd;
            if (!(core.flags & F_DELTA)) {
                pdfFwd = pdfFwd / (tHit * tHit) * cosNext;
                if (i >= 2) {
                    R3 pd = vs.p - vp.p;
                    const Real plen = length(pd);
                    pd = pd / plen;
                    pdfBwd = pdfBwd / (plen * plen) * absdot(pd, vp.ng);
                }
            }
            mis_put(lm, lane, MIS_IMP, i + 1, pdfFwd);
            mis_put(lm, lane, MIS_RAD, i - 1, pdfBwd);
            mis_put(lm, lane, MIS_CONV, i, tHit * tHit / fabs(absdot(d, vs.ng) * cosNext));   // edge (i, i+1)
            vp.p = vs.p; vp.ng = vs.ng;
            vs = nv;
            ++i;
            core.j = i;
            lane_store(lm, lane, W_VS, vs);
            lane_store(lm, lane, W_VSP, vp);
            if (i < core.s) {
                WalkStep ws;
                if (!surface_sample_next(sc, vs, nm, normalize(vp.p - vs.p), MODE_IMPORTANCE, L.pss.next2D(SMP_EMITTER), ws)) { result_clear(out); return false; }
                if (!ws.delta && !vs.degenerate) { core.connectable |= 1u << i; core.flags |= F_ANYCONN; }
                core.flags = ws.delta ? (core.flags | F_DELTA) : (core.flags & ~F_DELTA);
                core.weight *= ws.weightFwd;
                core.pdfFwd = ws.pdfFwd; core.pdfBwd = ws.pdfBwd;
                L.emit(RAY_CLOSEST, vs.p, ws.wo, sc.epsilon, INFINITY);
                return true;
            }
            if (!vs.degenerate) { core.connectable |= 1u << core.s; core.flags |= F_ANYCONN; }
            core.pstate = PS_CONNECT;
            break;
        }
        case PS_CONNECT: {
            result_clear(out);
            if (!(core.flags & F_ANYCONN)) return false;     // pathsampler.cpp:161-174
            const int s = core.s, t = core.t;
            Vtx vt; PredRec vtp;
            lane_load(lm, lane, W_VT, vt);
            if (t >= 2) lane_load(lm, lane, W_VTP, vtp);
            if (s == 0) {                                    // pure sensor path: vt must be on an emitter (:213-224)
                if (vt.type != V_SURFACE || vt.emitter < 0) return false;
                const R3 n = vt.ns;                          // cast(): pRec.n = its.shFrame.n (records.inl:154-155)
                R3 wo = vtp.p - vt.p;
                const Real dist = length(wo);
                wo = wo / dist;
                const Real dp = dot(wo, n);
                if (!(dp > 0.)) return false;                // evalDirection (area.cpp:140-148) / |n.wo| = 1/pi
                core.weight = core.weight * emitter_radiance(sc, vt.emitter);   // radiance * pi * (1/pi)
                const DevEmitter &em = sc.emitters[vt.emitter];
                core.connectable |= 1u << 1;                 // emitter sample: area measure, not degenerate
                mis_put(lm, lane, MIS_IMP, 1, em.invArea * em.pdfDiscrete);                       // vs->evalPdf: pdfEmitterPosition
                mis_put(lm, lane, MIS_IMP, 2, R_INV_PI * dp / (dist * dist) * absdot(wo, vtp.ng));   // vt->evalPdf(vs, vtPred, EImportance)
                // connection edge of a supernode: length 0, generalized geometric term = 1 (edge.cpp:229-234, 561-571)
                if (pc.excludeDirect && depth <= 2) return false;
                core.pstate = PS_FINISH;
                break;
            }
            Vtx vs; PredRec vsp;
            lane_load(lm, lane, W_VS, vs);
            if (s >= 2) lane_load(lm, lane, W_VSP, vsp);
            if (vs.degenerate || vt.degenerate) return false;  // :253-257
            R3 d = vs.p - vt.p;                              // from vt towards vs
            const Real len = length(d);
            if (len == 0.) return false;
            d = d / len;
            R3 fs, ft;
            Mat ms, mt;
            if (s == 1) {                                    // vs->eval(vsPred, vt, EImportance)
                const Real dp = dot(-d, vs.ns);
                fs = r3(dp > 0. ? R_INV_PI : 0.);
            } else {
                ms = load_material(sc, vs.mat);
                fs = surface_eval(sc, vs, ms, normalize(vsp.p - vs.p), -d, MODE_IMPORTANCE);
            }
            if (t == 1) {
                const Real imp = cam_importance(sc.cam, cam_inv_dir(sc.cam, d));
                const Real dp = absdot(vt.ns, d);
                ft = r3(dp != 0. ? imp / dp : imp);
            } else {
                mt = load_material(sc, vt.mat);
                ft = surface_eval(sc, vt, mt, normalize(vtp.p - vt.p), d, MODE_RADIANCE);
            }
            R3 value = core.weight * fs * ft;
            if (is_zero(value)) return false;
            // generalized geometric term (edge.cpp:245-267), applied before the visibility test: an occluded
            // connection is dropped whatever its value
            value *= absdot(vs.ns, d) * absdot(vt.ns, d) / (len * len);
            core.weight = value;
            // the four densities next to the connection (path.cpp:835-859)
            core.connectable |= (1u << s) | (1u << (s + 1));   // measure forced to EArea (:263-265)
            if (s == 1) {
                const Real dp = dot(-d, vs.ns);
                mis_put(lm, lane, MIS_IMP, s + 1, R_INV_PI * fmax(dp, 0.) / (len * len) * absdot(d, vt.ng));
                mis_put(lm, lane, MIS_RAD, s - 1, 1.0);
            } else {
                mis_put(lm, lane, MIS_IMP, s + 1, surface_pdf_area(vs, ms, vsp.p, vt.p, vt.ng));
                mis_put(lm, lane, MIS_RAD, s - 1, surface_pdf_area(vs, ms, vt.p, vsp.p, vsp.ng));
            }
            if (t == 1) {
                mis_put(lm, lane, MIS_RAD, s, cam_importance(sc.cam, cam_inv_dir(sc.cam, d)) / (len * len) * absdot(d, vs.ng));
                mis_put(lm, lane, MIS_IMP, s + 2, 1.0);
                R2 sp = r2(0., 0.);
                if (!cam_sample_position(sc.cam, vs.p - vt.p, sp)) core.flags |= F_SPOS_FAIL;   // :298-303
                spos_put(lm, lane, sp);
            } else {
                mis_put(lm, lane, MIS_RAD, s, surface_pdf_area(vt, mt, vtp.p, vs.p, vs.ng));
                mis_put(lm, lane, MIS_IMP, s + 2, surface_pdf_area(vt, mt, vs.p, vtp.p, vtp.ng));
            }
            // pathConnectAndCollapse (edge.cpp:572-606): vt and vs are always "on surface" here
            core.pstate = PS_CONNECT_SHADOW;
            L.emit(RAY_SHADOW, vt.p, d, sc.epsilon, len * (1. - sc.shadowEpsilon));
            return true;
        }
        case PS_CONNECT_SHADOW: {
            result_clear(out);
            if (hit.tri >= 0) return false;                  // occluded
            if (pc.excludeDirect && depth <= 2) return false;
            if (core.flags & F_SPOS_FAIL) return false;
            core.pstate = PS_FINISH;
            break;
        }
        case PS_FINISH: {
            const int s = core.s, t = core.t;
            MisArrays A;
            A.connectable = core.connectable;
            for (int i = 0; i <= k; ++i) {
                A.pdfImp[i] = mis_get(lm, lane, MIS_IMP, i); A.pdfRad[i] = mis_get(lm, lane, MIS_RAD, i); A.conv[i] = mis_get(lm, lane, MIS_CONV, i);
            }
            const Real mis = mis_weight(A, s, t, pc.lightImage != 0);
            const int nStrats = pc.lightImage ? depth + 1 : depth;
            const R3 value = core.weight * (mis * (Real) nStrats);
            out.s = s; out.t = t; out.mis = mis;
            out.n = 1; out.pos = spos_get(lm, lane); out.val = value; out.lum = luminance(value);
            return false;
        }
        default:
            result_clear(out);
            return false;
        }
    }
}

// ------------------------------------------------------------------ unidirectional path tracer, resumable
// PathSampler EUnidirectional (pathsampler.cpp:529-567) + MIPathTracer::Li (integrators/path/path.cpp:123-312)
// with strictNormals=false, hideEmitters=false, minDepth=0, directTracing=false.  core.weight = throughput,
// core.j = depth counter; the emitter-side vertex slot holds PtExtra.
DR_D bool pt_advance(Lane &L, const Hit &hit, PathResult &out) {
    const DevScene &sc = L.sc;
    const PathCfg &pc = L.pc;
    PathCore &core = L.core;
    const LaneMem &lm = L.lm;
    const int lane = L.lane;
    PtExtra px;
    Vtx v;
    if (core.pstate != PS_START) lane_load(lm, lane, W_VS, px);
    for (;;) {
        switch (core.pstate) {
        case PS_START: {
            result_clear(out);
            out.s = out.t = -1;
            const R2 u0 = L.pss.next2D(SMP_SENSOR);
            const R2 samplePos = r2(u0.x * sc.cam.resX, u0.y * sc.cam.resY);
            spos_put(lm, lane, samplePos);
            const R3 dl = cam_sample_to_dir(sc.cam, samplePos.x / sc.cam.resX, samplePos.y / sc.cam.resY);
            const Real invZ = 1.0 / dl.z;
            px.Li = r3(0.); px.pending = r3(0.); px.refN = r3(0.); px.eta = 1.0; px.bsPdf = 0.;
            core.weight = r3(1.);
            core.flags = F_PT_FIRST | (pc.excludeDirect ? 0u : (F_PT_EMITTED | F_PT_DIRECT));
            core.j = 1;
            core.pstate = PS_PT_HIT;
            lane_store(lm, lane, W_VS, px);
            L.emit(RAY_CLOSEST, cam_pos(sc.cam), cam_xform_dir(sc.cam, dl), sc.cam.nearClip * invZ, sc.cam.farClip * invZ);
            return true;
        }
        case PS_PT_HIT: {                                    // the camera ray or a BSDF-sampled ray arrived
            if (hit.tri < 0) { core.pstate = PS_FINISH; break; }
            Vtx prev;
            lane_load(lm, lane, W_VT, prev);                 // origin of the ray (undefined for the camera ray, unused then)
            const R3 o = (core.flags & F_PT_FIRST) ? cam_pos(sc.cam) : prev.p;
            const R3 d = core.d;
            Real tHit;
            fill_vertex(sc, hit, o, d, v, tHit);
            if (!(core.flags & F_PT_FIRST)) {
                // emitter hit by the BSDF-sampled ray: MIS against direct sampling (path.cpp:242-290)
                if (v.emitter >= 0) {
                    const R3 value = dot(v.ns, -d) > 0. ? emitter_radiance(sc, v.emitter) : r3(0.);
                    Real lumPdf = 0.;
                    // pdfEmitterDirect (scene.cpp:1057-1060, area.cpp:172-180, shape.cpp:116-126) with dRec.setQuery(ray, its)
                    if (!(core.flags & F_DELTA) && dot(d, px.refN) >= 0. && dot(d, v.ns) < 0.) {
                        const DevEmitter &em = sc.emitters[v.emitter];
                        lumPdf = em.invArea * (tHit * tHit) / absdot(d, v.ns) * em.pdfDiscrete;
                    }
                    if ((core.flags & F_PT_DIRECT) && (core.flags & F_PT_NONSPEC))
                        px.Li += core.weight * value * ((px.bsPdf * px.bsPdf) / (px.bsPdf * px.bsPdf + lumPdf * lumPdf));
                }
                core.flags = (core.flags & ~F_PT_EMITTED) | F_PT_DIRECT;   // rRec.type = ERadianceNoEmission
                if (core.j++ >= pc.rrDepth) {                // Russian roulette (path.cpp:297-306)
                    const Real q = fmin(max3(core.weight) * px.eta * px.eta, 0.95);
                    if (L.pss.next1D(SMP_SENSOR) >= q) { core.pstate = PS_FINISH; break; }
                    core.weight = core.weight / q;
                }
            }
            core.flags &= ~F_PT_FIRST;
            core.pstate = PS_PT_SHADE;
            break;
        }
        case PS_PT_SHADE: {                                  // top of the loop body for the vertex v (in registers)
            const R3 d = core.d;
            const Mat m = load_material(sc, v.mat);
            if (v.emitter >= 0 && (core.flags & F_PT_EMITTED) && (core.flags & F_PT_NONSPEC) && dot(v.ns, -d) > 0.)
                px.Li += core.weight * emitter_radiance(sc, v.emitter);
            if (core.j >= pc.maxDepth && pc.maxDepth > 0) { core.pstate = PS_FINISH; break; }
            const R3 wi = to_local(v, -d);
            px.refN = mat_transmissive_or_backside(m) ? r3(0.) : v.ns;    // records.inl:160-164
            lane_store(lm, lane, W_VT, v);
            // wi of the shading point is needed again after the shadow ray: keep -d's source in VTP.p
            PredRec keep; keep.p = d; keep.ng = r3(0.);
            lane_store(lm, lane, W_VTP, keep);
            // ---- direct illumination (scene.cpp:879-904, area.cpp:156-170, shape.cpp:102-114)
            if ((core.flags & F_PT_DIRECT) && mat_has_smooth(m.type)) {
                const R2 u = L.pss.next2D(SMP_SENSOR);
                EmitterPoint ep;
                sample_emitter_point(sc, u.x, u.y, ep);
                R3 dd = ep.p - v.p;
                const Real distSq = dot(dd, dd), dist = sqrt(distSq);
                dd = dd / dist;
                const Real dp = absdot(dd, ep.n);
                Real pdf = sc.emitters[ep.emitter].invArea * (dp != 0. ? distSq / dp : 0.);
                if (dot(dd, px.refN) >= 0. && dot(dd, ep.n) < 0. && pdf != 0.) {
                    const R3 value = emitter_radiance(sc, ep.emitter) / pdf / ep.emPdf;
                    pdf *= ep.emPdf;
                    const R3 wo = to_local(v, dd);
                    const R3 bsdfVal = bsdf_eval(m, wi, wo, MODE_RADIANCE, MEAS_SOLID_ANGLE);
                    px.pending = r3(0.);
                    if (!is_zero(bsdfVal)) {
                        const Real bp = bsdf_pdf(m, wi, wo, MEAS_SOLID_ANGLE);
                        px.pending = core.weight * value * bsdfVal * ((pdf * pdf) / (pdf * pdf + bp * bp));
                    }
                    core.pstate = PS_PT_NEE;
                    lane_store(lm, lane, W_VS, px);
                    L.emit(RAY_SHADOW, v.p, dd, sc.epsilon, dist * (1. - sc.shadowEpsilon));
                    return true;
                }
            }
            core.pstate = PS_PT_BSDF;
            break;
        }
        case PS_PT_NEE: {
            if (hit.tri < 0) px.Li += px.pending;
            lane_load(lm, lane, W_VT, v);
            PredRec keep;
            lane_load(lm, lane, W_VTP, keep);
            core.d = keep.p;                                 // incoming direction at v
            core.pstate = PS_PT_BSDF;
            break;
        }
        case PS_PT_BSDF: {                                   // BSDF sampling (path.cpp:222-240)
            const R3 d = core.d;
            const Mat m = load_material(sc, v.mat);
            const R3 wi = to_local(v, -d);
            const R2 ub = L.pss.next2D(SMP_SENSOR);
            BsdfSample bs;
            bsdf_sample(m, wi, MODE_RADIANCE, ub.x, ub.y, sc.epsilon, bs);
            if (is_zero(bs.weight)) { core.pstate = PS_FINISH; break; }
            if (!(bs.sampledType & BT_DELTA)) core.flags |= F_PT_NONSPEC;
            core.flags = (bs.sampledType & BT_DELTA) ? (core.flags | F_DELTA) : (core.flags & ~F_DELTA);
            // throughput and eta are only read again if the ray hits something (path.cpp:268-274)
            core.weight *= bs.weight;
            px.eta *= bs.eta;
            px.bsPdf = bs.pdf;
            core.pstate = PS_PT_HIT;
            lane_store(lm, lane, W_VS, px);
            L.emit(RAY_CLOSEST, v.p, to_world(v, bs.wo), sc.epsilon, INFINITY);
            return true;
        }
        case PS_FINISH: {
            out.s = out.t = -1; out.mis = 0.;
            out.n = 1; out.pos = spos_get(lm, lane); out.val = px.Li; out.lum = luminance(px.Li);
            return false;
        }
        default:
            result_clear(out);
            return false;
        }
    }
}

DR_D bool path_advance(Lane &L, const Hit &hit, PathResult &out) {
    return L.pc.technique == DR_TECH_MMLT ? mmlt_advance(L, hit, out) : pt_advance(L, hit, out);
}

// findMaxDimensions (pssmlt_utils.h:27-77): MMLT vectors depend on the chain's depth
DR_D void chain_dims(const PathCfg &pc, int depth, int dimS, int dimE, int dimD, int dims[3]) {
    if (pc.technique == DR_TECH_MMLT) {
        int m = (depth + 2) * 3; if (m & 1) m++;
        dims[0] = m; dims[1] = m; dims[2] = 1;
    } else { dims[0] = dimS; dims[1] = dimE; dims[2] = dimD; }
}

DR_D bool invalid_strict(Real x) { return isnan(x) || isinf(x) || x <= 0.; }   // drmlt_proc.cpp:428
DR_D bool invalid_loose(Real x) { return isnan(x) || isinf(x) || x < 0.; }     // drmlt_proc.cpp:181
DR_D Real metropolis_clamp(Real x) { return x < 1.0 ? x : 1.0; }               // std::min(1, x): NaN -> 1

// MiraDRMLTSampler::getTransitionRatio over the three samplers (drmlt_sampler.cpp:400-414).
// dimStage holds the largest INDEX touched, so the last used coordinate is skipped.
DR_D Real mira_transition_ratio(const Pss &pss, const ChainParams &cp, const int maxIdx1[3], const int maxIdx2[3]) {
    Real ratio = 1.0;
    for (int s = 0; s < 3; ++s) {
        if (pss.identity1(s)) continue;
        Real num = 0., den = 0.;
        const int dimStage = max(maxIdx1[s], maxIdx2[s]);
        for (int i = 0; i < dimStage; i += 2) {
            const R2 y = pss.prop1(s, i >> 1), z = pss.prop2(s, i >> 1);
            num += kelemen_logpdf(z.x - y.x, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
            den += kelemen_logpdf(pss.xat(s, i) - y.x, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
            if (i + 1 < dimStage) {
                num += kelemen_logpdf(z.y - y.y, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
                den += kelemen_logpdf(pss.xat(s, i + 1) - y.y, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
            }
        }
        ratio *= exp(num - den);
    }
    return ratio;
}

// write the accepted proposal back as the new current state (DRMLTSampler::accept, drmlt_sampler.cpp:189-199)
DR_D void commit_state(const Pss &pss, double *Xbase, size_t n, const int offs[3], int slot, bool first, bool drmlt) {
    for (int s = 0; s < 3; ++s) {
        if (!pss.largeStep && (first ? pss.identity1(s) : pss.identity2(s))) continue;
        double *xs = Xbase + (size_t) offs[s] * n + slot;
        for (int k = 0; k < pss.dim[s]; k += 2) {
            R2 v = first ? pss.prop1(s, k >> 1) : pss.prop2(s, k >> 1);
            if (drmlt) { v.x = wrap_reflect(v.x); v.y = wrap_reflect(v.y); }
            xs[(size_t) k * n] = v.x;
            if (k + 1 < pss.dim[s]) xs[(size_t) (k + 1) * n] = v.y;
        }
    }
}

DR_D float3 normalized_value(const PathResult &r) {          // SplatList::normalize (pathsampler.cpp:1001-1028)
    const Real inv = r.lum > 0. ? 1.0 / r.lum : 1.0;
    return to_f3(r.val * inv);
}

// ------------------------------------------------------------------ the logic kernel
#define LOGIC_MAX_PATHS 4      // paths one logic step may complete without emitting a ray (dead-on-arrival paths)

__global__ void __launch_bounds__(128)
k_logic(const __grid_constant__ DevScene sc, const __grid_constant__ PathCfg pc, const __grid_constant__ PssParams pp,
        const __grid_constant__ ChainParams cp, const __grid_constant__ FilmParams fp, const __grid_constant__ ChainArrays ca,
        const __grid_constant__ LaneMem lm, const __grid_constant__ JobParams job, float4 *film, unsigned long long *counters,
        unsigned int *activeLanes) {
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t st[ST_COUNT];
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) st[i] = 0;
    bool active = false;
    if (lane < lm.n) {
        Lane L(sc, pc, lm, lane);
        lane_load(lm, lane, W_CORE, L.core);
        LaneCtl ctl;
        lane_load(lm, lane, W_CTL, ctl);
        const bool drmlt = pp.integrator == DR_INTEGRATOR_DRMLT;
        const int offs[3] = { 0, ca.dimS, ca.dimS + ca.dimE };
        Hit hit; hit.tri = -1; hit.t = 0.f; hit.u = hit.v = 0.f;
        if (lm.rayKind[lane] != RAY_NONE) {
            const float4 h = lm.hit[lane];
            hit.t = h.x; hit.u = h.y; hit.v = h.z; hit.tri = __float_as_int(h.w);
            lm.rayKind[lane] = RAY_NONE;
        }
        // a chain lane that had finished its budget resumes when the target was raised
        if (L.core.pstate == PS_IDLE && job.type == JOB_CHAIN && ctl.mut < job.mutTarget) { L.core.pstate = PS_START; ctl.phase = PH_STAGE1; ctl.large = 2u; }
        int completed = 0;
        while (L.core.pstate != PS_IDLE) {
            // ---- set up the primary-sample source of the path in flight
            Pss &pss = L.pss;
            pss.pp = &pp;
            long long item = 0;
            if (job.type == JOB_CHAIN) {
                L.depth = ca.depth[lane];
                int dims[3];
                chain_dims(pc, L.depth, ca.dimS, ca.dimE, ca.dimD, dims);
                pss.f32 = false; pss.stride = (size_t) ca.n;
                for (int s = 0; s < 3; ++s) { pss.xs[s] = ca.X + (size_t) offs[s] * ca.n + lane; pss.dim[s] = dims[s]; }
                pss.chain = ca.chainId[lane];
                pss.mut = ctl.mut;
                if (ctl.phase == PH_STAGE1 && ctl.large == 2u) {   // new mutation: draw the large-step coin (drmlt_proc.cpp:533)
                    ctl.large = (Real) keyed_uniform(pp.seed, S_COIN, pss.chain, ctl.mut, 0u) < cp.pLarge ? 1u : 0u;
                }
                pss.largeStep = ctl.phase != PH_INIT && ctl.large == 1u;
                pss.lightTracing = ctl.phase == PH_STAGE2 && cp.fixEmitterPath && ctl.tx == 1;   // nextStage(current->t == 1)
                pss.mode = ctl.phase == PH_INIT ? PSS_ARRAY : (ctl.phase == PH_STAGE1 ? PSS_STAGE1 : (ctl.phase == PH_STAGE2 ? PSS_STAGE2 : PSS_REVERSE));
            } else {
                item = (long long) lane + (long long) ctl.mut * lm.n;
                pss.largeStep = false; pss.lightTracing = false; pss.mut = 0;
                if (job.type == JOB_BOOT) {
                    const unsigned long long index = job.first + (unsigned long long) item;
                    L.depth = pc.technique == DR_TECH_MMLT ? (int) (index % (unsigned long long) pc.maxDepth) + 1 : -1;
                    pss.f32 = false; pss.stride = 0; pss.xs[0] = pss.xs[1] = pss.xs[2] = nullptr;
                    pss.dim[0] = pss.dim[1] = pss.dim[2] = 1 << 20;
                    pss.chain = index; pss.mode = PSS_BOOT;
                } else {
                    L.depth = job.depthIn ? job.depthIn[item] : -1;
                    pss.f32 = true; pss.stride = 1;
                    pss.xs[0] = job.us + item * job.ds; pss.xs[1] = job.ue + item * job.de; pss.xs[2] = job.ud + item * job.dd;
                    pss.dim[0] = job.ds; pss.dim[1] = job.de; pss.dim[2] = job.dd;
                    pss.chain = 0; pss.mode = PSS_ARRAY;
                }
            }
            if (L.core.pstate == PS_START) { L.core.pssPos = 0; L.core.pssMax = 0; L.core.nrays = 0; }
            L.pss_restore();
            PathResult r;
            if (path_advance(L, hit, r)) { L.pss_save(); active = true; break; }   // a ray is in flight
            hit.tri = -1;
            ++st[ST_PATHS];
            L.pss_save();
            pss.maxIdx[0] = L.core.pssMax & 0xff; pss.maxIdx[1] = (L.core.pssMax >> 8) & 0xff; pss.maxIdx[2] = (L.core.pssMax >> 16) & 0xff;

            // ================= the path is complete: job-level step =================
            if (job.type == JOB_BOOT) {
                job.lumOut[item] = (float) r.lum;
                ++ctl.mut;
                L.core.pstate = ((long long) lane + (long long) ctl.mut * lm.n < job.nItems) ? PS_START : PS_IDLE;
            } else if (job.type == JOB_EVAL) {
                dr_path_result o;
                memset(&o, 0, sizeof(o));
                o.luminance = (float) r.lum; o.n_splats = r.n; o.s = r.s; o.t = r.t; o.mis_weight = (float) r.mis;
                o.pos[0][0] = (float) r.pos.x; o.pos[0][1] = (float) r.pos.y;
                o.value[0][0] = (float) r.val.x; o.value[0][1] = (float) r.val.y; o.value[0][2] = (float) r.val.z;
                o.n_rays = L.core.nrays;
                job.out[item] = o;
                ++ctl.mut;
                L.core.pstate = ((long long) lane + (long long) ctl.mut * lm.n < job.nItems) ? PS_START : PS_IDLE;
            } else {
                ChainCore cc;
                lane_load(lm, lane, W_CHAIN, cc);
                bool mutationDone = false;
                Real a2 = 0.; bool acc2 = false;
                const bool largeStep = ctl.large == 1u;
                if (ctl.phase == PH_INIT) {
                    // seed replay (drmlt_proc.cpp:467-512): the bootstrap vector becomes the current state
                    cc.Lx = r.lum; cc.posx = make_float2((float) r.pos.x, (float) r.pos.y); cc.valx = normalized_value(r);
                    cc.cumW = 0.; ctl.tx = r.t;
                    ctl.phase = PH_STAGE1; ctl.large = 2u;
                    L.core.pstate = ctl.mut < job.mutTarget ? PS_START : PS_IDLE;
                } else if (!drmlt) {
                    // ---------------- PSSMLT (pssmlt_proc.cpp:175-272)
                    ++st[ST_MUT];
                    const Real yL = r.lum;
                    Real a = fmin(1.0, yL / cc.Lx);
                    if (isnan(yL) || yL < 0.) a = 0.;
                    a = isnan(a) ? 1.0 : a;                              // std::min(1, NaN) = 1
                    bool accept; Real currentWeight, proposedWeight;
                    if (a > 0.) {
                        if (cp.kelemenWeights) {
                            currentWeight = (1. - a) * cc.Lx / (cc.Lx / cp.b + cp.pLarge);
                            proposedWeight = (a + (largeStep ? 1. : 0.)) * yL / (yL / cp.b + cp.pLarge);
                        } else { currentWeight = 1. - a; proposedWeight = a; }
                        accept = (a == 1.) || ((Real) keyed_uniform(pp.seed, S_COIN, pss.chain, ctl.mut, 1u) < a);
                    } else {
                        currentWeight = cp.kelemenWeights ? cc.Lx / (cc.Lx / cp.b + cp.pLarge) : 1.;
                        proposedWeight = 0.; accept = false;
                    }
                    cc.cumW += currentWeight;
                    if (job.records) {
                        dr_step_record &rec = job.records[(size_t) lane * job.recordStride + (ctl.mut - job.mut0)];
                        rec.L_x = (float) cc.Lx; rec.L_y = (float) yL; rec.L_z = 0.f; rec.a1 = (float) a; rec.a2 = 0.f;
                        rec.large_step = largeStep; rec.accept1 = accept; rec.did_second = 0; rec.accept2 = 0;
                    }
                    ++st[ST_ACC_B];
                    if (largeStep) ++st[ST_LARGE_B]; else ++st[ST_BOLD_B];
                    const float3 yval = normalized_value(r);
                    if (accept) {
                        const float3 v = cc.valx * (float) cc.cumW;
                        if (film && !is_zero(v)) film_put(film, fp, cc.posx, v);
                        cc.cumW = proposedWeight;
                        commit_state(pss, ca.X, (size_t) ca.n, offs, lane, true, false);
                        cc.Lx = yL; cc.posx = make_float2((float) r.pos.x, (float) r.pos.y); cc.valx = yval; ctl.tx = r.t;
                        ++st[ST_ACC_A];
                        if (largeStep) ++st[ST_LARGE_A]; else ++st[ST_BOLD_A];
                    } else {
                        const float3 v = yval * (float) proposedWeight;
                        if (film && r.n && !is_zero(v)) film_put(film, fp, make_float2((float) r.pos.x, (float) r.pos.y), v);
                    }
                    mutationDone = true;
                } else if (ctl.phase == PH_STAGE1) {
                    // ---------------- DRMLT first stage (drmlt_proc.cpp:539-558; mixture :284-299)
                    ++st[ST_MUT];
                    cc.yL = r.lum; cc.ypos = make_float2((float) r.pos.x, (float) r.pos.y); cc.yval = normalized_value(r); cc.yn = r.n; cc.yt = r.t;
                    cc.maxIdx1 = L.core.pssMax;
                    cc.a1 = 0.; cc.acc1 = 0; cc.zL = 0.; cc.zn = 0;
                    bool doSecond;
                    if (cp.useMixture) {
                        if (!invalid_loose(cc.yL)) {
                            cc.a1 = metropolis_clamp(cc.yL / cc.Lx);
                            cc.acc1 = (cc.a1 >= 1.) || ((Real) keyed_uniform(pp.seed, S_COIN, pss.chain, ctl.mut, 1u) < cc.a1);
                        }
                        doSecond = !largeStep && ((Real) keyed_uniform(pp.seed, S_COIN, pss.chain, ctl.mut, 3u) < 0.5);
                    } else {
                        if (!invalid_strict(cc.yL)) {
                            cc.a1 = metropolis_clamp(cc.yL / cc.Lx);
                            cc.acc1 = (cc.a1 >= 1.) || ((Real) keyed_uniform(pp.seed, S_COIN, pss.chain, ctl.mut, 1u) < cc.a1);
                        }
                        doSecond = !cc.acc1 && (cp.timidAfterLarge || !largeStep);
                    }
                    if (doSecond) { ctl.phase = PH_STAGE2; L.core.pstate = PS_START; }
                    else mutationDone = true;
                } else if (ctl.phase == PH_STAGE2) {
                    cc.zL = r.lum; cc.zpos = make_float2((float) r.pos.x, (float) r.pos.y); cc.zval = normalized_value(r); cc.zn = r.n; cc.zt = r.t;
                    mutationDone = true;
                    if (cp.useMixture) {   // drmlt_proc.cpp:317-324: plain MH on the replaced proposal
                        cc.a1 = 0.; cc.acc1 = 0;
                        if (!invalid_loose(cc.zL)) {
                            a2 = metropolis_clamp(cc.zL / cc.Lx);
                            acc2 = (a2 >= 1.) || ((Real) keyed_uniform(pp.seed, S_COIN, pss.chain, ctl.mut, 2u) < a2);
                        }
                    } else if (!invalid_strict(cc.zL)) {
                        if (pp.type == DR_TYPE_GREEN) { ctl.phase = PH_REVERSE; L.core.pstate = PS_START; mutationDone = false; }
                        else if (pp.type == DR_TYPE_MIRA) {   // drmlt_proc.cpp:625-650
                            const Real aReverse = metropolis_clamp(cc.yL / cc.zL);
                            if (!(aReverse >= 1.)) {
                                int m1[3] = { (int) (cc.maxIdx1 & 0xff), (int) ((cc.maxIdx1 >> 8) & 0xff), (int) ((cc.maxIdx1 >> 16) & 0xff) };
                                const Real T = largeStep ? 1.0 : mira_transition_ratio(pss, cp, m1, pss.maxIdx);
                                if (!invalid_strict(T)) {
                                    a2 = metropolis_clamp((cc.zL / cc.Lx) * T * (1.0 - aReverse) / (1.0 - cc.a1));
                                    acc2 = (a2 >= 1.) || ((Real) keyed_uniform(pp.seed, S_COIN, pss.chain, ctl.mut, 2u) < a2);
                                }
                            }
                        } else {                               // orbital, drmlt_proc.cpp:655-669
                            if (cc.zL < cc.yL) { a2 = 0.; }
                            else if (cc.zL >= cc.Lx) { a2 = 1.0; acc2 = true; }
                            else {
                                a2 = (cc.zL - cc.yL) / (cc.Lx - cc.yL);
                                acc2 = (a2 >= 1.) || ((Real) keyed_uniform(pp.seed, S_COIN, pss.chain, ctl.mut, 2u) < a2);
                            }
                        }
                    }
                } else {                   // Green's reverse path (drmlt_proc.cpp:588-621)
                    const Real Lr = r.lum;
                    const Real aReverse = invalid_strict(Lr) ? 0. : metropolis_clamp(Lr / cc.zL);
                    if (aReverse != 1.) {
                        a2 = metropolis_clamp((cc.zL / cc.Lx) * (1. - aReverse) / (1. - cc.a1));
                        acc2 = (a2 >= 1.) || ((Real) keyed_uniform(pp.seed, S_COIN, pss.chain, ctl.mut, 2u) < a2);
                    }
                    mutationDone = true;
                }

                if (mutationDone && drmlt) {
                    // ---- splat with expectation weights (drmlt_proc.cpp:676-688; mixture :327-333)
                    const bool did2 = ctl.phase != PH_STAGE1;
                    const bool acc1 = cc.acc1 != 0;
                    const Real a1 = cc.a1;
                    if (job.records) {
                        dr_step_record &rec = job.records[(size_t) lane * job.recordStride + (ctl.mut - job.mut0)];
                        rec.L_x = (float) cc.Lx; rec.L_y = (float) cc.yL; rec.L_z = did2 ? (float) cc.zL : 0.f; rec.a1 = (float) a1; rec.a2 = (float) a2;
                        rec.large_step = largeStep; rec.accept1 = acc1; rec.did_second = did2; rec.accept2 = acc2;
                    }
                    if (film && !cp.acceptanceMap) {
                        Real wy, wz, wx;
                        if (cp.useMixture) { wy = did2 ? 0. : a1; wz = did2 ? a2 : 0.; wx = 1.0 - (did2 ? a2 : a1); }
                        else { wy = a1; wz = (1.0 - a1) * a2; wx = 1.0 - wy - wz; }
                        if (wx > 0.) film_put(film, fp, cc.posx, cc.valx * (float) wx);
                        if (wy > 0. && cc.yn) film_put(film, fp, cc.ypos, cc.yval * (float) wy);
                        if (wz > 0. && cc.zn) film_put(film, fp, cc.zpos, cc.zval * (float) wz);
                    }
                    // ---- accept / reject, statistics (drmlt_proc.cpp:691-769; mixture :335-378)
                    if (cp.useMixture) {
                        const bool accept = did2 ? acc2 : acc1;
                        ++st[ST_ACC_B];
                        if (!did2) { ++st[ST_FIRST_B]; if (largeStep) ++st[ST_LARGE_B]; else ++st[ST_BOLD_B]; } else ++st[ST_SECOND_B];
                        if (accept) {
                            ++st[ST_ACC_A];
                            if (!did2) { ++st[ST_FIRST_A]; if (largeStep) ++st[ST_LARGE_A]; else ++st[ST_BOLD_A]; } else ++st[ST_SECOND_A];
                            commit_state(pss, ca.X, (size_t) ca.n, offs, lane, !did2, true);
                            if (did2) { cc.Lx = cc.zL; cc.posx = cc.zpos; cc.valx = cc.zval; ctl.tx = cc.zt; }
                            else { cc.Lx = cc.yL; cc.posx = cc.ypos; cc.valx = cc.yval; ctl.tx = cc.yt; }
                        }
                    } else if (acc1 || acc2) {
                        // acceptance map: binned at the state that is being LEFT (proposed.* after the swap, :695-708)
                        if (film && cp.acceptanceMap && (acc1 ? !largeStep : true))
                            film_put(film, fp, cc.posx, acc1 ? make_float3(1.f, 0.f, 0.f) : make_float3(0.f, 1.f, 0.f));
                        commit_state(pss, ca.X, (size_t) ca.n, offs, lane, acc1, true);
                        if (acc1) { cc.Lx = cc.yL; cc.posx = cc.ypos; cc.valx = cc.yval; ctl.tx = cc.yt; }
                        else { cc.Lx = cc.zL; cc.posx = cc.zpos; cc.valx = cc.zval; ctl.tx = cc.zt; }
                        ++st[ST_ACC_B]; ++st[ST_ACC_A];
                        if (acc1) {
                            ++st[ST_FIRST_B]; ++st[ST_FIRST_A];
                            if (largeStep) { ++st[ST_LARGE_B]; ++st[ST_LARGE_A]; } else { ++st[ST_BOLD_B]; ++st[ST_BOLD_A]; }
                        } else {
                            ++st[ST_ACC_B]; ++st[ST_FIRST_B]; ++st[ST_SECOND_B]; ++st[ST_SECOND_A];
                            if (largeStep) { ++st[ST_LARGE_B]; ++st[ST_SECOND_LARGE_B]; ++st[ST_SECOND_LARGE_A]; }
                            else { ++st[ST_BOLD_B]; ++st[ST_SECOND_BOLD_B]; ++st[ST_SECOND_BOLD_A]; }
                        }
                    } else {
                        ++st[ST_ACC_B]; ++st[ST_FIRST_B];
                        if (largeStep) { ++st[ST_LARGE_B]; if (did2) { ++st[ST_SECOND_B]; ++st[ST_SECOND_LARGE_B]; ++st[ST_ACC_B]; } }
                        else { ++st[ST_BOLD_B]; if (did2) { ++st[ST_SECOND_B]; ++st[ST_SECOND_BOLD_B]; ++st[ST_ACC_B]; } }
                    }
                }
                if (mutationDone) {
                    ++ctl.mut;
                    ctl.phase = PH_STAGE1; ctl.large = 2u;
                    L.core.pstate = ctl.mut < job.mutTarget ? PS_START : PS_IDLE;
                }
                lane_store(lm, lane, W_CHAIN, cc);
            }
            if (++completed >= LOGIC_MAX_PATHS && L.core.pstate != PS_IDLE) { active = true; break; }   // resume next round
        }
        st[ST_RAYS] = L.rays;
        lane_store(lm, lane, W_CORE, L.core);
        lane_store(lm, lane, W_CTL, ctl);
    }
    // lanes with work left (a ray in flight or a path to start) keep the host loop going
    const unsigned int ballot = __ballot_sync(0xffffffffu, active);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(activeLanes, (unsigned int) __popc(ballot));
    // per-warp counters -> global (the reference's StatsCounter, statistics.h:80-110)
    unsigned int any = 0;
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) any |= st[i];
    if (__any_sync(0xffffffffu, any != 0)) {
#pragma unroll
        for (int i = 0; i < ST_COUNT; ++i) {
            unsigned int v = st[i];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counters[i], (unsigned long long) v);
        }
    }
}

// ------------------------------------------------------------------ the traversal kernel
__global__ void __launch_bounds__(256)
k_trace(const __grid_constant__ DevScene sc, const __grid_constant__ LaneMem lm) {
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lm.n) return;
    const int kind = lm.rayKind[lane];
    if (kind == RAY_NONE) return;
    const float4 a = lm.rayO[lane], b = lm.rayD[lane];
    Hit h;
    const float3 o = f3(a.x, a.y, a.z), d = f3(b.x, b.y, b.z);
    const bool found = kind == RAY_SHADOW ? traverse<true>(sc, o, d, a.w, b.w, h) : traverse<false>(sc, o, d, a.w, b.w, h);
    lm.hit[lane] = make_float4(h.t, h.u, h.v, __int_as_float(found ? h.tri : -1));
}

// ------------------------------------------------------------------ lane set-up
// JOB_CHAIN: seed replay + fillReplay (drmlt_proc.cpp:467-504): current = the seed's bootstrap vector; the lane then
// evaluates it (PH_INIT) before its first mutation.  JOB_BOOT / JOB_EVAL: lane l starts at item l.
__global__ void k_setup_lanes(const __grid_constant__ PssParams pp, const __grid_constant__ ChainArrays ca, const __grid_constant__ LaneMem lm,
                              const __grid_constant__ JobParams job) {
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lm.n) return;
    PathCore core;
    memset(&core, 0, sizeof(core));
    LaneCtl ctl;
    ctl.mut = 0; ctl.tx = -1; ctl.large = 2u; ctl.phase = PH_STAGE1;
    if (job.type == JOB_CHAIN) {
        const unsigned long long sidx = ca.seedIdx[lane];
        const int offs[3] = { 0, ca.dimS, ca.dimS + ca.dimE }, alloc[3] = { ca.dimS, ca.dimE, ca.dimD };
        for (int s = 0; s < 3; ++s)
            for (int k = 0; k < alloc[s]; ++k)
                ca.X[(size_t) (offs[s] + k) * ca.n + lane] = (double) keyed_uniform(pp.seed, S_BOOT, sidx, (uint32_t) s, (uint32_t) k);
        ctl.phase = PH_INIT;
        core.pstate = PS_START;
    } else {
        core.pstate = lane < job.nItems ? PS_START : PS_IDLE;
    }
    lane_store(lm, lane, W_CORE, core);
    lane_store(lm, lane, W_CTL, ctl);
    lm.rayKind[lane] = RAY_NONE;
}

// PSSMLT's "last splat" of the accumulated current state (pssmlt_proc.cpp:274-279); resets the weight
__global__ void k_flush_pssmlt(const __grid_constant__ LaneMem lm, const __grid_constant__ FilmParams fp, float4 *film) {
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lm.n) return;
    ChainCore cc;
    lane_load(lm, lane, W_CHAIN, cc);
    const float3 c = cc.valx * (float) cc.cumW;
    if (!is_zero(c)) film_put(film, fp, cc.posx, c);
    cc.cumW = 0.;
    lane_store(lm, lane, W_CHAIN, cc);
}
