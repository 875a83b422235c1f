// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
//
// orc_path.hpp: bidirectional path vertices/edges, random walks, MIS and the three path
// sampling techniques.  Restates
//   src/libbidir/vertex.cpp:29-350   (makeEndpoint, sampleNext)
//   src/libbidir/vertex.cpp:958-1205 (eval, evalPdf), :1384-1433 (cast), :1578-1586 (getSamplePosition)
//   src/libbidir/edge.cpp:27-84 (sampleNext), :221-271 (evalCached), :558-690 (pathConnectAndCollapse)
//   src/libbidir/path.cpp:500-535 (randomWalk), :763-1028 (miWeight)
//   src/libbidir/pathsampler.cpp:79-571 (sampleSplats: MMLT / BDPT / PT)
//   src/integrators/path/path.cpp:123-312 (MIPathTracer::Li)
// Vacuum only (no participating media, no ENull BSDFs): edge weights/pdfs are 1.
#pragma once
#include "orc_scene.hpp"
#include "orc_bsdf.hpp"

namespace orc {

struct Sampler {
    virtual ~Sampler() {}
    virtual Float next1D() = 0;
    Vec2 next2D() { Float a = next1D(); Float b = next1D(); return Vec2(a, b); }
};

// A sampler that replays an explicit primary-sample vector (replayed-u parity entry point).
struct ArraySampler : Sampler {
    const float *u; int n; int pos = 0; bool overflow = false;
    ArraySampler(const float *u_, int n_) : u(u_), n(n_) {}
    Float next1D() override {
        if (pos >= n) { overflow = true; ++pos; return 0.5; }
        return (Float) u[pos++];
    }
};

enum EVertexType { EInvalid = 0, EEmitterSupernode, ESensorSupernode, EEmitterSample, ESensorSample, ESurfaceInteraction };

struct PathVertex {
    int type = EInvalid;
    int measure = EInvalidMeasure;
    bool degenerate = false;
    RGB weight[2];
    Float pdf[2] = { 0, 0 };
    Float rrWeight = 1.0;
    // ESurfaceInteraction
    Intersection its;
    // EEmitterSample / ESensorSample (PositionSamplingRecord)
    Vec3 p, n;
    Vec2 uv;
    int emitter = -1;

    bool isEmitterSupernode() const { return type == EEmitterSupernode; }
    bool isSensorSupernode() const { return type == ESensorSupernode; }
    bool isSupernode() const { return type == EEmitterSupernode || type == ESensorSupernode; }
    bool isSensorSample() const { return type == ESensorSample; }
    bool isSurfaceInteraction() const { return type == ESurfaceInteraction; }
    // vertex.h:624-628 -- area lights and the pinhole camera are both EOnSurface
    bool isOnSurface() const { return type == ESurfaceInteraction || type == EEmitterSample || type == ESensorSample; }
    bool isDegenerate() const { return degenerate; }
    bool isConnectable() const { return !degenerate && measure != EDiscrete; }   // vertex.h:782
    Vec3 getPosition() const { return type == ESurfaceInteraction ? its.p : p; }
    Vec3 getShadingNormal() const { return type == ESurfaceInteraction ? its.sh.n : n; }
    Vec3 getGeometricNormal() const { return type == ESurfaceInteraction ? its.ng : n; }
};

struct PathEdge {
    Vec3 d;
    Float length = 0;
    RGB weight[2];
    Float pdf[2] = { 0, 0 };
};

struct PathCtx {
    const Scene *scene;
    uint64_t rays = 0;
};

// ------------------------------------------------------------------ emitter / sensor endpoint sampling
// scene.cpp:1066-1082 + area.cpp:98-102 + trimesh.cpp:429-440 + triangle.cpp:24-60
inline RGB sampleEmitterPosition(const Scene &sc, Vec2 sample, Vec3 &p, Vec3 &n, Vec2 &uv, Float &pdf, int &emitter) {
    Float emPdf;
    size_t index = sc.emitterPDF.sampleReuse(sample.x, emPdf);
    const EmitterRec &em = sc.emitters[index];
    size_t triLocal = em.areaDistr.sampleReuse(sample.y);
    int tri = em.firstTri + (int) triLocal;
    const uint32_t i0 = sc.idx[3 * tri], i1 = sc.idx[3 * tri + 1], i2 = sc.idx[3 * tri + 2];
    const Vec3 &p0 = sc.P[i0], &p1 = sc.P[i1], &p2 = sc.P[i2];
    Vec2 bary = squareToUniformTriangle(sample);
    Vec3 sideA = p1 - p0, sideB = p2 - p0;
    p = p0 + sideA * bary.x + sideB * bary.y;
    if ((sc.triFlags[tri] & DR_TRI_SMOOTH) && !sc.N.empty())
        n = normalize(sc.N[i0] * (1.0 - bary.x - bary.y) + sc.N[i1] * bary.x + sc.N[i2] * bary.y);
    else
        n = normalize(cross(sideA, sideB));
    uv = bary;
    pdf = em.invArea * emPdf;
    emitter = (int) index;
    return em.power() / emPdf;
}

inline Float pdfEmitterDiscrete(const Scene &sc, int emitter) { return sc.emitterPDF[emitter]; }
inline Float pdfEmitterPosition(const Scene &sc, int emitter) {   // scene.cpp:1084-1087, trimesh.cpp:374-376
    return sc.emitters[emitter].invArea * pdfEmitterDiscrete(sc, emitter);
}

// DirectSamplingRecord (include/mitsuba/render/common.h + records.inl:160-178)
struct DirectRec {
    Vec3 ref, refN;
    Vec3 p, n, d;
    Float dist = 0, pdf = 0;
    int measure = EInvalidMeasure;
    int emitter = -1;
};

// scene.cpp:879-904 + area.cpp:156-170 + shape.cpp:102-114
inline RGB sampleEmitterDirect(PathCtx &ctx, DirectRec &dRec, Vec2 sample, bool testVisibility) {
    const Scene &sc = *ctx.scene;
    Float emPdf;
    size_t index = sc.emitterPDF.sampleReuse(sample.x, emPdf);
    const EmitterRec &em = sc.emitters[index];
    // Shape::sampleDirect -> TriMesh::samplePosition
    size_t triLocal = em.areaDistr.sampleReuse(sample.y);
    int tri = em.firstTri + (int) triLocal;
    const uint32_t i0 = sc.idx[3 * tri], i1 = sc.idx[3 * tri + 1], i2 = sc.idx[3 * tri + 2];
    const Vec3 &p0 = sc.P[i0], &p1 = sc.P[i1], &p2 = sc.P[i2];
    Vec2 bary = squareToUniformTriangle(sample);
    Vec3 sideA = p1 - p0, sideB = p2 - p0;
    dRec.p = p0 + sideA * bary.x + sideB * bary.y;
    if ((sc.triFlags[tri] & DR_TRI_SMOOTH) && !sc.N.empty())
        dRec.n = normalize(sc.N[i0] * (1.0 - bary.x - bary.y) + sc.N[i1] * bary.x + sc.N[i2] * bary.y);
    else
        dRec.n = normalize(cross(sideA, sideB));
    dRec.pdf = em.invArea;
    dRec.d = dRec.p - dRec.ref;
    Float distSquared = lengthSquared(dRec.d);
    dRec.dist = std::sqrt(distSquared);
    dRec.d /= dRec.dist;
    Float dp = absDot(dRec.d, dRec.n);
    dRec.pdf *= dp != 0 ? (distSquared / dp) : 0.0;
    dRec.measure = ESolidAngle;
    RGB value(0.0);
    if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0 && dRec.pdf != 0) {
        value = em.radiance / dRec.pdf;
    } else {
        dRec.pdf = 0.0;
    }
    if (dRec.pdf != 0) {
        if (testVisibility) {
            Ray ray; ray.o = dRec.ref; ray.d = dRec.d; ray.mint = sc.epsilon; ray.maxt = dRec.dist * (1 - sc.shadowEpsilon);
            if (sc.rayIntersectShadow(ray, &ctx.rays)) return RGB(0.0);
        }
        dRec.emitter = (int) index;
        dRec.pdf *= emPdf;
        value /= emPdf;
        return value;
    }
    return RGB(0.0);
}

// scene.cpp:1057-1060 + area.cpp:172-180 + shape.cpp:116-126
inline Float pdfEmitterDirect(const Scene &sc, const DirectRec &dRec) {
    if (!(dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0)) return 0.0;
    Float pdfPos = sc.emitters[dRec.emitter].invArea, r;
    if (dRec.measure == ESolidAngle) r = pdfPos * (dRec.dist * dRec.dist) / absDot(dRec.d, dRec.n);
    else if (dRec.measure == EArea) r = pdfPos;
    else r = 0.0;
    return r * pdfEmitterDiscrete(sc, dRec.emitter);
}

// ------------------------------------------------------------------ PathVertex / PathEdge operations
struct Path {
    std::vector<PathVertex> v;
    std::vector<PathEdge> e;
    size_t vertexCount() const { return v.size(); }
    size_t edgeCount() const { return e.size(); }
    PathVertex *vertexOrNull(int i) { return (i < 0 || i >= (int) v.size()) ? nullptr : &v[i]; }
    PathEdge *edgeOrNull(int i) { return (i < 0 || i >= (int) e.size()) ? nullptr : &e[i]; }
    void initialize(int mode) {   // path.cpp:493-498 + vertex.cpp:29-35
        v.clear(); e.clear();
        PathVertex sv;
        sv.type = (mode == EImportance) ? EEmitterSupernode : ESensorSupernode;
        // hasDegenerateEmitters() = false (area lights), hasDegenerateSensor() = true (pinhole)
        sv.degenerate = (mode == EImportance) ? false : true;
        v.push_back(sv);
    }
};

// edge.cpp:27-84 (vacuum)
inline bool edgeSampleNext(PathCtx &ctx, PathEdge &edge, const Ray &ray, PathVertex &succ, int mode) {
    const Scene &sc = *ctx.scene;
    bool surface = sc.rayIntersect(ray, succ.its, &ctx.rays);
    if (!surface) return false;
    succ.type = ESurfaceInteraction;
    const Mat m = materialAt(sc, succ.its);
    succ.degenerate = !(bsdfHasSmooth(m) || succ.its.emitter >= 0);
    edge.length = succ.its.t;
    if (edge.length == 0) return false;
    edge.weight[ERadiance] = edge.weight[EImportance] = RGB(1.0);
    edge.pdf[ERadiance] = edge.pdf[EImportance] = 1.0;
    edge.d = ray.d;
    if (mode == ERadiance) edge.d = -edge.d;
    return true;
}

// vertex.cpp:37-350
inline bool vertexSampleNext(PathCtx &ctx, PathVertex &cur, Sampler *sampler, const PathVertex *pred,
                             const PathEdge *predEdge, PathEdge &succEdge, PathVertex &succ, int mode,
                             bool russianRoulette, RGB *throughput) {
    const Scene &sc = *ctx.scene;
    Ray ray = sc.makeRay(Vec3(), Vec3());
    succEdge = PathEdge();
    succ = PathVertex();
    cur.rrWeight = 1.0;

    switch (cur.type) {
    case EEmitterSupernode: {   // :50-72
        Float pdf; RGB result = sampleEmitterPosition(sc, sampler->next2D(), succ.p, succ.n, succ.uv, pdf, succ.emitter);
        if (result.isZero()) return false;
        cur.weight[EImportance] = result;
        cur.pdf[EImportance] = pdf;
        cur.measure = EArea;
        succ.type = EEmitterSample;
        succ.degenerate = false;
        succEdge.weight[EImportance] = RGB(1.0);
        succEdge.pdf[EImportance] = 1.0;
        return true;
    }
    case ESensorSupernode: {   // :74-97 + perspective.cpp:300-308
        (void) sampler->next2D();
        succ.p = sc.cam.pos; succ.n = sc.cam.dir;
        cur.weight[ERadiance] = RGB(1.0);
        cur.pdf[ERadiance] = 1.0;
        cur.measure = EDiscrete;
        succ.type = ESensorSample;
        succ.degenerate = false;
        succEdge.weight[ERadiance] = RGB(1.0);
        succEdge.pdf[ERadiance] = 1.0;
        return true;
    }
    case EEmitterSample: {   // :99-124 + area.cpp:130-138
        Vec3 local = squareToCosineHemisphere(sampler->next2D());
        Vec3 d = Frame(cur.n).toWorld(local);
        Float dpdf = squareToCosineHemispherePdf(local);
        RGB result(1.0);
        cur.weight[EImportance] = result;
        cur.weight[ERadiance] = result * dpdf * (1.0 / absDot(d, cur.n));
        cur.pdf[EImportance] = dpdf;
        cur.pdf[ERadiance] = 1.0;
        cur.measure = ESolidAngle;
        ray.o = cur.p; ray.d = d;
        break;
    }
    case ESensorSample: {   // :126-151 + perspective.cpp:318-345
        Vec2 sample = sampler->next2D();
        cur.uv = Vec2(sample.x * sc.cam.resX, sample.y * sc.cam.resY);
        Vec3 dl = sc.cam.sampleToDir(sample.x, sample.y);
        Vec3 d = sc.cam.xformDir(dl);
        Float dpdf = sc.cam.normalization / (dl.z * dl.z * dl.z);
        RGB result(1.0);
        cur.weight[EImportance] = result * dpdf * (1.0 / absDot(d, cur.n));
        cur.weight[ERadiance] = result;
        cur.pdf[EImportance] = 1.0;
        cur.pdf[ERadiance] = dpdf;
        cur.measure = ESolidAngle;
        ray.o = cur.p; ray.d = d;
        break;
    }
    case ESurfaceInteraction: {   // :153-271
        const Intersection &its = cur.its;
        const Mat bsdf = materialAt(sc, its);
        Vec3 wi = normalize(pred->getPosition() - its.p);
        BSDFRecord bRec(its.toLocal(wi), mode);
        Vec2 rndPoint = sampler->next2D();
        // EUsesSampler BSDFs draw one more number inside sample() (roughdielectric.cpp:555); drawing it up front is
        // equivalent: the only exit before the draw returns a zero weight, which ends the walk
        Float extra = bsdfUsesSampler(bsdf) ? sampler->next1D() : 0.5;
        cur.weight[mode] = bsdfSample(bsdf, bRec, cur.pdf[mode], rndPoint, sc.epsilon, extra);
        if (cur.weight[mode].isZero()) return false;
        cur.measure = bsdfMeasure(bRec.sampledType);
        Vec3 wo = its.toWorld(bRec.wo);
        Float wiDotGeoN = dot(its.ng, wi), woDotGeoN = dot(its.ng, wo);
        if (wiDotGeoN * Frame::cosTheta(bRec.wi) <= 0 || woDotGeoN * Frame::cosTheta(bRec.wo) <= 0) return false;
        bRec.reverse();
        cur.pdf[1 - mode] = bsdfPdf(bsdf, bRec, cur.measure);
        if (cur.pdf[1 - mode] <= RCPOVERFLOW) return false;
        if (!bsdfNonSymmetric(bsdf)) {
            cur.weight[1 - mode] = cur.weight[mode] * (cur.pdf[mode] / cur.pdf[1 - mode]);
            if (cur.measure == ESolidAngle)
                cur.weight[1 - mode] *= std::abs(Frame::cosTheta(bRec.wo) / Frame::cosTheta(bRec.wi));
        } else {
            cur.weight[1 - mode] = bsdfEval(bsdf, bRec, cur.measure) / cur.pdf[1 - mode];
        }
        bRec.reverse();
        // adjoint BSDF for shading normals (option.adjointComp = true, path.h:118-120)
        if (mode == EImportance)
            cur.weight[EImportance] *= std::abs((Frame::cosTheta(bRec.wi) * woDotGeoN) / (Frame::cosTheta(bRec.wo) * wiDotGeoN));
        else
            cur.weight[EImportance] *= std::abs((Frame::cosTheta(bRec.wo) * wiDotGeoN) / (Frame::cosTheta(bRec.wi) * woDotGeoN));
        if (throughput && mode == ERadiance && bRec.eta != 1) (*throughput) *= bRec.eta * bRec.eta;
        ray.o = its.p; ray.d = wo;
        break;
    }
    default:
        return false;
    }

    if (throughput) {   // :307-322
        (*throughput) *= cur.weight[mode];
        if (russianRoulette) {
            Float q = std::min(throughput->max(), (Float) 0.95f);
            if (sampler->next1D() > q) { cur.measure = EInvalidMeasure; return false; }
            cur.rrWeight = 1.0 / q;
            (*throughput) *= cur.rrWeight;
        }
    }
    if (!edgeSampleNext(ctx, succEdge, ray, succ, mode)) { cur.measure = EInvalidMeasure; return false; }
    if (cur.measure == ESolidAngle) {   // :334-347
        cur.measure = EArea;
        cur.pdf[mode] /= succEdge.length * succEdge.length;
        if (succ.isOnSurface()) cur.pdf[mode] *= absDot(ray.d, succ.getGeometricNormal());
        if (predEdge->length != 0.0) {
            cur.pdf[1 - mode] /= predEdge->length * predEdge->length;
            if (pred->isOnSurface()) cur.pdf[1 - mode] *= absDot(predEdge->d, pred->getGeometricNormal());
        }
    }
    return true;
}

// path.cpp:500-535
inline int randomWalk(PathCtx &ctx, Path &path, Sampler *sampler, int nSteps, int rrStart, int mode) {
    RGB throughput(1.0);
    for (int i = 0; i < nSteps || nSteps == -1; ++i) {
        size_t nv = path.v.size();
        // keep references stable: reserve
        path.v.reserve(nv + 2); path.e.reserve(path.e.size() + 2);
        PathVertex succV; PathEdge succE;
        PathVertex *cur = &path.v[nv - 1];
        const PathVertex *pred = nv < 2 ? nullptr : &path.v[nv - 2];
        const PathEdge *predEdge = path.e.empty() ? nullptr : &path.e.back();
        bool rr = rrStart != -1 && i >= rrStart;
        if (!vertexSampleNext(ctx, *cur, sampler, pred, predEdge, succE, succV, mode, rr, &throughput))
            return i;
        path.e.push_back(succE);
        path.v.push_back(succV);
    }
    return nSteps;
}

// vertex.cpp:958-1092 (adjointComp = true)
inline RGB vertexEval(const Scene &sc, const PathVertex &cur, const PathVertex *pred, const PathVertex *succ, int mode,
                      int measure = EArea) {
    RGB result(0.0);
    switch (cur.type) {
    case EEmitterSupernode:
        if (mode != EImportance || pred != nullptr || succ->type != EEmitterSample) return RGB(0.0);
        return sc.emitters[succ->emitter].radiance * PI;   // area.cpp:104-106
    case ESensorSupernode:
        if (mode != ERadiance || pred != nullptr || succ->type != ESensorSample) return RGB(0.0);
        return RGB(measure == EDiscrete ? 1.0 : 0.0);      // perspective.cpp:310-312
    case EEmitterSample: {
        Vec3 target;
        if (mode == EImportance && pred->type == EEmitterSupernode) target = succ->getPosition();
        else if (mode == ERadiance && succ->type == EEmitterSupernode) target = pred->getPosition();
        else return RGB(0.0);
        Vec3 wo = normalize(target - cur.p);
        int m = measure == EArea ? ESolidAngle : measure;
        Float dp = dot(wo, cur.n);                          // area.cpp:140-148
        if (m != ESolidAngle || dp < 0) dp = 0.0;
        result = RGB(INV_PI * dp);
        Float adp = absDot(cur.n, wo);
        if (measure != EDiscrete && adp != 0) result = result / adp;
        return result;
    }
    case ESensorSample: {
        Vec3 target;
        if (mode == ERadiance && pred->type == ESensorSupernode) target = succ->getPosition();
        else if (mode == EImportance && succ->type == ESensorSupernode) target = pred->getPosition();
        else return RGB(0.0);
        Vec3 wo = normalize(target - cur.p);
        int m = measure == EArea ? ESolidAngle : measure;
        if (m != ESolidAngle) return RGB(0.0);              // perspective.cpp:357-365
        result = RGB(sc.cam.importance(sc.cam.invDir(wo)));
        Float dp = absDot(cur.n, wo);
        if (measure != EDiscrete && dp != 0) result = result / dp;
        return result;
    }
    case ESurfaceInteraction: {
        const Intersection &its = cur.its;
        const Mat bsdf = materialAt(sc, its);
        Vec3 wi = normalize(pred->getPosition() - its.p);
        Vec3 wo = normalize(succ->getPosition() - its.p);
        BSDFRecord bRec(its.toLocal(wi), its.toLocal(wo), mode);
        if (measure == EArea) measure = ESolidAngle;
        result = bsdfEval(bsdf, bRec, measure);
        Float wiDotGeoN = dot(its.ng, wi), woDotGeoN = dot(its.ng, wo);
        if (wiDotGeoN * Frame::cosTheta(bRec.wi) <= 0 || woDotGeoN * Frame::cosTheta(bRec.wo) <= 0) return RGB(0.0);
        if (mode == EImportance)
            result *= std::abs((Frame::cosTheta(bRec.wi) * woDotGeoN) / (Frame::cosTheta(bRec.wo) * wiDotGeoN));
        if (measure != EDiscrete && Frame::cosTheta(bRec.wo) != 0) result = result / std::abs(Frame::cosTheta(bRec.wo));
        return result;
    }
    }
    return result;
}

// vertex.cpp:1094-1205
inline Float vertexEvalPdf(const Scene &sc, const PathVertex &cur, const PathVertex *pred, const PathVertex *succ, int mode,
                           int measure = EArea) {
    Vec3 wo; Float dist = 0.0, result = 0.0;
    switch (cur.type) {
    case EEmitterSupernode:
        if (mode != EImportance || pred != nullptr || succ->type != EEmitterSample) return 0.0;
        return pdfEmitterPosition(sc, succ->emitter);
    case ESensorSupernode:
        if (mode != ERadiance || pred != nullptr || succ->type != ESensorSample) return 0.0;
        return measure == EDiscrete ? 1.0 : 0.0;            // perspective.cpp:314-316
    case EEmitterSample: {
        if (mode == ERadiance && succ->type == EEmitterSupernode) return 1.0;
        else if (mode != EImportance || pred->type != EEmitterSupernode) return 0.0;
        wo = succ->getPosition() - cur.p; dist = length(wo); wo /= dist;
        int m = measure == EArea ? ESolidAngle : measure;
        Float dp = dot(wo, cur.n);                          // area.cpp:150-158
        if (m != ESolidAngle || dp < 0) dp = 0.0;
        result = INV_PI * dp;
        break;
    }
    case ESensorSample: {
        if (mode == EImportance && succ->type == ESensorSupernode) return 1.0;
        else if (mode != ERadiance || pred->type != ESensorSupernode) return 0.0;
        wo = succ->getPosition() - cur.p; dist = length(wo); wo /= dist;
        int m = measure == EArea ? ESolidAngle : measure;
        result = (m != ESolidAngle) ? 0.0 : sc.cam.importance(sc.cam.invDir(wo));   // perspective.cpp:347-355
        break;
    }
    case ESurfaceInteraction: {
        const Intersection &its = cur.its;
        const Mat bsdf = materialAt(sc, its);
        wo = succ->getPosition() - its.p; dist = length(wo); wo /= dist;
        Vec3 wi = normalize(pred->getPosition() - its.p);
        BSDFRecord bRec(its.toLocal(wi), its.toLocal(wo), mode);
        result = bsdfPdf(bsdf, bRec, measure == EArea ? ESolidAngle : measure);
        Float wiDotGeoN = dot(its.ng, wi), woDotGeoN = dot(its.ng, wo);
        if (wiDotGeoN * Frame::cosTheta(bRec.wi) <= 0 || woDotGeoN * Frame::cosTheta(bRec.wo) <= 0) return 0.0;
        break;
    }
    default:
        return 0.0;
    }
    if (measure == EArea) {
        result /= dist * dist;
        if (succ->isOnSurface()) result *= absDot(wo, succ->getGeometricNormal());
    }
    return result;
}

// vertex.cpp:1384-1404 (desired = EEmitterSample; a pinhole has no shape so ESensorSample never succeeds)
inline bool vertexCastEmitter(PathVertex &v) {
    if (v.type == EEmitterSample) return true;
    if (v.type != ESurfaceInteraction) return false;
    if (v.its.emitter < 0) return false;
    v.type = EEmitterSample;
    v.p = v.its.p; v.n = v.its.sh.n; v.uv = v.its.uv;   // records.inl:154-155
    v.emitter = v.its.emitter;
    v.measure = EArea;
    v.degenerate = false;
    return true;
}
inline bool vertexCastSensor(PathVertex &v) {
    if (v.type == ESensorSample) return true;
    return false;   // vertex.cpp:1405-1413: its.shape->getSensor() is NULL for every mesh
}

// vertex.cpp:1578-1586
inline bool vertexGetSamplePosition(const Scene &sc, const PathVertex &sensorSample, const PathVertex &v, Vec2 &result) {
    return sc.cam.getSamplePosition(v.getPosition() - sensorSample.getPosition(), result);
}

// edge.cpp:558-690 (vacuum, no ENull surfaces: any hit is an occluder)
inline bool pathConnectAndCollapse(PathCtx &ctx, PathEdge &edge, const PathVertex &vs, const PathVertex &vt, int &interactions) {
    const Scene &sc = *ctx.scene;
    if (vs.isEmitterSupernode() || vt.isSensorSupernode()) {
        Float radianceTransport = vt.isSensorSupernode() ? 1.0 : 0.0, importanceTransport = 1 - radianceTransport;
        edge.length = 0.0; edge.d = Vec3(0.0);
        edge.pdf[ERadiance] = radianceTransport; edge.pdf[EImportance] = importanceTransport;
        edge.weight[ERadiance] = RGB(radianceTransport); edge.weight[EImportance] = RGB(importanceTransport);
        interactions = 0;
    } else {
        Vec3 vsp = vs.getPosition(), vtp = vt.getPosition();
        edge.d = vsp - vtp;
        edge.length = length(edge.d);
        interactions = 0;
        if (edge.length == 0) return false;
        edge.d /= edge.length;
        Float lengthFactor = vs.isOnSurface() ? (1 - sc.shadowEpsilon) : 1;
        Ray ray; ray.o = vtp; ray.d = edge.d; ray.mint = vt.isOnSurface() ? sc.epsilon : 0; ray.maxt = edge.length * lengthFactor;
        edge.weight[ERadiance] = edge.weight[EImportance] = RGB(1.0);
        edge.pdf[ERadiance] = edge.pdf[EImportance] = 1.0;
        if (sc.rayIntersectShadow(ray, &ctx.rays)) return false;
    }
    edge.d = -edge.d;
    return true;
}

// edge.cpp:221-271 with what = EGeneralizedGeometricTerm (ECosineImp|ECosineRad|EInverseSquareFalloff|ETransmittance)
inline RGB edgeEvalCachedGG(const PathEdge &edge, const PathVertex &pred, const PathVertex &succ) {
    RGB result(1.0);
    if (edge.length == 0) return result;
    if (pred.isOnSurface() && pred.isConnectable()) result *= absDot(pred.getShadingNormal(), edge.d);
    if (succ.isOnSurface() && succ.isConnectable()) result *= absDot(succ.getShadingNormal(), edge.d);
    result = result / (edge.length * edge.length);
    result *= edge.weight[EImportance] * edge.pdf[EImportance];
    return result;
}

// path.cpp:763-1028.  sampleDirect is always false here (MMLT forces it off; BDPT is run with
// directSampling=false); no ENull vertices exist.
inline Float miWeight(const Scene &sc, Path &emitterSubpath, const PathEdge *connectionEdge, Path &sensorSubpath,
                      int s, int t, bool lightImage) {
    int k = s + t + 1, n = k + 1;
    const PathVertex *vsPred = emitterSubpath.vertexOrNull(s - 1), *vtPred = sensorSubpath.vertexOrNull(t - 1),
                     *vs = &emitterSubpath.v[s], *vt = &sensorSubpath.v[t];
    std::vector<Float> pdfImp(n), pdfRad(n);
    std::vector<char> connectable(n);
    int pos = 0;
    for (int i = 0; i <= s; ++i) connectable[pos++] = emitterSubpath.v[i].isConnectable();
    for (int i = t; i >= 0; --i) connectable[pos++] = sensorSubpath.v[i].isConnectable();

    pos = 0;
    pdfImp[pos++] = 1.0;
    for (int i = 0; i < s; ++i) pdfImp[pos++] = emitterSubpath.v[i].pdf[EImportance] * emitterSubpath.e[i].pdf[EImportance];
    pdfImp[pos++] = vertexEvalPdf(sc, *vs, vsPred, vt, EImportance, EArea) * connectionEdge->pdf[EImportance];
    if (t > 0) {
        pdfImp[pos++] = vertexEvalPdf(sc, *vt, vs, vtPred, EImportance, EArea) * sensorSubpath.e[t - 1].pdf[EImportance];
        for (int i = t - 1; i > 0; --i) pdfImp[pos++] = sensorSubpath.v[i].pdf[EImportance] * sensorSubpath.e[i - 1].pdf[EImportance];
    }
    pos = 0;
    if (s > 0) {
        for (int i = 0; i < s - 1; ++i) pdfRad[pos++] = emitterSubpath.v[i + 1].pdf[ERadiance] * emitterSubpath.e[i].pdf[ERadiance];
        pdfRad[pos++] = vertexEvalPdf(sc, *vs, vt, vsPred, ERadiance, EArea) * emitterSubpath.e[s - 1].pdf[ERadiance];
    }
    pdfRad[pos++] = vertexEvalPdf(sc, *vt, vtPred, vs, ERadiance, EArea) * connectionEdge->pdf[ERadiance];
    for (int i = t; i > 0; --i) pdfRad[pos++] = sensorSubpath.v[i - 1].pdf[ERadiance] * sensorSubpath.e[i - 1].pdf[ERadiance];
    pdfRad[pos++] = 1.0;

    auto vertexAt = [&](int i) -> const PathVertex * { return i <= s ? &emitterSubpath.v[i] : &sensorSubpath.v[k - i]; };
    // specular chains: area -> projected solid angle (path.cpp:875-899)
    for (int i = 1; i <= k - 3; ++i) {
        if (i == s || !(connectable[i] && !connectable[i + 1])) continue;
        const PathVertex *cur = vertexAt(i), *succ = vertexAt(i + 1);
        const PathEdge *edge = i < s ? &emitterSubpath.e[i] : &sensorSubpath.e[k - i - 1];
        pdfImp[i + 1] *= edge->length * edge->length / std::abs(
            (succ->isOnSurface() ? dot(edge->d, succ->getGeometricNormal()) : 1) *
            (cur->isOnSurface() ? dot(edge->d, cur->getGeometricNormal()) : 1));
    }
    for (int i = k - 1; i >= 3; --i) {
        if (i - 1 == s || !(connectable[i] && !connectable[i - 1])) continue;
        const PathVertex *cur = vertexAt(i), *succ = vertexAt(i - 1);
        const PathEdge *edge = i <= s ? &emitterSubpath.e[i - 1] : &sensorSubpath.e[k - i];
        pdfRad[i - 1] *= edge->length * edge->length / std::abs(
            (succ->isOnSurface() ? dot(edge->d, succ->getGeometricNormal()) : 1) *
            (cur->isOnSurface() ? dot(edge->d, cur->getGeometricNormal()) : 1));
    }

    double weight = 1, pdf = 1.0;   // power heuristic in ratio form, double precision (path.cpp:979-1025)
    for (int i = s + 1; i < k; ++i) {
        double next = pdf * (double) pdfImp[i] / (double) pdfRad[i], value = next;
        int tPrime = k - i - 1;
        if (connectable[i] && connectable[i + 1] && (lightImage || tPrime > 1)) weight += value * value;
        pdf = next;
    }
    pdf = 1.0;
    for (int i = s - 1; i >= 0; --i) {
        double next = pdf * (double) pdfRad[i + 1] / (double) pdfImp[i + 1], value = next;
        int tPrime = k - i - 1;
        if (connectable[i] && connectable[i + 1] && (lightImage || tPrime > 1)) weight += value * value;
        pdf = next;
    }
    return (Float) (1.0 / weight);
}

// ------------------------------------------------------------------ SplatList (pathsampler.h:317-380)
struct SplatList {
    std::vector<std::pair<Vec2, RGB>> splats;
    Float luminance = 0;
    int nSamples = 0;
    int s = -1, t = -1;
    Float misWeight = 0;   // diagnostic (MMLT single-strategy weight)
    void append(const Vec2 &pos, const RGB &value) { splats.push_back({ pos, value }); luminance += value.luminance(); ++nSamples; }
    void accum(size_t i, const RGB &value) { splats[i].second += value; luminance += value.luminance(); ++nSamples; }
    void clear() { luminance = 0; nSamples = 0; splats.clear(); misWeight = 0; }
    size_t size() const { return splats.size(); }
    // pathsampler.cpp:1001-1028; importanceMap (two-stage MLT): w*h floats or null
    void normalize(const float *importanceMap = nullptr, int w = 0, int h = 0) {
        if (importanceMap) {
            luminance = 0.0;
            for (auto &sp : splats) {
                if (sp.second.isZero()) continue;
                int x = std::min(std::max(0, (int) sp.first.x), w - 1), y = std::min(std::max(0, (int) sp.first.y), h - 1);
                Float lumValue = importanceMap[x + (size_t) y * w];
                Float recip = 1.0 / lumValue;             // Spectrum::operator/= (spectrum.h:447-456)
                sp.second *= recip;
                luminance += sp.second.luminance();
            }
        }
        if (luminance > 0) {
            Float inv = 1.0 / luminance;
            for (auto &sp : splats) sp.second *= inv;
        }
    }
};

struct PathSamplerConfig {
    int technique;       // dr_technique
    int maxDepth, rrDepth;
    bool excludeDirectIllum;   // separateDirect
    bool lightImage;
};

// MIPathTracer::Li (src/integrators/path/path.cpp:123-312) with the plugin defaults the PathSampler
// uses: strictNormals=false, hideEmitters=false, minDepth=0, directTracing=false
inline RGB pathTracerLi(PathCtx &ctx, Sampler *sampler, Ray ray, int maxDepth, int rrDepth, bool excludeDirect) {
    const Scene &sc = *ctx.scene;
    // rRec.type: ERadiance, or ERadiance & ~(EDirectSurfaceRadiance|EEmittedRadiance) (pathsampler.cpp:558-561);
    // after the first bounce it becomes ERadianceNoEmission (path.cpp:293), which has the direct bit set again.
    bool typeEmitted = !excludeDirect, typeDirect = !excludeDirect;
    Intersection its;
    RGB Li(0.0);
    bool non_specular = false;
    int depth = 1;
    sc.rayIntersect(ray, its, &ctx.rays);
    RGB throughput(1.0);
    Float eta = 1.0;
    while (depth <= maxDepth || maxDepth < 0) {
        if (!its.valid()) break;
        const Mat bsdf = materialAt(sc, its);
        if (its.emitter >= 0 && typeEmitted && non_specular) {
            if (dot(its.sh.n, -ray.d) > 0) Li += throughput * sc.emitters[its.emitter].radiance;
        }
        if (depth >= maxDepth && maxDepth > 0) break;

        DirectRec dRec;
        dRec.ref = its.p;
        dRec.refN = bsdfTransmissiveOrBackside(bsdf) ? Vec3(0.0) : its.sh.n;
        if (typeDirect && bsdfHasSmooth(bsdf)) {
            RGB value = sampleEmitterDirect(ctx, dRec, sampler->next2D(), true);
            if (!value.isZero()) {
                BSDFRecord bRec(its.wi, its.toLocal(dRec.d), ERadiance);
                const RGB bsdfVal = bsdfEval(bsdf, bRec);
                if (!bsdfVal.isZero()) {
                    Float bPdf = (dRec.measure == ESolidAngle) ? bsdfPdf(bsdf, bRec) : 0;
                    Float weight = (dRec.pdf * dRec.pdf) / (dRec.pdf * dRec.pdf + bPdf * bPdf);
                    Li += throughput * value * bsdfVal * weight;
                }
            }
        }
        Float bsdfPdfV;
        BSDFRecord bRec(its.wi, ERadiance);
        const Vec2 bsdfPoint = sampler->next2D();
        const Float bsdfExtra = bsdfUsesSampler(bsdf) ? sampler->next1D() : 0.5;      // bRec.sampler->next1D() (roughdielectric.cpp:555)
        RGB bsdfWeight = bsdfSample(bsdf, bRec, bsdfPdfV, bsdfPoint, sc.epsilon, bsdfExtra);
        if (bsdfWeight.isZero()) break;
        non_specular |= !(bRec.sampledType & EDelta);
        const Vec3 wo = its.toWorld(bRec.wo);
        bool hitEmitter = false;
        RGB value;
        ray = sc.makeRay(its.p, wo);
        if (sc.rayIntersect(ray, its, &ctx.rays)) {
            if (its.emitter >= 0) {
                value = (dot(its.sh.n, -ray.d) > 0) ? sc.emitters[its.emitter].radiance : RGB(0.0);   // area.cpp:112-117
                dRec.p = its.p; dRec.n = its.sh.n; dRec.measure = ESolidAngle;                     // records.inl:170-178
                dRec.emitter = its.emitter; dRec.d = ray.d; dRec.dist = its.t;
                hitEmitter = true;
            }
        } else {
            break;
        }
        throughput *= bsdfWeight;
        eta *= bRec.eta;
        if (hitEmitter && typeDirect) {
            const Float lumPdf = (!(bRec.sampledType & EDelta)) ? pdfEmitterDirect(sc, dRec) : 0;
            if (non_specular) {
                Float w = (bsdfPdfV * bsdfPdfV) / (bsdfPdfV * bsdfPdfV + lumPdf * lumPdf);
                Li += throughput * value * w;
            }
        }
        if (!its.valid()) break;
        typeEmitted = false; typeDirect = true;   // rRec.type = ERadianceNoEmission
        if (depth++ >= rrDepth) {
            Float q = std::min(throughput.max() * eta * eta, (Float) 0.95f);
            if (sampler->next1D() >= q) break;
            throughput = throughput / q;
        }
    }
    return Li;
}


// MIDirectIntegrator::Li (src/integrators/direct/direct.cpp:144-305) for a camera ray, with the plugin defaults
// strictNormals=false, hideEmitters=false and emitterSamples = bsdfSamples = shadingSamples; `u` supplies the
// 2-D samples: first the emitter samples, then the BSDF samples.
// `extra` (optional): one more number per BSDF sample for EUsesSampler BSDFs (bRec.sampler->next1D(), roughdielectric.cpp:555)
inline RGB directLi(PathCtx &ctx, Ray ray, int shadingSamples, const Vec2 *u, const Float *extra = nullptr) {
    const Scene &sc = *ctx.scene;
    Intersection its;
    RGB Li(0.0);
    if (!sc.rayIntersect(ray, its, &ctx.rays)) return Li;
    if (its.emitter >= 0 && dot(its.sh.n, -ray.d) > 0) Li += sc.emitters[its.emitter].radiance;   // its.Le(-ray.d), area.cpp:112-117
    const Mat bsdf = materialAt(sc, its);
    const int nE = shadingSamples, nB = shadingSamples;
    const Float fracLum = nE / (Float) (nE + nB), fracBSDF = nB / (Float) (nE + nB), weightLum = 1.0 / nE, weightBSDF = 1.0 / nB;
    auto mi = [](Float a, Float b) { a *= a; b *= b; return a / (a + b); };
    DirectRec dRec;
    dRec.ref = its.p;
    dRec.refN = bsdfTransmissiveOrBackside(bsdf) ? Vec3(0.0) : its.sh.n;
    if (bsdfHasSmooth(bsdf) && !sc.emitters.empty()) {
        for (int i = 0; i < nE; ++i) {
            RGB value = sampleEmitterDirect(ctx, dRec, u[i], true);
            if (value.isZero()) continue;
            BSDFRecord bRec(its.wi, its.toLocal(dRec.d), ERadiance);
            const RGB bsdfVal = bsdfEval(bsdf, bRec);
            if (bsdfVal.isZero()) continue;
            const Float bPdf = bsdfPdf(bsdf, bRec);        // area lights are on a surface
            Li += value * bsdfVal * (mi(dRec.pdf * fracLum, bPdf * fracBSDF) * weightLum);
        }
    }
    for (int i = 0; i < nB; ++i) {
        Float bPdf;
        BSDFRecord bRec(its.wi, ERadiance);
        RGB bsdfVal = bsdfSample(bsdf, bRec, bPdf, u[nE + i], sc.epsilon, extra ? extra[i] : 0.5);
        if (bsdfVal.isZero()) continue;
        const Vec3 wo = its.toWorld(bRec.wo);
        Ray bsdfRay = sc.makeRay(its.p, wo);
        Intersection bIts;
        if (!sc.rayIntersect(bsdfRay, bIts, &ctx.rays)) continue;
        if (bIts.emitter < 0) continue;
        const RGB value = (dot(bIts.sh.n, -bsdfRay.d) > 0) ? sc.emitters[bIts.emitter].radiance : RGB(0.0);
        dRec.p = bIts.p; dRec.n = bIts.sh.n; dRec.measure = ESolidAngle;      // dRec.setQuery (records.inl:170-178)
        dRec.emitter = bIts.emitter; dRec.d = bsdfRay.d; dRec.dist = bIts.t;
        const Float lumPdf = (!(bRec.sampledType & EDelta)) ? pdfEmitterDirect(sc, dRec) : 0;
        Li += value * bsdfVal * (mi(bPdf * fracBSDF, lumPdf * fracLum) * weightBSDF);
    }
    return Li;
}

} // namespace orc
