// path.cuh -- u -> path -> (pixel, RGB, luminance): the device restructuring of
// PathSampler::sampleSplats (src/libbidir/pathsampler.cpp:79-571) for technique = path and mmlt.
//
// The reference builds two pool-allocated Path objects (~300 B per vertex) and then sweeps them
// three times (weights, connection, miWeight).  Here a subpath is walked once: only the current
// vertex, its predecessor's position/normal and four small per-path arrays (pdfImp, pdfRad, the
// area<->projected-solid-angle conversion factor of every edge, a connectable bit mask) survive,
// which is exactly what Path::miWeight (src/libbidir/path.cpp:763-1028) consumes.
// Bookkeeping conventions follow SURVEY.md Appendix E.
#pragma once
#include "traverse.cuh"
#include "bsdf.cuh"
#include "pss.cuh"

#define DR_MAXK 14            // maxDepth + 3 <= DR_MAXK  (maxDepth <= 11)

struct PathCfg {
    int technique, maxDepth, rrDepth;
    int excludeDirect;        // separateDirect (directSamples >= 0)
    int lightImage;
};

struct PathResult {           // single-splat techniques (path, mmlt)
    float lum;                // SplatList::luminance, un-normalised
    int n;                    // number of splats (0 or 1)
    float2 pos;
    float3 val;
    int s, t;
    float mis;
};

enum { V_EMITTER_SAMPLE = 3, V_SENSOR_SAMPLE = 4, V_SURFACE = 5 };

struct Vtx {
    float3 p, ng, ns, ss;     // position, geometric normal, shading normal, shading tangent s (t = ns x ss)
    int mat, emitter;
    int type;
    bool degenerate;
};

DR_D float3 to_local(const Vtx &v, float3 w) { return f3(dot(w, v.ss), dot(w, cross(v.ns, v.ss)), dot(w, v.ns)); }
DR_D float3 to_world(const Vtx &v, float3 w) { return v.ss * w.x + cross(v.ns, v.ss) * w.y + v.ns * w.z; }

// ---------------------------------------------------------------- camera (src/sensors/perspective.cpp)
DR_D float3 cam_xform_dir(const DevCamera &c, float3 v) {
    return f3(c.m[0] * v.x + c.m[1] * v.y + c.m[2] * v.z, c.m[4] * v.x + c.m[5] * v.y + c.m[6] * v.z, c.m[8] * v.x + c.m[9] * v.y + c.m[10] * v.z);
}
DR_D float3 cam_inv_dir(const DevCamera &c, float3 v) {
    return f3(c.m[0] * v.x + c.m[4] * v.y + c.m[8] * v.z, c.m[1] * v.x + c.m[5] * v.y + c.m[9] * v.z, c.m[2] * v.x + c.m[6] * v.y + c.m[10] * v.z);
}
DR_D float3 cam_sample_to_dir(const DevCamera &c, float sx, float sy) {   // :150-157, :336-339
    return normalize(f3((1.f - 2.f * sx) * c.tanHalf, (1.f - 2.f * sy) * c.tanHalf / c.aspect, 1.0f));
}
DR_D float cam_importance(const DevCamera &c, float3 d) {   // :191-245
    if (d.z <= 0.f) return 0.0f;
    float inv = 1.0f / d.z;
    float px = d.x * inv, py = d.y * inv;
    if (px < -c.rectX || px > c.rectX || py < -c.rectY || py > c.rectY) return 0.0f;
    return c.normalization * inv * inv * inv;
}
DR_D bool cam_sample_position(const DevCamera &c, float3 dWorld, float2 &pos) {   // :367-385
    float3 l = cam_inv_dir(c, dWorld);
    if (l.z <= 0.f) return false;
    float sx = 0.5f * (1.f - l.x / (l.z * c.tanHalf));
    float sy = 0.5f * (1.f - l.y * c.aspect / (l.z * c.tanHalf));
    if (sx < 0.f || sx > 1.f || sy < 0.f || sy > 1.f) return false;
    pos = make_float2(sx * c.resX, sy * c.resY);
    return true;
}

// ---------------------------------------------------------------- hit -> vertex (skdtree.h:343-426)
DR_D void fill_vertex(const DevScene &sc, const Hit &hit, Vtx &v) {
    const float4 *tp = sc.tris + 3 * (size_t) hit.tri;
    const float4 t0 = ldg4(tp), t1 = ldg4(tp + 1), t2 = ldg4(tp + 2);
    const float3 p0 = f3(t0.x, t0.y, t0.z), e1 = f3(t0.w, t1.x, t1.y), e2 = f3(t1.z, t1.w, t2.x);
    const uint32_t mf = (uint32_t) __float_as_int(t2.z);
    v.p = p0 + e1 * hit.u + e2 * hit.v;
    float3 face = cross(e1, e2);
    float len = length(face);
    if (len != 0.f) face = face / len;
    if (mf & 0x80000000u) {
        const float4 *np = sc.normals + 3 * (size_t) hit.tri;
        const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
        const float3 n0 = f3(a.x, a.y, a.z), n1 = f3(a.w, b.x, b.y), n2 = f3(b.z, b.w, c.x);
        v.ns = normalize(n0 * (1.f - hit.u - hit.v) + n1 * hit.u + n2 * hit.v);
        if (dot(face, v.ns) < 0.f) face = -face;
    } else {
        v.ns = face;
    }
    v.ng = face;
    v.ss = normalize(e1 - v.ns * dot(v.ns, e1));    // computeShadingFrame, dpdu = p1 - p0
    v.mat = (int) (mf & 0x00ffffffu);
    v.emitter = __float_as_int(t2.w);
    v.type = V_SURFACE;
}

// Scene::rayIntersect -> closest hit with the adaptive epsilon; counts the ray
DR_D bool trace_closest(const DevScene &sc, float3 o, float3 d, float mint, float maxt, Hit &hit, uint32_t &rays) {
    ++rays;
    return traverse<false>(sc, o, d, adaptive_mint(sc, o, mint), maxt, hit);
}
DR_D bool trace_shadow(const DevScene &sc, float3 o, float3 d, float mint, float maxt, uint32_t &rays) {
    ++rays;
    Hit hit;
    return traverse<true>(sc, o, d, adaptive_mint(sc, o, mint), maxt, hit);
}

// ---------------------------------------------------------------- emitters
// DiscreteDistribution::sample + sampleReuse (include/mitsuba/core/pmf.h:134-178) on a double CDF
DR_D int cdf_sample_reuse(const double *cdf, int n, float &v, float &pdf) {
    // lower_bound over cdf[0..n]: first entry >= v, minus one, clamped to [0, n-1]
    const double dv = (double) v;
    int lo = 0, hi = n + 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] < dv) lo = mid + 1; else hi = mid; }
    int index = min(n - 1, max(0, lo - 1));
    while (cdf[index + 1] - cdf[index] == 0.0 && index < n) ++index;
    const double a = cdf[index], b = cdf[index + 1];
    pdf = (float) (b - a);
    v = (float) ((dv - a) / (b - a));
    return index;
}

struct EmitterPoint { float3 p, n; int emitter; float pdfArea; float emPdf; };

// Scene::sampleEmitterPosition (scene.cpp:1066-1082) -> TriMesh::samplePosition (trimesh.cpp:429-440)
// -> Triangle::sample (triangle.cpp:24-60)
DR_D void sample_emitter_point(const DevScene &sc, float sx, float sy, EmitterPoint &ep) {
    float emPdf;
    const int e = cdf_sample_reuse(sc.emitterCdf, sc.nEmitters, sx, emPdf);
    const DevEmitter &em = sc.emitters[e];
    float triPdf;
    const int tri = cdf_sample_reuse(sc.emCdf + em.cdfOffset, (int) em.nTris, sy, triPdf);
    const float4 *tp = sc.emTris + 6 * (size_t) (em.firstEmTri + tri);
    const float4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
    const float3 p0 = f3(a.x, a.y, a.z), e1 = f3(a.w, b.x, b.y), e2 = f3(b.z, b.w, c.x);
    const float sq = safe_sqrtf(1.0f - sx);
    const float bx = 1.f - sq, by = sq * sy;                 // squareToUniformTriangle
    ep.p = p0 + e1 * bx + e2 * by;
    if (__float_as_int(c.y)) {
        const float4 d = ldg4(tp + 3), f = ldg4(tp + 4), g = ldg4(tp + 5);
        const float3 n0 = f3(d.x, d.y, d.z), n1 = f3(d.w, f.x, f.y), n2 = f3(f.z, f.w, g.x);
        ep.n = normalize(n0 * (1.0f - bx - by) + n1 * bx + n2 * by);
    } else {
        ep.n = normalize(cross(e1, e2));
    }
    ep.emitter = e; ep.pdfArea = em.invArea * emPdf; ep.emPdf = emPdf;
}
DR_D float3 emitter_radiance(const DevScene &sc, int e) { const DevEmitter &em = sc.emitters[e]; return f3(em.radiance[0], em.radiance[1], em.radiance[2]); }

// ---------------------------------------------------------------- surface vertex helpers
// PathVertex::eval for ESurfaceInteraction (vertex.cpp:1026-1063), measure = EArea -> ESolidAngle
DR_D float3 surface_eval(const DevScene &sc, const Vtx &v, const Mat &m, float3 wiW, float3 woW, int mode) {
    const float3 wi = to_local(v, wiW), wo = to_local(v, woW);
    float3 r = bsdf_eval(m, wi, wo, mode, MEAS_SOLID_ANGLE);
    const float wiDotGeoN = dot(v.ng, wiW), woDotGeoN = dot(v.ng, woW);
    if (wiDotGeoN * wi.z <= 0.f || woDotGeoN * wo.z <= 0.f) return f3(0.f);
    if (mode == MODE_IMPORTANCE) r *= fabsf((wi.z * woDotGeoN) / (wo.z * wiDotGeoN));
    if (wo.z != 0.f) r = r / fabsf(wo.z);
    return r;
}
// PathVertex::evalPdf for ESurfaceInteraction in area measure (vertex.cpp:1150-1204)
DR_D float surface_pdf_area(const Vtx &v, const Mat &m, float3 fromPos, float3 toPos, float3 toNg) {
    float3 woW = toPos - v.p;
    const float dist = length(woW);
    woW = woW / dist;
    const float3 wiW = normalize(fromPos - v.p);
    const float3 wi = to_local(v, wiW), wo = to_local(v, woW);
    float r = bsdf_pdf(m, wi, wo, MEAS_SOLID_ANGLE);
    if (dot(v.ng, wiW) * wi.z <= 0.f || dot(v.ng, woW) * wo.z <= 0.f) return 0.f;
    return r / (dist * dist) * absdot(woW, toNg);
}

// One BSDF sampling step of a random walk (vertex.cpp:153-271 + :334-347), shared by both subpaths.
struct WalkStep { float3 wo; float3 weightFwd; float pdfFwd, pdfBwd; bool delta; float eta; };
DR_D bool surface_sample_next(const DevScene &sc, const Vtx &v, const Mat &m, float3 wiW, int mode, float2 u, WalkStep &ws) {
    const float3 wi = to_local(v, wiW);
    BsdfSample bs;
    bsdf_sample(m, wi, mode, u.x, u.y, sc.epsilon, bs);
    if (is_zero(bs.weight)) return false;
    const int measure = (bs.sampledType & BT_DELTA) ? MEAS_DISCRETE : MEAS_SOLID_ANGLE;
    ws.wo = to_world(v, bs.wo);
    const float wiDotGeoN = dot(v.ng, wiW), woDotGeoN = dot(v.ng, ws.wo);
    if (wiDotGeoN * wi.z <= 0.f || woDotGeoN * bs.wo.z <= 0.f) return false;
    ws.pdfFwd = bs.pdf;
    ws.pdfBwd = bsdf_pdf(m, bs.wo, wi, measure);          // bRec.reverse()
    if (ws.pdfBwd <= DR_RCPOVERFLOW) return false;
    ws.weightFwd = bs.weight;
    // adjoint BSDF for shading normals acts on weight[EImportance] only (vertex.cpp:252-262)
    if (mode == MODE_IMPORTANCE) ws.weightFwd *= fabsf((wi.z * woDotGeoN) / (bs.wo.z * wiDotGeoN));
    ws.delta = measure == MEAS_DISCRETE;
    ws.eta = bs.eta;
    return true;
}

// ---------------------------------------------------------------- MIS (path.cpp:763-1028, no direct sampling / ENull)
struct MisArrays {
    float pdfImp[DR_MAXK + 1], pdfRad[DR_MAXK + 1];
    float conv[DR_MAXK + 1];       // conv[g]: len^2 / |cos_g cos_{g+1}| (geometric normals) of edge (g, g+1)
    uint32_t connectable;          // bit g
};
DR_D float mis_weight(MisArrays &A, int s, int t, bool lightImage) {
    const int k = s + t + 1;
    // entering specular chains: area density -> projected solid angle (path.cpp:875-899)
    for (int i = 1; i <= k - 3; ++i) {
        if (i == s || !(((A.connectable >> i) & 1u) && !((A.connectable >> (i + 1)) & 1u))) continue;
        A.pdfImp[i + 1] *= A.conv[i];
    }
    for (int i = k - 1; i >= 3; --i) {
        if (i - 1 == s || !(((A.connectable >> i) & 1u) && !((A.connectable >> (i - 1)) & 1u))) continue;
        A.pdfRad[i - 1] *= A.conv[i - 1];
    }
    double weight = 1.0, pdf = 1.0;
    for (int i = s + 1; i < k; ++i) {
        const double next = pdf * (double) A.pdfImp[i] / (double) A.pdfRad[i];
        const int tPrime = k - i - 1;
        if (((A.connectable >> i) & 1u) && ((A.connectable >> (i + 1)) & 1u) && (lightImage || tPrime > 1)) weight += next * next;
        pdf = next;
    }
    pdf = 1.0;
    for (int i = s - 1; i >= 0; --i) {
        const double next = pdf * (double) A.pdfRad[i + 1] / (double) A.pdfImp[i + 1];
        const int tPrime = k - i - 1;
        if (((A.connectable >> i) & 1u) && ((A.connectable >> (i + 1)) & 1u) && (lightImage || tPrime > 1)) weight += next * next;
        pdf = next;
    }
    return (float) (1.0 / weight);
}

// ---------------------------------------------------------------- MMLT (pathsampler.cpp:84-320)
DR_D void eval_mmlt(const DevScene &sc, const PathCfg &pc, Pss &pss, int depth, PathResult &out, uint32_t &rays) {
    out.lum = 0.f; out.n = 0; out.val = f3(0.f); out.pos = make_float2(0.f, 0.f); out.mis = 0.f;
    int s, t, nStrats;
    const float decision = pss.next1D(SMP_DIRECT);
    if (pc.lightImage) { nStrats = depth + 1; s = min((int) (nStrats * decision), nStrats - 1); t = nStrats - s; }
    else { nStrats = depth; s = min((int) (nStrats * decision), nStrats - 1); t = 1 + (nStrats - s); }
    out.s = s; out.t = t;
    if (depth == 1) return;
    const int k = s + t + 1;
    MisArrays A;
    A.connectable = 0;
    bool anyConnectable = false;      // some vertex with index >= 2 on either subpath is connectable
    float3 weight = f3(1.f);

    // ---- sensor subpath: vertices 0 (supernode, g = k), 1 (sensor sample, g = k-1), 2.. (surfaces)
    Vtx vt, vtPred;                   // vertex t and t-1 (vtPred only valid when t >= 2)
    float3 firstHit = f3(0.f);
    {
        (void) pss.next2D(SMP_SENSOR);                     // sampleSensorPosition consumes 2 (vertex.cpp:79)
        A.pdfRad[k] = 1.0f;
        A.pdfRad[k - 1] = 1.0f;                            // supernode pdf[ERadiance] (perspective.cpp:305)
        vt.p = sc.cam.pos; vt.ng = vt.ns = sc.cam.dir; vt.type = V_SENSOR_SAMPLE; vt.degenerate = false;
        vt.mat = -1; vt.emitter = -1; vt.ss = f3(0.f);
        A.connectable |= 1u << (k - 1);                    // sensor sample: never discrete, not degenerate
        float3 prevP = vt.p, prevNg = vt.ng;               // vertex j-1 while standing on j
        float prevLen = 0.f;                               // length of the edge (j-1, j)
        float3 prevD = f3(0.f);                            // its direction (world, pointing j-1 -> j)
        for (int j = 1; j < t; ++j) {                      // sampleNext from vertex j creates vertex j+1
            const int g = k - j;
            float3 d; float pdfFwdSA, pdfBwd; bool delta = false; float3 wFwd;
            if (j == 1) {                                  // vertex.cpp:126-151, perspective.cpp:318-345
                const float2 u = pss.next2D(SMP_SENSOR);
                const float3 dl = cam_sample_to_dir(sc.cam, u.x, u.y);
                d = cam_xform_dir(sc.cam, dl);
                pdfFwdSA = sc.cam.normalization / (dl.z * dl.z * dl.z);
                pdfBwd = 1.0f; wFwd = f3(1.f);
            } else {
                const Mat m = load_material(sc, vt.mat);
                WalkStep ws;
                if (!surface_sample_next(sc, vt, m, normalize(prevP - vt.p), MODE_RADIANCE, pss.next2D(SMP_SENSOR), ws)) return;
                d = ws.wo; pdfFwdSA = ws.pdfFwd; pdfBwd = ws.pdfBwd; delta = ws.delta; wFwd = ws.weightFwd;
                if (!delta) A.connectable |= (vt.degenerate ? 0u : 1u) << g;
                if (!delta && !vt.degenerate) anyConnectable = true;
            }
            weight *= wFwd;
            Hit hit;
            if (!trace_closest(sc, vt.p, d, sc.epsilon, INFINITY, hit, rays)) return;
            if (hit.t == 0.f) return;
            Vtx nv;
            fill_vertex(sc, hit, nv);
            const Mat nm = load_material(sc, nv.mat);
            nv.degenerate = !(mat_has_smooth(nm.type) || nv.emitter >= 0);
            // solid angle -> area (vertex.cpp:334-347); delta interactions keep their discrete pdfs
            const float cosNext = absdot(d, nv.ng);
            if (!delta) {
                pdfFwdSA = pdfFwdSA / (hit.t * hit.t) * cosNext;
                if (j >= 2) pdfBwd = pdfBwd / (prevLen * prevLen) * absdot(prevD, prevNg);
            }
            A.pdfRad[g - 1] = pdfFwdSA;                    // density of vertex j+1
            A.pdfImp[g + 1] = pdfBwd;                      // density of vertex j-1
            A.conv[g - 1] = hit.t * hit.t / fabsf(absdot(d, vt.ng) * cosNext);   // edge (g-1, g)
            if (j == 1) firstHit = nv.p;
            prevP = vt.p; prevNg = vt.ng; prevLen = hit.t; prevD = d;
            vtPred = vt; vt = nv;
        }
        if (t >= 2) {                                      // last vertex: measure stays invalid => connectable iff !degenerate
            A.connectable |= (vt.degenerate ? 0u : 1u) << (k - t);
            if (!vt.degenerate) anyConnectable = true;
        }
    }

    // ---- emitter subpath: vertices 0 (supernode, g = 0), 1 (emitter sample), 2.. (surfaces)
    Vtx vs, vsPred;
    A.pdfImp[0] = 1.0f;
    A.connectable |= 1u;                                   // area lights: supernode not degenerate, measure never discrete
    if (s >= 1) {
        EmitterPoint ep;
        const float2 u0 = pss.next2D(SMP_EMITTER);
        sample_emitter_point(sc, u0.x, u0.y, ep);
        const DevEmitter &em = sc.emitters[ep.emitter];
        weight *= emitter_radiance(sc, ep.emitter) * (DR_PI * em.area / ep.emPdf);   // m_power / emPdf
        A.pdfImp[1] = ep.pdfArea;
        vs.p = ep.p; vs.ng = vs.ns = ep.n; vs.type = V_EMITTER_SAMPLE; vs.degenerate = false; vs.emitter = ep.emitter; vs.mat = -1;
        vs.ss = f3(0.f);
        A.connectable |= 1u << 1;
        float3 prevP = vs.p, prevNg = vs.ng, prevD = f3(0.f); float prevLen = 0.f;
        for (int i = 1; i < s; ++i) {
            float3 d; float pdfFwdSA, pdfBwd; bool delta = false; float3 wFwd;
            if (i == 1) {                                  // vertex.cpp:99-124, area.cpp:130-138
                const float2 u = pss.next2D(SMP_EMITTER);
                const float3 local = square_to_cosine_hemisphere(u.x, u.y);
                float3 fs, ft;
                coordinate_system(vs.ns, fs, ft);
                d = fs * local.x + ft * local.y + vs.ns * local.z;
                pdfFwdSA = DR_INV_PI * local.z; pdfBwd = 1.0f; wFwd = f3(1.f);
            } else {
                const Mat m = load_material(sc, vs.mat);
                WalkStep ws;
                if (!surface_sample_next(sc, vs, m, normalize(prevP - vs.p), MODE_IMPORTANCE, pss.next2D(SMP_EMITTER), ws)) return;
                d = ws.wo; pdfFwdSA = ws.pdfFwd; pdfBwd = ws.pdfBwd; delta = ws.delta; wFwd = ws.weightFwd;
                if (!delta) A.connectable |= (vs.degenerate ? 0u : 1u) << i;
                if (!delta && !vs.degenerate) anyConnectable = true;
            }
            weight *= wFwd;
            Hit hit;
            if (!trace_closest(sc, vs.p, d, sc.epsilon, INFINITY, hit, rays)) return;
            if (hit.t == 0.f) return;
            Vtx nv;
            fill_vertex(sc, hit, nv);
            const Mat nm = load_material(sc, nv.mat);
            nv.degenerate = !(mat_has_smooth(nm.type) || nv.emitter >= 0);
            const float cosNext = absdot(d, nv.ng);
            if (!delta) {
                pdfFwdSA = pdfFwdSA / (hit.t * hit.t) * cosNext;
                if (i >= 2) pdfBwd = pdfBwd / (prevLen * prevLen) * absdot(prevD, prevNg);
            }
            A.pdfImp[i + 1] = pdfFwdSA;
            A.pdfRad[i - 1] = pdfBwd;
            A.conv[i] = hit.t * hit.t / fabsf(absdot(d, vs.ng) * cosNext);   // edge (i, i+1)
            prevP = vs.p; prevNg = vs.ng; prevLen = hit.t; prevD = d;
            vsPred = vs; vs = nv;
        }
        if (s >= 2) {
            A.connectable |= (vs.degenerate ? 0u : 1u) << s;
            if (!vs.degenerate) anyConnectable = true;
        }
    }
    if (!anyConnectable) return;                            // pathsampler.cpp:161-174

    // ---- connection
    float3 value;
    float2 samplePos = make_float2(0.f, 0.f);
    if (s == 0) {                                           // pure sensor path: vt must be on an emitter (:213-224)
        if (vt.type != V_SURFACE || vt.emitter < 0) return;
        const float3 n = vt.ns;                             // cast(): pRec.n = its.shFrame.n (records.inl:154-155)
        float3 wo = vtPred.p - vt.p;
        const float dist = length(wo);
        wo = wo / dist;
        const float dp = dot(wo, n);
        if (!(dp > 0.f)) return;                            // evalDirection (area.cpp:140-148) / |n.wo| = 1/pi
        value = weight * emitter_radiance(sc, vt.emitter);  // radiance * pi * (1/pi)
        const DevEmitter &em = sc.emitters[vt.emitter];
        A.connectable |= 1u << 1;                           // emitter sample: area measure, not degenerate
        A.pdfImp[1] = em.invArea * em.pdfDiscrete;          // vs->evalPdf: pdfEmitterPosition
        A.pdfImp[2] = DR_INV_PI * dp / (dist * dist) * absdot(wo, vtPred.ng);   // vt->evalPdf(vs, vtPred, EImportance)
        // connection edge of a supernode: length 0, generalized geometric term = 1 (edge.cpp:229-234, 561-571)
    } else {
        if (vs.degenerate || vt.degenerate) return;         // :253-257
        float3 d = vs.p - vt.p;                             // from vt towards vs
        const float len = length(d);
        if (len == 0.f) return;
        d = d / len;
        // vs->eval(vsPred, vt, EImportance)
        float3 fs, ft;
        Mat ms, mt;
        if (s == 1) {
            const float dp = dot(-d, vs.ns);
            fs = f3(dp > 0.f ? DR_INV_PI : 0.f);
        } else {
            ms = load_material(sc, vs.mat);
            fs = surface_eval(sc, vs, ms, normalize(vsPred.p - vs.p), -d, MODE_IMPORTANCE);
        }
        if (t == 1) {
            const float imp = cam_importance(sc.cam, cam_inv_dir(sc.cam, d));
            const float dp = absdot(vt.ns, d);
            ft = f3(dp != 0.f ? imp / dp : imp);
        } else {
            mt = load_material(sc, vt.mat);
            ft = surface_eval(sc, vt, mt, normalize(vtPred.p - vt.p), d, MODE_RADIANCE);
        }
        value = weight * fs * ft;
        if (is_zero(value)) return;
        // pathConnectAndCollapse (edge.cpp:572-606): vt and vs are always "on surface" here
        if (trace_shadow(sc, vt.p, d, sc.epsilon, len * (1.f - sc.shadowEpsilon), rays)) return;
        if (pc.excludeDirect && depth <= 2) return;
        // generalized geometric term (edge.cpp:245-267)
        value *= absdot(vs.ns, d) * absdot(vt.ns, d) / (len * len);
        // the four densities next to the connection (path.cpp:835-859)
        A.connectable |= (1u << s) | (1u << (s + 1));       // measure forced to EArea (:263-265)
        if (s == 1) {
            const float dp = dot(-d, vs.ns);
            A.pdfImp[s + 1] = DR_INV_PI * fmaxf(dp, 0.f) / (len * len) * absdot(d, vt.ng);
            A.pdfRad[s - 1] = 1.0f;
        } else {
            A.pdfImp[s + 1] = surface_pdf_area(vs, ms, vsPred.p, vt.p, vt.ng);
            A.pdfRad[s - 1] = surface_pdf_area(vs, ms, vt.p, vsPred.p, vsPred.ng);
        }
        if (t == 1) {
            A.pdfRad[s] = cam_importance(sc.cam, cam_inv_dir(sc.cam, d)) / (len * len) * absdot(d, vs.ng);
            A.pdfImp[s + 2] = 1.0f;
        } else {
            A.pdfRad[s] = surface_pdf_area(vt, mt, vtPred.p, vs.p, vs.ng);
            A.pdfImp[s + 2] = surface_pdf_area(vt, mt, vs.p, vtPred.p, vtPred.ng);
        }
        if (t == 1 && !cam_sample_position(sc.cam, vs.p - vt.p, samplePos)) return;   // :298-303
    }
    if (s == 0 && pc.excludeDirect && depth <= 2) return;
    const float mis = mis_weight(A, s, t, pc.lightImage != 0);
    value *= mis * (float) nStrats;
    if (t >= 2) cam_sample_position(sc.cam, firstHit - sc.cam.pos, samplePos);           // :309-312
    out.mis = mis;
    out.n = 1; out.pos = samplePos; out.val = value; out.lum = luminance(value);
}

// ---------------------------------------------------------------- unidirectional path tracer
// PathSampler EUnidirectional (pathsampler.cpp:529-567) + MIPathTracer::Li (integrators/path/path.cpp:123-312)
// with strictNormals=false, hideEmitters=false, minDepth=0, directTracing=false.
DR_D void eval_pt(const DevScene &sc, const PathCfg &pc, Pss &pss, PathResult &out, uint32_t &rays) {
    out.s = out.t = -1; out.mis = 0.f;
    const float2 u0 = pss.next2D(SMP_SENSOR);
    const float2 samplePos = make_float2(u0.x * sc.cam.resX, u0.y * sc.cam.resY);
    const float3 dl = cam_sample_to_dir(sc.cam, samplePos.x / sc.cam.resX, samplePos.y / sc.cam.resY);
    const float invZ = 1.0f / dl.z;
    float3 o = sc.cam.pos, d = cam_xform_dir(sc.cam, dl);
    bool typeEmitted = !pc.excludeDirect, typeDirect = !pc.excludeDirect;
    float3 Li = f3(0.f), throughput = f3(1.f);
    float eta = 1.0f;
    bool nonSpecular = false;
    int depth = 1;
    Hit hit;
    Vtx v;
    bool valid = trace_closest(sc, o, d, sc.cam.nearClip * invZ, sc.cam.farClip * invZ, hit, rays);
    if (valid) fill_vertex(sc, hit, v);
    while (depth <= pc.maxDepth || pc.maxDepth < 0) {
        if (!valid) break;
        const Mat m = load_material(sc, v.mat);
        if (v.emitter >= 0 && typeEmitted && nonSpecular && dot(v.ns, -d) > 0.f) Li += throughput * emitter_radiance(sc, v.emitter);
        if (depth >= pc.maxDepth && pc.maxDepth > 0) break;
        const float3 wi = to_local(v, -d);
        const float3 refN = mat_transmissive_or_backside(m) ? f3(0.f) : v.ns;    // records.inl:160-164
        // ---- direct illumination (scene.cpp:879-904, area.cpp:156-170, shape.cpp:102-114)
        if (typeDirect && mat_has_smooth(m.type)) {
            const float2 u = pss.next2D(SMP_SENSOR);
            EmitterPoint ep;
            sample_emitter_point(sc, u.x, u.y, ep);
            float3 dd = ep.p - v.p;
            const float distSq = dot(dd, dd), dist = sqrtf(distSq);
            dd = dd / dist;
            const float dp = absdot(dd, ep.n);
            float pdf = sc.emitters[ep.emitter].invArea * (dp != 0.f ? distSq / dp : 0.f);
            if (dot(dd, refN) >= 0.f && dot(dd, ep.n) < 0.f && pdf != 0.f) {
                if (!trace_shadow(sc, v.p, dd, sc.epsilon, dist * (1.f - sc.shadowEpsilon), rays)) {
                    const float3 value = emitter_radiance(sc, ep.emitter) / pdf / ep.emPdf;
                    pdf *= ep.emPdf;
                    const float3 wo = to_local(v, dd);
                    const float3 bsdfVal = bsdf_eval(m, wi, wo, MODE_RADIANCE, MEAS_SOLID_ANGLE);
                    if (!is_zero(bsdfVal)) {
                        const float bp = bsdf_pdf(m, wi, wo, MEAS_SOLID_ANGLE);
                        Li += throughput * value * bsdfVal * ((pdf * pdf) / (pdf * pdf + bp * bp));
                    }
                }
            }
        }
        // ---- BSDF sampling
        const float2 ub = pss.next2D(SMP_SENSOR);
        BsdfSample bs;
        bsdf_sample(m, wi, MODE_RADIANCE, ub.x, ub.y, sc.epsilon, bs);
        if (is_zero(bs.weight)) break;
        nonSpecular |= !(bs.sampledType & BT_DELTA);
        const float3 refP = v.p;
        o = v.p; d = to_world(v, bs.wo);
        bool hitEmitter = false;
        float3 value = f3(0.f);
        float lumPdf = 0.f;
        valid = trace_closest(sc, o, d, sc.epsilon, INFINITY, hit, rays);
        if (!valid) break;
        fill_vertex(sc, hit, v);
        if (v.emitter >= 0) {
            value = dot(v.ns, -d) > 0.f ? emitter_radiance(sc, v.emitter) : f3(0.f);
            hitEmitter = true;
            // pdfEmitterDirect (scene.cpp:1057-1060, area.cpp:172-180, shape.cpp:116-126) with dRec.setQuery(ray, its)
            if (!(bs.sampledType & BT_DELTA) && dot(d, refN) >= 0.f && dot(d, v.ns) < 0.f) {
                const DevEmitter &em = sc.emitters[v.emitter];
                lumPdf = em.invArea * (hit.t * hit.t) / absdot(d, v.ns) * em.pdfDiscrete;
            }
        }
        (void) refP;
        throughput *= bs.weight;
        eta *= bs.eta;
        if (hitEmitter && typeDirect && nonSpecular)
            Li += throughput * value * ((bs.pdf * bs.pdf) / (bs.pdf * bs.pdf + lumPdf * lumPdf));
        typeEmitted = false; typeDirect = true;            // rRec.type = ERadianceNoEmission
        if (depth++ >= pc.rrDepth) {
            const float q = fminf(max3(throughput) * eta * eta, 0.95f);
            if (pss.next1D(SMP_SENSOR) >= q) break;
            throughput = throughput / q;
        }
    }
    out.n = 1; out.pos = samplePos; out.val = Li; out.lum = luminance(Li);
}
