// path.cuh -- building blocks of u -> path -> (pixel, RGB, luminance): camera, hit -> vertex,
// emitter sampling, BSDF walk step, connection evaluation and the MIS sweep.  The resumable
// evaluators that string them together (one wavefront stage per ray) live in wave.cuh.
//
// The reference builds two pool-allocated Path objects (~300 B per vertex) and then sweeps them
// three times (weights, connection, miWeight; src/libbidir/pathsampler.cpp:79-571).  Here a subpath
// is walked once: only the current vertex, its predecessor's position/normal and four small
// per-path arrays (pdfImp, pdfRad, the area<->projected-solid-angle conversion factor of every
// edge, a connectable bit mask) survive, which is exactly what Path::miWeight
// (src/libbidir/path.cpp:763-1028) consumes.  Bookkeeping conventions follow SURVEY.md Appendix E.
#pragma once
#include "traverse.cuh"
#include "bsdf.cuh"
#include "pss.cuh"

#define DR_MAXK 14            // maxDepth + 3 <= DR_MAXK  (maxDepth <= 11)

struct PathCfg {
    int technique, maxDepth, rrDepth;
    int excludeDirect;        // separateDirect (directSamples >= 0)
    int lightImage;
    int hasRoughDielectric;   // some triangle's BSDF draws an extra number per sample (pssmlt_utils.h:35-45)
    int bdBatch;              // technique=bdpt: all connections of a path in one round (machine.cuh BdConn)
};

struct PathResult {           // single-splat techniques (path, mmlt)
    Real lum;                // SplatList::luminance, un-normalised
    int n;                    // splat 0 exists (0 or 1)
    int nl;                   // bdpt: light-image splats in the lane's in-flight list
    R2 pos;
    R3 val;
    int s, t;
    Real mis;
};

enum { V_EMITTER_SAMPLE = 3, V_SENSOR_SAMPLE = 4, V_SURFACE = 5 };

struct Vtx {                  // 128 bytes (one cache line), stored verbatim in lane memory between wavefront stages
    R3 p, ng, ns, ss;         // position, geometric normal, shading normal, shading tangent s (t = ns x ss)
    int mat, emitter;
    int type;
    int degenerate;
    R2 uv;                    // its.uv: interpolated texture coordinates, or the barycentric pair (skdtree.h:399-405)
};

DR_D R3 to_local(const Vtx &v, R3 w) { return r3(dot(w, v.ss), dot(w, cross(v.ns, v.ss)), dot(w, v.ns)); }
DR_D R3 to_world(const Vtx &v, R3 w) { return v.ss * w.x + cross(v.ns, v.ss) * w.y + v.ns * w.z; }

// ---------------------------------------------------------------- camera (src/sensors/perspective.cpp)
DR_D R3 cam_pos(const DevCamera &c) { return r3(c.pos[0], c.pos[1], c.pos[2]); }
DR_D R3 cam_dir(const DevCamera &c) { return r3(c.dir[0], c.dir[1], c.dir[2]); }
DR_D R3 cam_xform_dir(const DevCamera &c, R3 v) {
    return r3(c.m[0] * v.x + c.m[1] * v.y + c.m[2] * v.z, c.m[4] * v.x + c.m[5] * v.y + c.m[6] * v.z, c.m[8] * v.x + c.m[9] * v.y + c.m[10] * v.z);
}
DR_D R3 cam_inv_dir(const DevCamera &c, R3 v) {
    return r3(c.inv[0] * v.x + c.inv[1] * v.y + c.inv[2] * v.z, c.inv[3] * v.x + c.inv[4] * v.y + c.inv[5] * v.z, c.inv[6] * v.x + c.inv[7] * v.y + c.inv[8] * v.z);
}
DR_D R3 cam_sample_to_dir(const DevCamera &c, Real sx, Real sy) {   // :132-157, :336-339 (sample in [0,1]^2 over the crop window)
    const Real fx = c.relOffX + sx * c.relSizeX, fy = c.relOffY + sy * c.relSizeY;
    return normalize(r3((1. - 2. * fx) * c.tanHalf, (1. - 2. * fy) * c.tanHalf / c.aspect, 1.0));
}
DR_D Real cam_importance(const DevCamera &c, R3 d) {   // :191-245
    if (d.z <= 0.) return 0.0;
    Real inv = 1.0 / d.z;
    Real px = d.x * inv, py = d.y * inv;
    if (px < c.rectMinX || px > c.rectMaxX || py < c.rectMinY || py > c.rectMaxY) return 0.0;
    return c.normalization * inv * inv * inv;
}
DR_D bool cam_sample_position(const DevCamera &c, R3 dWorld, R2 &pos) {   // :367-385
    R3 l = cam_inv_dir(c, dWorld);
    if (l.z <= 0.) return false;
    Real sx = (0.5 * (1. - l.x / (l.z * c.tanHalf)) - c.relOffX) / c.relSizeX;
    Real sy = (0.5 * (1. - l.y * c.aspect / (l.z * c.tanHalf)) - c.relOffY) / c.relSizeY;
    if (sx < 0. || sx > 1. || sy < 0. || sy > 1.) return false;
    pos = r2(sx * c.resX, sy * c.resY);
    return true;
}

// ---------------------------------------------------------------- hit -> vertex (skdtree.h:343-426)
// The float traversal only SELECTS the triangle.  The hit itself is recomputed here in double from
// the double-precision ray (o, d) and the exact float vertices, so that t, the barycentrics and the
// vertex position agree with the reference's double-precision TriAccel test (triaccel.h:91-157).
DR_D void fill_vertex(const DevScene &sc, const Hit &hit, R3 o, R3 d, Vtx &v, Real &tOut) {
    const float4 *tp = sc.tris + 3 * (size_t) hit.tri;
    const float4 t0 = ldg4(tp), t1 = ldg4(tp + 1), t2 = ldg4(tp + 2);
    const R3 p0 = r3(t0.x, t0.y, t0.z), p1 = r3(t0.w, t1.x, t1.y), p2 = r3(t1.z, t1.w, t2.x);
    const uint32_t mf = (uint32_t) __float_as_int(t2.z);
    const R3 e1 = p1 - p0, e2 = p2 - p0;
    Real t = 0., bu = 0., bv = 0.;                  // det == 0 cannot pass the traversal's tests; t = 0 then ends the path
    {
        const R3 pvec = cross(d, e2);
        const Real det = dot(e1, pvec);
        if (det != 0.) {
            const Real inv = 1.0 / det;
            const R3 tvec = o - p0;
            const R3 qvec = cross(tvec, e1);
            bu = dot(tvec, pvec) * inv; bv = dot(d, qvec) * inv; t = dot(e2, qvec) * inv;
        }
    }
    tOut = t;
    v.p = p0 * (1. - bu - bv) + p1 * bu + p2 * bv;
    R3 face = cross(e1, e2);
    Real len = length(face);
    if (len != 0.) face = face / len;
    if (mf & 0x80000000u) {
        const float4 *np = sc.normals + 3 * (size_t) hit.tri;
        const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
        const R3 n0 = r3(a.x, a.y, a.z), n1 = r3(a.w, b.x, b.y), n2 = r3(b.z, b.w, c.x);
        v.ns = normalize(n0 * (1. - bu - bv) + n1 * bu + n2 * bv);
        if (dot(face, v.ns) < 0.) face = -face;
    } else {
        v.ns = face;
    }
    v.ng = face;
    R3 dpdu = e1;                                   // its.dpdu = p1 - p0 unless the mesh carries UV tangents (skdtree.h:373-380)
    v.uv = r2(bu, bv);
    if (mf & DR_MF_HAS_UV) {
        const float4 *up = sc.uvs + 2 * (size_t) hit.tri;
        const float4 ua = ldg4(up), ub = ldg4(up + 1);
        const Real b0 = 1. - bu - bv;
        v.uv = r2((Real) ua.x * b0 + (Real) ua.z * bu + (Real) ub.x * bv, (Real) ua.y * b0 + (Real) ua.w * bu + (Real) ub.y * bv);
        if (mf & DR_MF_UV_TANGENTS) {               // TriMesh::computeUVTangents (trimesh.cpp:741-759), per triangle
            const Real du1 = (Real) ua.z - (Real) ua.x, dv1 = (Real) ua.w - (Real) ua.y, du2 = (Real) ub.x - (Real) ua.x, dv2 = (Real) ub.y - (Real) ua.y;
            const Real det = du1 * dv2 - dv1 * du2;
            if (det == 0.) {
                const R3 n = cross(e1, e2);
                R3 dpdv;
                coordinate_system(n / length(n), dpdu, dpdv);
            } else {
                const Real invDet = 1.0 / det;
                dpdu = (dv2 * e1 - dv1 * e2) * invDet;
            }
        }
    }
    v.ss = normalize(dpdu - v.ns * dot(v.ns, dpdu));    // computeShadingFrame
    v.mat = (int) (mf & 0x00ffffffu);
    v.emitter = __float_as_int(t2.w);
    v.type = V_SURFACE;
    v.degenerate = 0;
}

// ---------------------------------------------------------------- emitters
// DiscreteDistribution::sample + sampleReuse (include/mitsuba/core/pmf.h:134-178) on a double CDF
DR_D int cdf_sample_reuse(const double *cdf, int n, Real &v, Real &pdf) {
    // lower_bound over cdf[0..n]: first entry >= v, minus one, clamped to [0, n-1]
    const double dv = (double) v;
    int lo = 0, hi = n + 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] < dv) lo = mid + 1; else hi = mid; }
    int index = min(n - 1, max(0, lo - 1));
    while (cdf[index + 1] - cdf[index] == 0.0 && index < n) ++index;
    const double a = cdf[index], b = cdf[index + 1];
    pdf = (Real) (b - a);
    v = (Real) ((dv - a) / (b - a));
    return index;
}

struct EmitterPoint { R3 p, n; int emitter; Real pdfArea; Real emPdf; };

// Scene::sampleEmitterPosition (scene.cpp:1066-1082) -> TriMesh::samplePosition (trimesh.cpp:429-440)
// -> Triangle::sample (triangle.cpp:24-60)
DR_D void sample_emitter_point(const DevScene &sc, Real sx, Real sy, EmitterPoint &ep) {
    Real emPdf;
    const int e = cdf_sample_reuse(sc.emitterCdf, sc.nEmitters, sx, emPdf);
    const DevEmitter &em = sc.emitters[e];
    Real triPdf;
    const int tri = cdf_sample_reuse(sc.emCdf + em.cdfOffset, (int) em.nTris, sy, triPdf);
    const float4 *tp = sc.emTris + 6 * (size_t) (em.firstEmTri + tri);
    const float4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
    const R3 p0 = r3(a.x, a.y, a.z), p1 = r3(a.w, b.x, b.y), p2 = r3(b.z, b.w, c.x);
    const R3 e1 = p1 - p0, e2 = p2 - p0;
    const Real sq = safe_sqrt(1.0 - sx);
    const Real bx = 1. - sq, by = sq * sy;                 // squareToUniformTriangle
    ep.p = p0 + e1 * bx + e2 * by;
    if (__float_as_int(c.y)) {
        const float4 d = ldg4(tp + 3), f = ldg4(tp + 4), g = ldg4(tp + 5);
        const R3 n0 = r3(d.x, d.y, d.z), n1 = r3(d.w, f.x, f.y), n2 = r3(f.z, f.w, g.x);
        ep.n = normalize(n0 * (1.0 - bx - by) + n1 * bx + n2 * by);
    } else {
        ep.n = normalize(cross(e1, e2));
    }
    ep.emitter = e; ep.pdfArea = em.invArea * emPdf; ep.emPdf = emPdf;
}
DR_D R3 emitter_radiance(const DevScene &sc, int e) { const DevEmitter &em = sc.emitters[e]; return r3(em.radiance[0], em.radiance[1], em.radiance[2]); }

// ---------------------------------------------------------------- surface vertex helpers
// PathVertex::eval for ESurfaceInteraction (vertex.cpp:1026-1063), measure = EArea -> ESolidAngle
DR_D R3 surface_eval(const DevScene &sc, const Vtx &v, const Mat &m, R3 wiW, R3 woW, int mode) {
    const R3 wi = to_local(v, wiW), wo = to_local(v, woW);
    R3 r = bsdf_eval(m, wi, wo, mode, MEAS_SOLID_ANGLE);
    const Real wiDotGeoN = dot(v.ng, wiW), woDotGeoN = dot(v.ng, woW);
    if (wiDotGeoN * wi.z <= 0. || woDotGeoN * wo.z <= 0.) return r3(0.);
    if (mode == MODE_IMPORTANCE) r *= fabs((wi.z * woDotGeoN) / (wo.z * wiDotGeoN));
    if (wo.z != 0.) r = r / fabs(wo.z);
    return r;
}
// PathVertex::evalPdf for ESurfaceInteraction in area measure (vertex.cpp:1150-1204)
DR_D Real surface_pdf_area(const Vtx &v, const Mat &m, R3 fromPos, R3 toPos, R3 toNg) {
    R3 woW = toPos - v.p;
    const Real dist = length(woW);
    woW = woW / dist;
    const R3 wiW = normalize(fromPos - v.p);
    const R3 wi = to_local(v, wiW), wo = to_local(v, woW);
    Real r = bsdf_pdf(m, wi, wo, MEAS_SOLID_ANGLE);
    if (dot(v.ng, wiW) * wi.z <= 0. || dot(v.ng, woW) * wo.z <= 0.) return 0.;
    return r / (dist * dist) * absdot(woW, toNg);
}

// One BSDF sampling step of a random walk (vertex.cpp:153-271 + :334-347), shared by both subpaths.
struct WalkStep { R3 wo; R3 weightFwd; Real pdfFwd, pdfBwd; bool delta; Real eta; };
// uz: the extra number of an EUsesSampler BSDF (mat_uses_sampler), drawn by the caller right after u
DR_D bool surface_sample_next(const DevScene &sc, const Vtx &v, const Mat &m, R3 wiW, int mode, R2 u, Real uz, WalkStep &ws) {
    const R3 wi = to_local(v, wiW);
    BsdfSample bs;
    bsdf_sample(m, wi, mode, u.x, u.y, uz, sc.epsilon, bs);
    if (is_zero(bs.weight)) return false;
    const int measure = (bs.sampledType & BT_DELTA) ? MEAS_DISCRETE : MEAS_SOLID_ANGLE;
    ws.wo = to_world(v, bs.wo);
    const Real wiDotGeoN = dot(v.ng, wiW), woDotGeoN = dot(v.ng, ws.wo);
    if (wiDotGeoN * wi.z <= 0. || woDotGeoN * bs.wo.z <= 0.) return false;
    ws.pdfFwd = bs.pdf;
    ws.pdfBwd = bsdf_pdf(m, bs.wo, wi, measure);          // bRec.reverse()
    if (ws.pdfBwd <= R_RCPOVERFLOW) return false;
    ws.weightFwd = bs.weight;
    // adjoint BSDF for shading normals acts on weight[EImportance] only (vertex.cpp:252-262)
    if (mode == MODE_IMPORTANCE) ws.weightFwd *= fabs((wi.z * woDotGeoN) / (bs.wo.z * wiDotGeoN));
    ws.delta = measure == MEAS_DISCRETE;
    ws.eta = bs.eta;
    return true;
}

// ---------------------------------------------------------------- MIS (path.cpp:763-1028, no direct sampling / ENull)
struct MisArrays {
    Real pdfImp[DR_MAXK + 1], pdfRad[DR_MAXK + 1];
    Real conv[DR_MAXK + 1];       // conv[g]: len^2 / |cos_g cos_{g+1}| (geometric normals) of edge (g, g+1)
    uint32_t connectable;          // bit g
};
DR_D Real mis_weight(MisArrays &A, int s, int t, bool lightImage) {
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
    return (Real) (1.0 / weight);
}

