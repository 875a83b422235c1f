// k_bdpt.cu -- technique=bdpt: bidirectional path tracing with multiple importance sampling as a resumable
// wavefront stage (one ray in flight per lane, like every other stage of machine.cuh).
//
// PathSampler::sampleSplats, EBidirectional branch (src/libbidir/pathsampler.cpp:321-527) with m_sampleDirect = false
// (the reference overflows its direct sampler otherwise):
//   1. two random walks with Russian roulette from rrDepth (Path::randomWalk path.cpp:500-535, PathVertex::sampleNext
//      vertex.cpp:37-350): emitter subpath of <= maxDepth steps, sensor subpath of <= maxDepth + 1 steps.  Both are kept
//      in the lane's subpath arrays (bv / bx): every vertex can be a connection end point;
//   2. every pair (s, t), s from the longest emitter prefix down to 0, t from the longest allowed sensor prefix down:
//      PathVertex::eval at both ends, shadow ray (pathConnectAndCollapse edge.cpp:558-690), generalized geometric term
//      (edge.cpp:221-271), Path::miWeight (path.cpp:763-1028);  t >= 2 accumulates into splat 0 (the pixel of the sensor
//      subpath), t = 1 appends a light-image splat at the pixel the emitter-side vertex projects to.
// The result (BdAcc + the lane's in-flight splat list) goes to k_chain like that of any other technique.
#include "bdpt.cuh"

namespace {

// ---- connections ---------------------------------------------------------------------------------------------------
struct EndPoint { Vtx v; R3 predP, predNg; };

DR_D void bd_load_end(const Machine &M, int lane, int side, int idx, EndPoint &e) {
    rec_load(e.v, bvp(M, lane, side, idx));
    e.predP = r3(0.); e.predNg = r3(0.);
    if (idx >= 2) {
        Vtx p;
        rec_load(p, bvp(M, lane, side, idx - 1));
        e.predP = p.p; e.predNg = p.ng;
    }
}

// Path::miWeight for strategy (s, t) from the cached per-vertex densities and the four recomputed ones
DR_D Real bd_mis(const Machine &M, int lane, int s, int t, const Real pdfs[4]) {
    const int k = s + t + 1;
    MisArrays A;
    A.connectable = 0;
#pragma unroll 1
    for (int i = 0; i <= k; ++i) {
        const int side = i <= s ? BD_E : BD_S, idx = i <= s ? i : k - i;
        BExtra x;
        rec_load(x, bxp(M, lane, side, idx));
        Vtx v;
        bool connectable;
        if (idx == 0) connectable = side == BD_E;                 // emitter supernode: area measure | sensor supernode: discrete
        else {
            rec_load(v, bvp(M, lane, side, idx));
            connectable = !v.degenerate && !x.discrete;
        }
        if (i == s || i == s + 1) connectable = true;             // measure forced to EArea; non-degenerate was checked
        if (connectable) A.connectable |= 1u << i;
        Real imp, rad, conv = 0.;
        if (i <= s) {
            imp = i == 0 ? 1.0 : x.fwd;
            rad = x.bwd;
            conv = x.conv;                                        // edge (i, i+1) of the emitter subpath
        } else {
            rad = i == k ? 1.0 : x.fwd;
            imp = x.bwd;
            // edge (i, i+1) is edge (idx-1, idx) of the sensor subpath
            if (idx >= 1) { BExtra y; rec_load(y, bxp(M, lane, BD_S, idx - 1)); conv = y.conv; }
        }
        A.pdfImp[i] = imp; A.pdfRad[i] = rad; A.conv[i] = conv;
    }
    A.pdfImp[s + 1] = pdfs[0];
    if (s >= 1) A.pdfRad[s - 1] = pdfs[1];
    A.pdfRad[s] = pdfs[2];
    if (s + 2 <= k) A.pdfImp[s + 2] = pdfs[3];
    for (int i = k + 1; i <= DR_MAXK; ++i) { A.pdfImp[i] = 0.; A.pdfRad[i] = 0.; A.conv[i] = 0.; }
    return mis_weight(A, s, t, M.pc.lightImage != 0);
}

// The per-vertex part of the MIS arrays of a whole path, loaded ONCE (batched connections): what bd_mis re-reads for every pair
struct MisBase {
    Real eFwd[BD_MAXV], eBwd[BD_MAXV], eConv[BD_MAXV];    // emitter subpath vertex v: fwd, bwd, conv of edge (v, v + 1)
    Real sFwd[BD_MAXV], sBwd[BD_MAXV], sConv[BD_MAXV];    // sensor subpath
    uint32_t eOk, sOk;                                     // bit v: vertex v is connectable (non-degenerate, not sampled from a delta lobe)
};
DR_D void bd_mis_base(const Machine &M, int lane, int ns, int nt, MisBase &B) {
    B.eOk = 1u; B.sOk = 0u;                                // emitter supernode: area measure | sensor supernode: discrete
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        const int last = side == BD_E ? ns : nt;
#pragma unroll 1
        for (int v = 0; v <= last && v < BD_MAXV; ++v) {
            BExtra x;
            rec_load(x, bxp(M, lane, side, v));
            if (side == BD_E) { B.eFwd[v] = x.fwd; B.eBwd[v] = x.bwd; B.eConv[v] = x.conv; if (v >= 1 && !x.pad[0] && !x.discrete) B.eOk |= 1u << v; }
            else { B.sFwd[v] = x.fwd; B.sBwd[v] = x.bwd; B.sConv[v] = x.conv; if (v >= 1 && !x.pad[0] && !x.discrete) B.sOk |= 1u << v; }
        }
    }
}
DR_D Real bd_mis_cached(const Machine &M, const MisBase &B, int s, int t, const Real pdfs[4]) {
    const int k = s + t + 1;
    MisArrays A;
    A.connectable = 0;
#pragma unroll 1
    for (int i = 0; i <= k; ++i) {
        const bool eSide = i <= s;
        const int idx = eSide ? i : k - i;
        bool connectable = ((eSide ? B.eOk : B.sOk) >> idx) & 1u;
        if (i == s || i == s + 1) connectable = true;             // measure forced to EArea; non-degenerate was checked
        if (connectable) A.connectable |= 1u << i;
        Real imp, rad, conv = 0.;
        if (eSide) { imp = i == 0 ? 1.0 : B.eFwd[idx]; rad = B.eBwd[idx]; conv = B.eConv[idx]; }
        else { rad = i == k ? 1.0 : B.sFwd[idx]; imp = B.sBwd[idx]; if (idx >= 1) conv = B.sConv[idx - 1]; }
        A.pdfImp[i] = imp; A.pdfRad[i] = rad; A.conv[i] = conv;
    }
    A.pdfImp[s + 1] = pdfs[0];
    if (s >= 1) A.pdfRad[s - 1] = pdfs[1];
    A.pdfRad[s] = pdfs[2];
    if (s + 2 <= k) A.pdfImp[s + 2] = pdfs[3];
    for (int i = k + 1; i <= DR_MAXK; ++i) { A.pdfImp[i] = 0.; A.pdfRad[i] = 0.; A.conv[i] = 0.; }
    return mis_weight(A, s, t, M.pc.lightImage != 0);
}

// add the finished connection (s, t) with value `value` (geometric term included, MIS weight not yet) to the splat list
DR_D void bd_finish_connection(const Machine &M, int lane, Core &c, BdAcc &acc, R3 value, const MisBase *base = nullptr) {
    const int s = c.s, t = c.t;
    value *= base ? bd_mis_cached(M, *base, s, t, acc.pdfs) : bd_mis(M, lane, s, t, acc.pdfs);
    if (t < 2) {
        if (acc.nl < BD_MAXS - 1) {
            float4 *sp = M.lm.bsplat + (((size_t) lane * 4 + 3) * BD_MAXS + acc.nl) * 2;
            sp[0] = make_float4(c.spos.x, c.spos.y, 0.f, 0.f);
            sp[1] = make_float4((float) value.x, (float) value.y, (float) value.z, 0.f);
            ++acc.nl;
        }
    } else acc.val0 += value;
    acc.lum += luminance(value);
}

// un-rounded ray of a connection: origin, direction, [tmin, tmax]
struct RayD { R3 o, d; Real tmin, tmax; };

// advance (s, t) to the next pair in the reference's order: t runs from maxT(s) down to minT(s), s from ns down to 0
// (pathsampler.cpp:369-377).  False: no pair is left.
DR_D bool bd_advance(const PathCfg &pc, const BdAcc &acc, int &s, int &t) {
    --t;
    for (;;) {
        if (s < 0) return false;
        const int minT = max(max(2 - s, pc.lightImage ? 0 : 2), 1);   // t = 0 needs a sensor with a shape: never for a pinhole
        if (t >= minT) return true;
        --s;
        if (s < 0) return false;
        t = min(acc.nt, pc.maxDepth + 1 - s);
    }
}

// Evaluate connection (s, t): 0 = contributes nothing, 1 = complete without a shadow ray (s = 0: the sensor subpath hit an
// emitter), 2 = needs the shadow ray `ray`.  value (geometric term included, MIS weight not yet), the four densities next to
// the connection and, for t = 1, the pixel of the light-image splat are returned through `value`, `pdfs`, `spos`.
// (es / xs: emitter-side end point s and its record, unused for s = 0; et / xt: sensor-side end point t)
DR_D int bd_eval_connection(const Machine &M, int s, int t, const EndPoint &es, const BExtra &xs, const EndPoint &et, const BExtra &xt,
                            R3 &value, Real pdfs[4], float2 &spos, RayD &ray) {
    const DevScene &sc = M.sc;
    const PathCfg &pc = M.pc;
    const int depth = s + t - 1;
    if (pc.excludeDirect && depth <= 2) return 0;
    if (s == 0) {                                        // pure sensor path: vt must be on an emitter (:213-224 / :391-398)
        const Vtx &vt = et.v;
        if (vt.type != V_SURFACE || vt.emitter < 0) return 0;
        R3 wo = et.predP - vt.p;
        const Real dist = length(wo);
        wo = wo / dist;
        const Real dp = dot(wo, vt.ns);
        if (!(dp > 0.)) return 0;
        value = xt.prefix * emitter_radiance(sc, vt.emitter);
        if (is_zero(value)) return 0;
        const DevEmitter &em = sc.emitters[vt.emitter];
        pdfs[0] = em.invArea * em.pdfDiscrete;                                   // pdfImp[1]
        pdfs[1] = 0.;
        pdfs[2] = 1.0;                                                            // pdfRad[0]: vt->evalPdf(vtPred, vs, ERadiance) towards the supernode
        pdfs[3] = R_INV_PI * dp / (dist * dist) * absdot(wo, et.predNg);         // pdfImp[2]
        return 1;
    }
    const Vtx &vs = es.v, &vt = et.v;
    if (vs.degenerate || vt.degenerate) return 0;
    R3 d = vs.p - vt.p;                                  // from vt towards vs
    const Real len = length(d);
    if (len == 0.) return 0;
    d = d / len;
    R3 fs, ft;
    Mat ms, mt;
    if (s == 1) {
        const Real dp = dot(-d, vs.ns);
        fs = r3(dp > 0. ? R_INV_PI : 0.);
    } else {
        ms = load_material(sc, vs.mat, vs.uv);
        fs = surface_eval(sc, vs, ms, normalize(es.predP - vs.p), -d, MODE_IMPORTANCE);
    }
    if (t == 1) {
        const Real imp = cam_importance(sc.cam, cam_inv_dir(sc.cam, d));
        const Real dp = absdot(vt.ns, d);
        ft = r3(dp != 0. ? imp / dp : imp);
    } else {
        mt = load_material(sc, vt.mat, vt.uv);
        ft = surface_eval(sc, vt, mt, normalize(et.predP - vt.p), d, MODE_RADIANCE);
    }
    value = xs.prefix * xt.prefix * fs * ft;
    if (is_zero(value)) return 0;
    value *= absdot(vs.ns, d) * absdot(vt.ns, d) / (len * len);
    if (s == 1) {
        const Real dp = dot(-d, vs.ns);
        pdfs[0] = R_INV_PI * fmax(dp, 0.) / (len * len) * absdot(d, vt.ng);
        pdfs[1] = 1.0;
    } else {
        pdfs[0] = surface_pdf_area(vs, ms, es.predP, vt.p, vt.ng);
        pdfs[1] = surface_pdf_area(vs, ms, vt.p, es.predP, es.predNg);
    }
    if (t == 1) {
        pdfs[2] = cam_importance(sc.cam, cam_inv_dir(sc.cam, d)) / (len * len) * absdot(d, vs.ng);
        pdfs[3] = 1.0;
        R2 sp = r2(0., 0.);
        if (!cam_sample_position(sc.cam, vs.p - vt.p, sp)) return 0;            // :298-303 / :506-508
        spos = make_float2((float) sp.x, (float) sp.y);
    } else {
        pdfs[2] = surface_pdf_area(vt, mt, et.predP, vs.p, vs.ng);
        pdfs[3] = surface_pdf_area(vt, mt, vs.p, et.predP, et.predNg);
    }
    ray.o = vt.p; ray.d = d; ray.tmin = sc.epsilon; ray.tmax = len * (1. - sc.shadowEpsilon);
    return 2;
}

// One pair per round: evaluate connections in the reference's order starting AFTER (c.s, c.t) until one needs a shadow ray.
// Returns true when a ray was emitted; false when every pair has been handled (the path is complete).
DR_D bool bd_next_connection(const Machine &M, int lane, Core &c, BdAcc &acc, RayF &ray) {
    int s = c.s, t = c.t;
    for (;;) {
        if (!bd_advance(M.pc, acc, s, t)) return false;
        c.s = (uint8_t) s; c.t = (uint8_t) t;
        R3 value; RayD rd; float2 spos = c.spos;
        EndPoint es, et;
        BExtra xs, xt;
        bd_load_end(M, lane, BD_S, t, et);
        rec_load(xt, bxp(M, lane, BD_S, t));
        if (s >= 1) { bd_load_end(M, lane, BD_E, s, es); rec_load(xs, bxp(M, lane, BD_E, s)); }
        const int kind = bd_eval_connection(M, s, t, es, xs, et, xt, value, acc.pdfs, spos, rd);
        if (kind == 0) continue;
        c.spos = spos;
        if (kind == 1) { bd_finish_connection(M, lane, c, acc, value); continue; }
        c.weight = value;
        c.pstate = PS_BD_SHADOW;
        emit_ray(M, lane, c, rd.o, rd.d, rd.tmin, rd.tmax, ray);
        return true;
    }
}

// ---- batched connections, one WARP per path (k_bd_connect): pair p of the reference's order (s from ns down to 0, t from maxT(s)
// down to minT(s), pathsampler.cpp:369-377) is record p of the lane and bit p of its visibility mask.
DR_D int bd_pair_count(const PathCfg &pc, const BdAcc &acc) {
    int total = 0;
    for (int s = acc.ns; s >= 0; --s) {
        const int minT = max(max(2 - s, pc.lightImage ? 0 : 2), 1), maxT = min(acc.nt, pc.maxDepth + 1 - s);
        total += max(0, maxT - minT + 1);
    }
    return total;
}
DR_D void bd_pair(const PathCfg &pc, const BdAcc &acc, int p, int &s, int &t) {
    for (s = acc.ns; s >= 0; --s) {
        const int minT = max(max(2 - s, pc.lightImage ? 0 : 2), 1), maxT = min(acc.nt, pc.maxDepth + 1 - s);
        const int n = max(0, maxT - minT + 1);
        if (p < n) { t = maxT - p; return; }
        p -= n;
    }
    s = 0; t = 0;
}
// Evaluate every pair (thread `self` of the warp takes pairs self, self + 32, ...), keep the records, send the shadow rays of the
// pairs that can contribute through Q_BDS.  Returns the number of rays in flight (warp-uniform).
DR_D int bd_connect_all(const Machine &M, int lane, Core &c, const BdAcc &acc, unsigned self) {
    const int stride = M.lm.bdStride;
    BdConn *recs = M.lm.bconn + (size_t) lane * stride;
    const int next = M.parity ^ 1;
    const int total = min(bd_pair_count(M.pc, acc), stride);
    int nRays = 0;
    for (int p = (int) self; p < total; p += 32) {
        int s, t;
        bd_pair(M.pc, acc, p, s, t);
        EndPoint es, et;
        BExtra xs, xt;
        bd_load_end(M, lane, BD_S, t, et);
        rec_load(xt, bxp(M, lane, BD_S, t));
        if (s >= 1) { bd_load_end(M, lane, BD_E, s, es); rec_load(xs, bxp(M, lane, BD_E, s)); }
        BdConn rec;
        memset(&rec, 0, sizeof(rec));
        RayD rd;
        const int kind = bd_eval_connection(M, s, t, es, xs, et, xt, rec.value, rec.pdfs, rec.spos, rd);
        rec.s = (uint8_t) s; rec.t = (uint8_t) t; rec.needsRay = kind == 2; rec.pad[0] = kind != 0;      // pad[0]: the pair can contribute
        rec_store(recs + p, rec);
        if (kind == 2) {
            Real tmin = rd.tmin;
            if (tmin == (Real) M.sc.epsilon) tmin *= fmax(fmax(fmax(fabs(rd.o.x), fabs(rd.o.y)), fabs(rd.o.z)), (Real) M.sc.epsilon);   // emit_ray
            double2 *dst = reinterpret_cast<double2 *>(M.lm.brayd + 8 * ((size_t) lane * stride + p));
            dst[0] = make_double2(rd.o.x, rd.o.y); dst[1] = make_double2(rd.o.z, rd.d.x); dst[2] = make_double2(rd.d.y, rd.d.z); dst[3] = make_double2(tmin, rd.tmax);
            // warp-aggregated append to the batch queue of the next round
            const unsigned act = __activemask();
            const int leader = __ffs(act) - 1;
            uint32_t base = 0;
            if ((int) self == leader) base = atomicAdd(&M.q.count[Q_BDS + next], (uint32_t) __popc(act));
            base = __shfl_sync(act, base, leader);
            const size_t slot = (size_t) next * M.q.bn + base + __popc(act & ((1u << self) - 1u));
            M.q.bitems[slot] = (uint32_t) lane * BD_MAXC + (uint32_t) p;
            M.q.brays[2 * slot] = make_float4((float) rd.o.x, (float) rd.o.y, (float) rd.o.z, (float) tmin);
            M.q.brays[2 * slot + 1] = make_float4((float) rd.d.x, (float) rd.d.y, (float) rd.d.z, (float) rd.tmax);
            ++nRays;
        }
    }
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) nRays += __shfl_xor_sync(0xffffffffu, nRays, o);
    if (self == 0) {
        M.lm.bcount[lane] = (uint32_t) total;
        M.lm.bvis[lane] = 0ull;
        M.lm.bpend[lane] = (uint32_t) nRays;
    }
    c.nrays += nRays;
    return nRays;
}
// ... and, once all shadow rays are back: the MIS weights of the visible pairs in parallel, then thread 0 accumulates them in the
// reference's order (the sums of the splat list do not depend on how the work was spread)
DR_D void bd_finish_all(const Machine &M, int lane, Core &c, BdAcc &acc, unsigned self) {
    const int stride = M.lm.bdStride;
    BdConn *recs = M.lm.bconn + (size_t) lane * stride;
    const int n = (int) M.lm.bcount[lane];
    const unsigned long long vis = M.lm.bvis[lane];
    if (n > (int) self) {
        MisBase base;
        bd_mis_base(M, lane, acc.ns, acc.nt, base);
        for (int p = (int) self; p < n; p += 32) {
            BdConn rec;
            rec_load(rec, recs + p);
            const bool ok = rec.pad[0] && (!rec.needsRay || ((vis >> p) & 1ull));
            if (ok) rec.value *= bd_mis_cached(M, base, rec.s, rec.t, rec.pdfs);
            rec.pad[0] = ok;
            rec_store(recs + p, rec);
        }
    }
    __syncwarp();
    if (self == 0)
        for (int p = 0; p < n; ++p) {
            BdConn rec;
            rec_load(rec, recs + p);
            if (!rec.pad[0]) continue;
            if (rec.t < 2) {
                if (acc.nl < BD_MAXS - 1) {
                    float4 *sp = M.lm.bsplat + (((size_t) lane * 4 + 3) * BD_MAXS + acc.nl) * 2;
                    sp[0] = make_float4(rec.spos.x, rec.spos.y, 0.f, 0.f);
                    sp[1] = make_float4((float) rec.value.x, (float) rec.value.y, (float) rec.value.z, 0.f);
                    ++acc.nl;
                }
            } else acc.val0 += rec.value;
            acc.lum += luminance(rec.value);
        }
}

} // namespace

// BATCH: the connections are not made here -- a lane whose two walks are over goes to the connection queue (k_bd_connect), so
// that the warps of this kernel only ever run walk steps (a connection phase is ~50x the work of a walk step: one such lane per
// warp would make every warp pay for it in every round).
template <bool BATCH>
__global__ void __launch_bounds__(128)
k_bdpt(const __grid_constant__ Machine M) {
    const DevScene &sc = M.sc;
    const PathCfg &pc = M.pc;
    const uint32_t cnt = M.q.count[Q_PT];
    const uint32_t *items = M.q.items + (size_t) Q_PT * M.q.n;
    for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < cnt; qi += gridDim.x * blockDim.x) {
        const int lane = (int) items[qi];
        Core c;
        rec_load(c, M.lm.core + lane);
        BdAcc acc;
        rec_load(acc, M.lm.bacc + lane);
        Hit hit;
        hit.t = hit.u = hit.v = 0.f; hit.tri = __ldcs(M.q.aux + (size_t) Q_PT * M.q.n + qi);
        UReader rd;
        reader_open(M, c, lane, rd);
        RayF ray;
        int dest = -1;                                       // ray queue, or -1 while undecided
        bool walkEnded = false;                              // the subpath of c.pstate ended: start the next phase
        if (c.pstate == PS_BD_EHIT || c.pstate == PS_BD_SHIT) {
            const int side = c.pstate == PS_BD_EHIT ? BD_E : BD_S;
            const int maxV = side == BD_E ? pc.maxDepth : pc.maxDepth + 1;       // emitterDepth / sensorDepth (pathsampler.cpp:53-71)
            const int j = c.j;
            if (hit.tri < 0) walkEnded = true;
            else {
                Vtx v;
                rec_load(v, bvp(M, lane, side, j));
                Vtx nv; Real tHit;
                fill_vertex(sc, hit, v.p, c.d, nv, tHit);
                if (tHit == 0.) walkEnded = true;
                else {
                    const Mat nm = load_material(sc, nv.mat, nv.uv);
                    nv.degenerate = !(mat_has_smooth(nm.type) || nv.emitter >= 0);
                    // solid angle -> area (vertex.cpp:334-347)
                    const Real cosNext = absdot(c.d, nv.ng);
                    Real pdfFwd = c.pdfFwd, pdfBwd = c.pdfBwd;
                    R3 predP = r3(0.);
                    if (j >= 2) { Vtx pv; rec_load(pv, bvp(M, lane, side, j - 1)); predP = pv.p;
                        if (!(c.flags & F_DELTA)) { R3 pd = v.p - pv.p; const Real plen = length(pd); pd = pd / plen; pdfBwd = pdfBwd / (plen * plen) * absdot(pd, pv.ng); } }
                    if (!(c.flags & F_DELTA)) pdfFwd = pdfFwd / (tHit * tHit) * cosNext;
                    BExtra xj, xn, xp;
                    rec_load(xj, bxp(M, lane, side, j));
                    rec_load(xn, bxp(M, lane, side, j + 1));
                    xj.conv = tHit * tHit / fabs(absdot(c.d, v.ng) * cosNext);
                    xj.discrete = (c.flags & F_DELTA) ? 1u : 0u;
                    xn.fwd = pdfFwd;
                    xn.pad[0] = nv.degenerate ? 1u : 0u;          // (read by bd_mis_base instead of the 128-byte vertex)
                    rec_store(bxp(M, lane, side, j), xj);
                    rec_store(bxp(M, lane, side, j + 1), xn);
                    rec_load(xp, bxp(M, lane, side, j - 1));
                    xp.bwd = pdfBwd;
                    rec_store(bxp(M, lane, side, j - 1), xp);
                    rec_store(bvp(M, lane, side, j + 1), nv);
                    if (side == BD_E) acc.ns = j + 1; else acc.nt = j + 1;
                    if (side == BD_S && j == 1) {            // pixel of splat 0 (pathsampler.cpp:355-362)
                        R2 sp = r2(0., 0.);
                        cam_sample_position(sc.cam, nv.p - cam_pos(sc.cam), sp);
                        acc.pos0 = make_float2((float) sp.x, (float) sp.y);
                        acc.has0 = 1;
                    }
                    // next step of the walk (step index j + 1), unless the subpath has its full length
                    if (j + 1 < maxV && bd_sample_surface(M, lane, c, rd, side, j + 1, nv, v.p, ray)) dest = Q_RAYC + (M.parity ^ 1);
                    else walkEnded = true;
                }
            }
            if (walkEnded) {
                if (side == BD_E) {
                    if (bd_sensor_start(M, lane, c, rd, acc, ray)) dest = Q_RAYC + (M.parity ^ 1);
                    else { c.s = (uint8_t) (acc.ns + 1); c.t = 0; c.pstate = PS_BD_SHADOW; }   // (cannot connect anything: falls through to the loop)
                }
                if (dest < 0) {                              // both walks are done: first connection
                    c.s = (uint8_t) acc.ns; c.t = (uint8_t) (min(acc.nt, pc.maxDepth + 1 - acc.ns) + 1);
                    if (BATCH) { c.pstate = PS_BD_BATCH; dest = Q_CONNECT; }
                    else if (bd_next_connection(M, lane, c, acc, ray)) dest = Q_RAYS + (M.parity ^ 1);
                }
            }
        } else if (!BATCH && c.pstate == PS_BD_SHADOW) {     // the shadow ray of connection (c.s, c.t) arrived
            if (hit.tri < 0) bd_finish_connection(M, lane, c, acc, c.weight);
            if (bd_next_connection(M, lane, c, acc, ray)) dest = Q_RAYS + (M.parity ^ 1);
        }
        if (dest < 0) { c.pstate = PS_BD_DONE; dest = Q_CHAIN + M.parity; }
        reader_close(rd, c);
        rec_store(M.lm.bacc + lane, acc);
        rec_store(M.lm.core + lane, c);
        q_push_ray(M.q, dest, (uint32_t) lane, ray);
    }
}

// Batched connections, one WARP per path: the warps [0, n0) take the lanes whose walks just ended (Q_CONNECT): evaluate all pairs, send
// the shadow rays; the warps [n0, n0 + n1) the lanes whose shadow rays are all back (first walk queue, unused by BDPT; filled by
// k_bd_shadow): MIS-weight and accumulate.  Finished paths go to the chain queue of this round.  (A connection phase is ~50x the work of a
// walk step and only ~1 / 18 of the lanes are in it: one thread per lane left ~150 warps per launch with 44 serial connections each.)
__global__ void __launch_bounds__(128)
k_bd_connect(const __grid_constant__ Machine M) {
    const uint32_t n0 = M.q.count[Q_CONNECT], n1 = M.q.count[Q_WALK];
    const uint32_t nWarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned self = threadIdx.x & 31u;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n0 + n1; w += nWarps) {
        const bool connect = w < n0;
        const int lane = (int) M.q.items[(size_t) (connect ? Q_CONNECT : Q_WALK) * M.q.n + (connect ? w : w - n0)];
        Core c;
        rec_load(c, M.lm.core + lane);
        BdAcc acc;
        rec_load(acc, M.lm.bacc + lane);
        bool parked = false;
        if (connect && bd_connect_all(M, lane, c, acc, self) > 0) parked = true;      // (the last of its shadow rays brings the lane back)
        if (!parked) bd_finish_all(M, lane, c, acc, self);
        if (self == 0) {
            if (!parked) c.pstate = PS_BD_DONE;
            rec_store(M.lm.bacc + lane, acc);
            rec_store(M.lm.core + lane, c);
            if (!parked) q_push(M.q, Q_CHAIN + M.parity, (uint32_t) lane);
        }
        __syncwarp();
    }
}

// The batched shadow rays of the round (Q_BDS): any-hit traversal, one thread per ray; the visibility bit goes to the lane's mask
// and the LAST ray of a lane to come back queues the lane for k_bdpt (which runs next, in the same round).
__global__ void __launch_bounds__(128)
k_bd_shadow(const __grid_constant__ Machine M) {
    const uint32_t cnt = M.q.count[Q_BDS + M.parity];
    const uint32_t *items = M.q.bitems + (size_t) M.parity * M.q.bn;
    const float4 *rays = M.q.brays + 2 * (size_t) M.parity * M.q.bn;
    if (blockIdx.x == 0 && threadIdx.x == 0 && cnt) atomicAdd(&M.counters[ST_RAYS], (unsigned long long) cnt);
    for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < cnt; qi += gridDim.x * blockDim.x) {
        const uint32_t item = __ldcs(items + qi);
        const int lane = (int) (item / BD_MAXC), ci = (int) (item % BD_MAXC);
        const float4 a = __ldcs(rays + 2 * (size_t) qi), b = __ldcs(rays + 2 * (size_t) qi + 1);
        Hit h;
        const bool occluded = traverse<true>(M.sc, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, b.w, M.lm.brayd + 8 * ((size_t) lane * M.lm.bdStride + ci), h);
        if (!occluded) atomicOr(M.lm.bvis + lane, 1ull << ci);
        __threadfence();
        if (atomicSub(M.lm.bpend + lane, 1u) == 1u) q_push(M.q, Q_WALK, (uint32_t) lane);
    }
}

void launch_bdpt(const Machine &M, const LaunchCfg &lc) {
    const unsigned g = stage_grid(lc.nLanes, 128);
    if (M.pc.bdBatch) {
        k_bd_shadow<<<stage_grid(lc.nLanes * 8, 128), 128, 0, lc.stream>>>(M);
        k_bdpt<true><<<g, 128, 0, lc.stream>>>(M);
        k_bd_connect<<<stage_grid(lc.nLanes * 4, 128), 128, 0, lc.stream>>>(M);
    } else k_bdpt<false><<<g, 128, 0, lc.stream>>>(M);
}
