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

// add the finished connection (s, t) with value `value` (geometric term included, MIS weight not yet) to the splat list
DR_D void bd_finish_connection(const Machine &M, int lane, Core &c, BdAcc &acc, R3 value) {
    const int s = c.s, t = c.t;
    value *= bd_mis(M, lane, s, t, acc.pdfs);
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

// Evaluate connections in the reference's order starting AFTER (c.s, c.t) until one needs a shadow ray.
// Returns true when a ray was emitted; false when every pair has been handled (the path is complete).
DR_D bool bd_next_connection(const Machine &M, int lane, Core &c, BdAcc &acc, RayF &ray) {
    const DevScene &sc = M.sc;
    const PathCfg &pc = M.pc;
    int s = c.s, t = c.t;
    for (;;) {
        // ---- advance (s, t): t runs from maxT(s) down to minT(s), s from ns down to 0 (pathsampler.cpp:369-377)
        --t;
        for (;;) {
            if (s < 0) return false;
            const int minT = max(max(2 - s, pc.lightImage ? 0 : 2), 1);   // t = 0 needs a sensor with a shape: never for a pinhole
            if (t >= minT) break;
            --s;
            if (s < 0) return false;
            t = min(acc.nt, pc.maxDepth + 1 - s);
        }
        c.s = (uint8_t) s; c.t = (uint8_t) t;
        const int depth = s + t - 1;
        if (pc.excludeDirect && depth <= 2) continue;
        EndPoint et;
        bd_load_end(M, lane, BD_S, t, et);
        BExtra xt;
        rec_load(xt, bxp(M, lane, BD_S, t));
        Real *pdfs = acc.pdfs;
        if (s == 0) {                                        // pure sensor path: vt must be on an emitter (:213-224 / :391-398)
            const Vtx &vt = et.v;
            if (vt.type != V_SURFACE || vt.emitter < 0) continue;
            R3 wo = et.predP - vt.p;
            const Real dist = length(wo);
            wo = wo / dist;
            const Real dp = dot(wo, vt.ns);
            if (!(dp > 0.)) continue;
            const R3 value = xt.prefix * emitter_radiance(sc, vt.emitter);
            if (is_zero(value)) continue;
            const DevEmitter &em = sc.emitters[vt.emitter];
            pdfs[0] = em.invArea * em.pdfDiscrete;                                   // pdfImp[1]
            pdfs[1] = 0.;
            pdfs[2] = 1.0;                                                            // pdfRad[0]: vt->evalPdf(vtPred, vs, ERadiance) towards the supernode
            pdfs[3] = R_INV_PI * dp / (dist * dist) * absdot(wo, et.predNg);         // pdfImp[2]
            bd_finish_connection(M, lane, c, acc, value);
            continue;
        }
        EndPoint es;
        bd_load_end(M, lane, BD_E, s, es);
        BExtra xs;
        rec_load(xs, bxp(M, lane, BD_E, s));
        const Vtx &vs = es.v, &vt = et.v;
        if (vs.degenerate || vt.degenerate) continue;
        R3 d = vs.p - vt.p;                                  // from vt towards vs
        const Real len = length(d);
        if (len == 0.) continue;
        d = d / len;
        R3 fs, ft;
        Mat ms, mt;
        if (s == 1) {
            const Real dp = dot(-d, vs.ns);
            fs = r3(dp > 0. ? R_INV_PI : 0.);
        } else {
            ms = load_material(sc, vs.mat);
            fs = surface_eval(sc, vs, ms, normalize(es.predP - vs.p), -d, MODE_IMPORTANCE);
        }
        if (t == 1) {
            const Real imp = cam_importance(sc.cam, cam_inv_dir(sc.cam, d));
            const Real dp = absdot(vt.ns, d);
            ft = r3(dp != 0. ? imp / dp : imp);
        } else {
            mt = load_material(sc, vt.mat);
            ft = surface_eval(sc, vt, mt, normalize(et.predP - vt.p), d, MODE_RADIANCE);
        }
        R3 value = xs.prefix * xt.prefix * fs * ft;
        if (is_zero(value)) continue;
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
            if (!cam_sample_position(sc.cam, vs.p - vt.p, sp)) continue;           // :298-303 / :506-508
            c.spos = make_float2((float) sp.x, (float) sp.y);
        } else {
            pdfs[2] = surface_pdf_area(vt, mt, et.predP, vs.p, vs.ng);
            pdfs[3] = surface_pdf_area(vt, mt, vs.p, et.predP, et.predNg);
        }
        c.weight = value;
        c.pstate = PS_BD_SHADOW;
        emit_ray(M, lane, c, vt.p, d, sc.epsilon, len * (1. - sc.shadowEpsilon), ray);
        return true;
    }
}

} // namespace

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
                    const Mat nm = load_material(sc, nv.mat);
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
                    if (bd_next_connection(M, lane, c, acc, ray)) dest = Q_RAYS + (M.parity ^ 1);
                }
            }
        } else if (c.pstate == PS_BD_SHADOW) {               // the shadow ray of connection (c.s, c.t) arrived
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

void launch_bdpt(const Machine &M, const LaunchCfg &lc) {
    const unsigned g = stage_grid(lc.nLanes, 128);
    k_bdpt<<<g, 128, 0, lc.stream>>>(M);
}
