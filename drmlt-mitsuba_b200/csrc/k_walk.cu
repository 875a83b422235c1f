// k_walk.cu -- MMLT wavefront stages that turn a closest hit into the next path vertex (k_walk) and that
// connect the two subpath ends (k_connect).  See machine.cuh for the machine, path.cuh for the building blocks.
//
// k_walk is instantiated once per BSDF model: the traversal kernel routes every hit to the walk queue of
// the material class it found, so each instance runs one BSDF's sampling code on full warps.
// Behavioural parity: PathVertex::sampleNext (src/libbidir/vertex.cpp:153-350), PathEdge::sampleNext
// (edge.cpp:27-84), the MMLT branch of PathSampler::sampleSplats (pathsampler.cpp:139-295),
// PathVertex::eval / evalPdf (vertex.cpp:958-1205), PathEdge::evalCached (edge.cpp:221-271).
#include "machine.cuh"

template <int BSDF>
DR_D void walk_lane(const Machine &M, uint32_t qi) {
    const DevScene &sc = M.sc;
    {
        const int lane = (int) M.q.items[(size_t) (Q_WALK + BSDF) * M.q.n + qi];
        Core c;
        rec_load(c, M.lm.core + lane);
        const bool emitterSide = c.pstate == PS_EMITTER_HIT;
        const int side = emitterSide ? SIDE_E : SIDE_S;
        const int k = c.depth + 2;                           // s + t + 1
        Hit hit;
        hit.t = hit.u = hit.v = 0.f; hit.tri = __ldcs(M.q.aux + (size_t) (Q_WALK + BSDF) * M.q.n + qi);
        int dest = Q_CHAIN + M.parity;                       // default: the path ends here (empty result)
        RayF ray;
        do {
            PredRec v, vp;                                   // vertex j (origin of the ray) and its predecessor: position + geometric normal
            int j = c.j;
            rec_load(v, geo_slot(M, lane, side, j));
            if (j >= 2) rec_load(vp, geo_slot(M, lane, side, j - 1));
            const R3 d = c.d;
            Vtx nv; Real tHit;
            fill_vertex(sc, hit, v.p, d, nv, tHit);
            if (tHit == 0.) { c.pstate = PS_EMPTY; break; }
            Mat nm = load_material(sc, nv.mat, nv.uv);
            nm.type = BSDF;                                  // compile-time constant for the BSDF switch
            nv.degenerate = !(mat_has_smooth(BSDF) || nv.emitter >= 0);
            // solid angle -> area (vertex.cpp:334-347); delta interactions keep their discrete pdfs
            const Real cosNext = absdot(d, nv.ng);
            Real pdfFwd = c.pdfFwd, pdfBwd = c.pdfBwd;
            if (!(c.flags & F_DELTA)) {
                pdfFwd = pdfFwd / (tHit * tHit) * cosNext;
                if (j >= 2) {
                    R3 pd = v.p - vp.p;
                    const Real plen = length(pd);
                    pd = pd / plen;
                    pdfBwd = pdfBwd / (plen * plen) * absdot(pd, vp.ng);
                }
            }
            misrec_store(M, lane, side, j, pdfFwd, pdfBwd, tHit * tHit / fabs(absdot(d, v.ng) * cosNext));
            if (!emitterSide && j == 1) {                    // pixel of the path (pathsampler.cpp:309-312)
                R2 sp = r2(0., 0.);
                cam_sample_position(sc.cam, nv.p - cam_pos(sc.cam), sp);
                c.spos = make_float2((float) sp.x, (float) sp.y);
            }
            ++j;
            c.j = (uint8_t) j;
            PredRec ng_;
            ng_.p = nv.p; ng_.ng = nv.ng; ng_.pad[0] = ng_.pad[1] = 0.;
            const int steps = emitterSide ? c.s : c.t;
            if (j < steps) {                                 // BSDF sampling step at vertex j (vertex.cpp:153-271)
                rec_store(geo_slot(M, lane, side, j), ng_);  // only position + geometric normal of an inner vertex are needed again
                WalkStep ws;
                UReader rd;
                reader_open(M, c, lane, rd);
                const R2 u = rd.next2D(emitterSide ? SMP_EMITTER : SMP_SENSOR);
                const Real uz = mat_uses_sampler(BSDF) ? rd.next1D(emitterSide ? SMP_EMITTER : SMP_SENSOR) : 0.5;   // bRec.sampler->next1D()
                reader_close(rd, c);
                if (!surface_sample_next(sc, nv, nm, normalize(v.p - nv.p), emitterSide ? MODE_IMPORTANCE : MODE_RADIANCE, u, uz, ws)) { c.pstate = PS_EMPTY; break; }
                const int bit = emitterSide ? j : k - j;
                if (!ws.delta && !nv.degenerate) { c.connectable |= 1u << bit; c.flags |= F_ANYCONN; }
                c.flags = ws.delta ? (c.flags | F_DELTA) : (c.flags & ~F_DELTA);
                c.weight *= ws.weightFwd;
                c.pdfFwd = ws.pdfFwd; c.pdfBwd = ws.pdfBwd;
                emit_ray(M, lane, c, nv.p, ws.wo, sc.epsilon, INFINITY, ray);
                dest = Q_RAYC + (M.parity ^ 1);
                break;
            }
            // last vertex of this subpath: the connection needs its full record; its measure stays invalid => connectable iff
            // not degenerate
            rec_store((emitterSide ? M.lm.vs : M.lm.vt) + lane, nv);
            if (!nv.degenerate) { c.connectable |= 1u << (emitterSide ? (int) c.s : k - (int) c.t); c.flags |= F_ANYCONN; }
            if (!emitterSide) dest = mmlt_emitter_launch(M, lane, c, ray) == Q_RAYC ? Q_RAYC + (M.parity ^ 1) : Q_CONNECT;
            else { c.pstate = PS_CONNECT; dest = Q_CONNECT; }
        } while (false);
        rec_store(M.lm.core + lane, c);
        q_push_ray(M.q, dest, (uint32_t) lane, ray);
    }
}

// ------------------------------------------------------------------ connection (pathsampler.cpp:161-295)
__global__ void __launch_bounds__(128, CONNECT_MINB)
k_connect(const __grid_constant__ Machine M) {
    const DevScene &sc = M.sc;
    const PathCfg &pc = M.pc;
    const uint32_t cnt = M.q.count[Q_CONNECT];
    const uint32_t *items = M.q.items + (size_t) Q_CONNECT * M.q.n;
    for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < cnt; qi += gridDim.x * blockDim.x) {
        const int lane = (int) items[qi];
        Core c;
        rec_load(c, M.lm.core + lane);
        Real conn[4] = { 0., 0., 0., 0. };                   // pdfImp[s+1], pdfRad[s-1], pdfRad[s], pdfImp[s+2]
        int dest = Q_CHAIN + M.parity;
        RayF ray;
        c.pstate = PS_EMPTY;
        do {
            if (!(c.flags & F_ANYCONN)) break;               // pathsampler.cpp:161-174
            const int s = c.s, t = c.t, depth = c.depth;
            Vtx vt; PredRec vtp;
            rec_load(vt, M.lm.vt + lane);
            if (t >= 2) rec_load(vtp, geo_slot(M, lane, SIDE_S, t - 1));
            if (s == 0) {                                    // pure sensor path: vt must be on an emitter (:213-224)
                if (vt.type != V_SURFACE || vt.emitter < 0) break;
                const R3 n = vt.ns;                          // cast(): pRec.n = its.shFrame.n (records.inl:154-155)
                R3 wo = vtp.p - vt.p;
                const Real dist = length(wo);
                wo = wo / dist;
                const Real dp = dot(wo, n);
                if (!(dp > 0.)) break;                       // evalDirection (area.cpp:140-148) / |n.wo| = 1/pi
                c.weight = c.weight * emitter_radiance(sc, vt.emitter);   // radiance * pi * (1/pi)
                const DevEmitter &em = sc.emitters[vt.emitter];
                c.connectable |= 1u << 1;                    // emitter sample: area measure, not degenerate
                conn[0] = em.invArea * em.pdfDiscrete;                                       // vs->evalPdf: pdfEmitterPosition
                conn[3] = R_INV_PI * dp / (dist * dist) * absdot(wo, vtp.ng);                 // vt->evalPdf(vs, vtPred, EImportance)
                // connection edge of a supernode: length 0, generalized geometric term = 1 (edge.cpp:229-234, 561-571)
                if (pc.excludeDirect && depth <= 2) break;
                c.pstate = PS_FINISH;
                break;
            }
            Vtx vs; PredRec vsp;
            rec_load(vs, M.lm.vs + lane);
            if (s >= 2) rec_load(vsp, geo_slot(M, lane, SIDE_E, s - 1));
            if (vs.degenerate || vt.degenerate) break;       // :253-257
            R3 d = vs.p - vt.p;                              // from vt towards vs
            const Real len = length(d);
            if (len == 0.) break;
            d = d / len;
            R3 fs, ft;
            Mat ms, mt;
            if (s == 1) {                                    // vs->eval(vsPred, vt, EImportance)
                const Real dp = dot(-d, vs.ns);
                fs = r3(dp > 0. ? R_INV_PI : 0.);
            } else {
                ms = load_material(sc, vs.mat, vs.uv);
                fs = surface_eval(sc, vs, ms, normalize(vsp.p - vs.p), -d, MODE_IMPORTANCE);
            }
            if (t == 1) {
                const Real imp = cam_importance(sc.cam, cam_inv_dir(sc.cam, d));
                const Real dp = absdot(vt.ns, d);
                ft = r3(dp != 0. ? imp / dp : imp);
            } else {
                mt = load_material(sc, vt.mat, vt.uv);
                ft = surface_eval(sc, vt, mt, normalize(vtp.p - vt.p), d, MODE_RADIANCE);
            }
            R3 value = c.weight * fs * ft;
            if (is_zero(value)) break;
            // generalized geometric term (edge.cpp:245-267), applied before the visibility test: an occluded
            // connection is dropped whatever its value
            value *= absdot(vs.ns, d) * absdot(vt.ns, d) / (len * len);
            c.weight = value;
            // the four densities next to the connection (path.cpp:835-859)
            c.connectable |= (1u << s) | (1u << (s + 1));    // measure forced to EArea (:263-265)
            if (s == 1) {
                const Real dp = dot(-d, vs.ns);
                conn[0] = R_INV_PI * fmax(dp, 0.) / (len * len) * absdot(d, vt.ng);
                conn[1] = 1.0;
            } else {
                conn[0] = surface_pdf_area(vs, ms, vsp.p, vt.p, vt.ng);
                conn[1] = surface_pdf_area(vs, ms, vt.p, vsp.p, vsp.ng);
            }
            if (t == 1) {
                conn[2] = cam_importance(sc.cam, cam_inv_dir(sc.cam, d)) / (len * len) * absdot(d, vs.ng);
                conn[3] = 1.0;
                R2 sp = r2(0., 0.);
                if (!cam_sample_position(sc.cam, vs.p - vt.p, sp)) c.flags |= F_SPOS_FAIL;   // :298-303
                c.spos = make_float2((float) sp.x, (float) sp.y);
            } else {
                conn[2] = surface_pdf_area(vt, mt, vtp.p, vs.p, vs.ng);
                conn[3] = surface_pdf_area(vt, mt, vs.p, vtp.p, vtp.ng);
            }
            // pathConnectAndCollapse (edge.cpp:572-606): vt and vs are always "on surface" here
            c.pstate = PS_CONNECT_SHADOW;
            emit_ray(M, lane, c, vt.p, d, sc.epsilon, len * (1. - sc.shadowEpsilon), ray);
            dest = Q_RAYS + (M.parity ^ 1);
        } while (false);
        if (c.pstate != PS_EMPTY) *reinterpret_cast<double4 *>(M.lm.conn + 4 * (size_t) lane) = make_double4(conn[0], conn[1], conn[2], conn[3]);
        rec_store(M.lm.core + lane, c);
        q_push_ray(M.q, dest, (uint32_t) lane, ray);
    }
}

static unsigned grid_for(int n, int threads) { return stage_grid(n, threads); }

// the walk queues of all BSDF models in one launch, every warp on one model (multiq_locate)
__global__ void __launch_bounds__(128, WALK_MINB)
k_walk(const __grid_constant__ Machine M) {
    static_assert(Q_WALK + N_WALK_CLASSES == Q_CONNECT && DR_BSDF_ROUGHPLASTIC == N_WALK_CLASSES - 1, "one walk queue per BSDF model");
    const uint32_t cnt[N_WALK_CLASSES] = { M.q.count[Q_WALK + 0], M.q.count[Q_WALK + 1], M.q.count[Q_WALK + 2], M.q.count[Q_WALK + 3], M.q.count[Q_WALK + 4], M.q.count[Q_WALK + 5],
                                           M.q.count[Q_WALK + 6] };
    const uint32_t nWarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;; w += nWarps) {
        int cls; uint32_t qi;
        if (!multiq_locate<N_WALK_CLASSES>(cnt, w, cls, qi)) break;
        if (qi >= cnt[cls]) continue;
        if (cls == DR_BSDF_DIFFUSE) walk_lane<DR_BSDF_DIFFUSE>(M, qi);
        else if (cls == DR_BSDF_DIELECTRIC) walk_lane<DR_BSDF_DIELECTRIC>(M, qi);
        else if (cls == DR_BSDF_CONDUCTOR) walk_lane<DR_BSDF_CONDUCTOR>(M, qi);
        else if (cls == DR_BSDF_ROUGHCONDUCTOR) walk_lane<DR_BSDF_ROUGHCONDUCTOR>(M, qi);
        else if (cls == DR_BSDF_ROUGHDIELECTRIC) walk_lane<DR_BSDF_ROUGHDIELECTRIC>(M, qi);
        else if (cls == DR_BSDF_PLASTIC) walk_lane<DR_BSDF_PLASTIC>(M, qi);
        else walk_lane<DR_BSDF_ROUGHPLASTIC>(M, qi);
    }
}

void launch_walk(const Machine &M, const LaunchCfg &lc, unsigned typeMask) {
    (void) typeMask;
    const unsigned g = grid_for(lc.nLanes + N_WALK_CLASSES * 128, 128);
    k_walk<<<g, 128, 0, lc.stream>>>(M);
    k_connect<<<g, 128, 0, lc.stream>>>(M);
}
