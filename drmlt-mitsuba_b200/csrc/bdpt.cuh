// bdpt.cuh -- pieces of the BDPT stage shared by begin.cuh (path start) and k_bdpt.cu (walks and connections).
#pragma once
#include "machine.cuh"


DR_D Vtx *bvp(const Machine &M, int lane, int side, int v) { return M.lm.bv + ((size_t) lane * 2 + side) * BD_MAXV + v; }
DR_D BExtra *bxp(const Machine &M, int lane, int side, int v) { return M.lm.bx + ((size_t) lane * 2 + side) * BD_MAXV + v; }

// Russian roulette of randomWalk step `i` (vertex.cpp:307-322): returns false when the walk is terminated
DR_D bool bd_roulette(const Machine &M, Core &c, UReader &rd, int sampler, int i, Real &rrWeight) {
    rrWeight = 1.0;
    if (i < M.pc.rrDepth) return true;
    const Real q = fmin(max3(c.weight), (Real) 0.95f);
    if (rd.next1D(sampler) > q) return false;
    rrWeight = 1.0 / q;
    c.weight *= rrWeight;
    return true;
}

// BSDF sampling step at surface vertex j of `side`; emits the ray on success
DR_D bool bd_sample_surface(const Machine &M, int lane, Core &c, UReader &rd, int side, int j, const Vtx &v, R3 predP, RayF &ray) {
    const int mode = side == BD_E ? MODE_IMPORTANCE : MODE_RADIANCE;
    const Mat m = load_material(M.sc, v.mat, v.uv);
    WalkStep ws;
    const R2 u = rd.next2D(side == BD_E ? SMP_EMITTER : SMP_SENSOR);
    const Real uz = mat_uses_sampler(m.type) ? rd.next1D(side == BD_E ? SMP_EMITTER : SMP_SENSOR) : 0.5;   // bRec.sampler->next1D()
    if (!surface_sample_next(M.sc, v, m, normalize(predP - v.p), mode, u, uz, ws)) return false;
    if (mode == MODE_RADIANCE && ws.eta != 1.) c.weight *= ws.eta * ws.eta;     // vertex.cpp:264-266
    c.weight *= ws.weightFwd;
    Real rrWeight;
    if (!bd_roulette(M, c, rd, side == BD_E ? SMP_EMITTER : SMP_SENSOR, j, rrWeight)) return false;
    BExtra cur, nxt;
    rec_load(cur, bxp(M, lane, side, j));
    memset(&nxt, 0, sizeof(nxt));
    nxt.prefix = cur.prefix * ws.weightFwd * rrWeight;
    rec_store(bxp(M, lane, side, j + 1), nxt);
    c.flags = ws.delta ? (c.flags | F_DELTA) : (c.flags & ~F_DELTA);
    c.pdfFwd = ws.pdfFwd; c.pdfBwd = ws.pdfBwd;
    c.j = (uint8_t) j;
    emit_ray(M, lane, c, v.p, ws.wo, M.sc.epsilon, INFINITY, ray);
    return true;
}

// start of the sensor subpath: supernode -> sensor sample -> first ray (vertex.cpp:74-97, 126-151)
DR_D bool bd_sensor_start(const Machine &M, int lane, Core &c, UReader &rd, BdAcc &acc, RayF &ray) {
    const DevScene &sc = M.sc;
    c.weight = r3(1.);
    (void) rd.next2D(SMP_SENSOR);                                  // sampleSensorPosition consumes 2
    // the supernode step returns before the throughput / Russian-roulette block (vertex.cpp:74-97): no roulette, and the
    // walk's throughput does not include this step's weight
    Real rrWeight = 1.0;
    Vtx vt;
    memset(&vt, 0, sizeof(vt));
    vt.p = cam_pos(sc.cam); vt.ng = vt.ns = cam_dir(sc.cam); vt.ss = r3(0.); vt.type = V_SENSOR_SAMPLE; vt.degenerate = 0; vt.mat = -1; vt.emitter = -1;
    rec_store(bvp(M, lane, BD_S, 1), vt);
    BExtra x0, x1;
    memset(&x0, 0, sizeof(x0)); memset(&x1, 0, sizeof(x1));
    x0.prefix = r3(1.); x0.fwd = 1.0; x0.discrete = 1;             // sensor supernode: measure EDiscrete
    x1.prefix = r3(rrWeight); x1.fwd = 1.0;                        // pdf[ERadiance] of the supernode (perspective.cpp:305)
    rec_store(bxp(M, lane, BD_S, 0), x0);
    rec_store(bxp(M, lane, BD_S, 1), x1);
    acc.nt = 1;
    if (M.pc.maxDepth + 1 < 2) return false;
    const R2 u = rd.next2D(SMP_SENSOR);
    const R3 dl = cam_sample_to_dir(sc.cam, u.x, u.y);
    // weight[ERadiance] = 1 (vertex.cpp:139-144)
    if (!bd_roulette(M, c, rd, SMP_SENSOR, 1, rrWeight)) return false;
    BExtra x2;
    memset(&x2, 0, sizeof(x2));
    x2.prefix = x1.prefix * rrWeight;
    rec_store(bxp(M, lane, BD_S, 2), x2);
    c.flags &= ~F_DELTA;
    c.pdfFwd = sc.cam.normalization / (dl.z * dl.z * dl.z); c.pdfBwd = 1.0;
    c.j = 1;
    c.pstate = PS_BD_SHIT;
    emit_ray(M, lane, c, vt.p, cam_xform_dir(sc.cam, dl), sc.epsilon, INFINITY, ray);
    return true;
}


// Start of a BDPT path (called by begin_path): emitter supernode -> emitter sample -> first emitter ray, or -- when the
// emitter subpath cannot be extended -- the start of the sensor subpath.  Returns the queue (Q_RAYC) or -1.
DR_D int bdpt_path_start(const Machine &M, int lane, Core &c, UReader &rd, RayF &ray) {
    const DevScene &sc = M.sc;
    BdAcc acc;
    memset(&acc, 0, sizeof(acc));
    c.flags = 0; c.weight = r3(1.); c.s = c.t = 0; c.connectable = 0;
    BExtra x0;
    memset(&x0, 0, sizeof(x0));
    x0.prefix = r3(1.); x0.fwd = 1.0; x0.bwd = 1.0;
    rec_store(bxp(M, lane, BD_E, 0), x0);
    bool launched = false;
    if (sc.nEmitters > 0 && M.pc.maxDepth >= 1) {                      // step 0: supernode -> emitter sample (vertex.cpp:50-72)
        EmitterPoint ep;
        const R2 u0 = rd.next2D(SMP_EMITTER);
        sample_emitter_point(sc, u0.x, u0.y, ep);
        const DevEmitter &em = sc.emitters[ep.emitter];
        const R3 w0 = emitter_radiance(sc, ep.emitter) * (R_PI * em.area / ep.emPdf);
        // the supernode step returns before the throughput / Russian-roulette block (vertex.cpp:50-72): the walk's
        // throughput stays 1 (it never sees the emitted power) and step 0 plays no roulette
        Real rrWeight = 1.0;
        if (!is_zero(w0)) {
            Vtx vs;
            memset(&vs, 0, sizeof(vs));
            vs.p = ep.p; vs.ng = vs.ns = ep.n; vs.ss = r3(0.); vs.type = V_EMITTER_SAMPLE; vs.degenerate = 0; vs.emitter = ep.emitter; vs.mat = -1;
            rec_store(bvp(M, lane, BD_E, 1), vs);
            BExtra x1;
            memset(&x1, 0, sizeof(x1));
            x1.prefix = w0 * rrWeight; x1.fwd = ep.pdfArea;
            rec_store(bxp(M, lane, BD_E, 1), x1);
            acc.ns = 1;
            if (M.pc.maxDepth >= 2) {                                  // step 1: emission direction (vertex.cpp:99-124, area.cpp:130-138)
                const R2 u = rd.next2D(SMP_EMITTER);
                const R3 local = square_to_cosine_hemisphere(u.x, u.y);
                R3 fs, ft;
                coordinate_system(vs.ns, fs, ft);
                if (bd_roulette(M, c, rd, SMP_EMITTER, 1, rrWeight)) {  // weight[EImportance] = 1
                    BExtra x2;
                    memset(&x2, 0, sizeof(x2));
                    x2.prefix = x1.prefix * rrWeight;
                    rec_store(bxp(M, lane, BD_E, 2), x2);
                    c.flags &= ~F_DELTA;
                    c.pdfFwd = R_INV_PI * local.z; c.pdfBwd = 1.0;
                    c.j = 1;
                    c.pstate = PS_BD_EHIT;
                    emit_ray(M, lane, c, vs.p, fs * local.x + ft * local.y + vs.ns * local.z, sc.epsilon, INFINITY, ray);
                    launched = true;
                }
            }
        }
    }
    if (!launched) launched = bd_sensor_start(M, lane, c, rd, acc, ray);
    rec_store(M.lm.bacc + lane, acc);
    if (launched) return Q_RAYC;
    c.pstate = PS_BD_DONE;                                             // nothing to trace: an empty splat list
    return -1;
}

