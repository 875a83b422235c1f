// k_pt.cu -- technique=path: the unidirectional path tracer as a resumable wavefront stage.
// PathSampler EUnidirectional (src/libbidir/pathsampler.cpp:529-567) + MIPathTracer::Li
// (src/integrators/path/path.cpp:123-312) with strictNormals=false, hideEmitters=false, minDepth=0,
// directTracing=false.  Core::weight = throughput, Core::j = depth counter; the emitter-side vertex slot of
// the lane holds PtExtra, the sensor-side slot the current shading point.
#include "machine.cuh"

#ifndef PT_MINB
#define PT_MINB 2
#endif
__global__ void __launch_bounds__(128, PT_MINB)
k_pt(const __grid_constant__ Machine M) {
    const DevScene &sc = M.sc;
    const PathCfg &pc = M.pc;
    const uint32_t cnt = M.q.count[Q_PT];
    const uint32_t *items = M.q.items + (size_t) Q_PT * M.q.n;
    for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < cnt; qi += gridDim.x * blockDim.x) {
        const int lane = (int) items[qi];
        Core c;
        rec_load(c, M.lm.core + lane);
        PtExtra px;
        rec_load(px, reinterpret_cast<const PtExtra *>(M.lm.vs + lane));
        Hit hit;
        hit.t = hit.u = hit.v = 0.f; hit.tri = __ldcs(M.q.aux + (size_t) Q_PT * M.q.n + qi);
        UReader rd;
        reader_open(M, c, lane, rd);
        Vtx v;
        int dest = Q_CHAIN + M.parity;
        RayF ray;
        // The direct-illumination sample of a vertex and the BSDF-sampled ray that continues the path leave in the SAME round
        // (a path costs one round per vertex instead of two): the shadow ray is queued with Q_DEFERRED, the traversal writes its
        // verdict to neeOcc[lane], and the lane -- back with its BSDF-sampled ray -- adds the pending contribution first, i.e. in
        // the reference's order (path.cpp:196-221 before :222-290).  Only the shadow ray of a path's LAST vertex brings the lane back itself.
        enum { GO_HIT, GO_SHADE, GO_BSDF, GO_DONE } go = c.pstate == PS_PT_NEE ? GO_DONE : GO_HIT;
        if (c.flags & F_PT_NEEPEND) {
            if (!M.lm.neeOcc[lane]) px.Li += px.pending;
            c.flags &= ~F_PT_NEEPEND;
        }
        if (c.pstate == PS_PT_NEE && hit.tri < 0) px.Li += px.pending;      // the last vertex's shadow ray arrived
        bool haveShadow = false;                             // this vertex's direct-illumination sample needs its shadow ray
        R3 shO = r3(0.), shD = r3(0.);
        Real shMax = 0.;
        Mat m;                                               // material of the vertex v: loaded (textures looked up) once, by GO_SHADE
        m.type = 0; m.flags = 0; m.alpha = 0.; m.table = nullptr; m.refl = m.trans = m.eta = m.k = r3(0.);
        for (bool running = true; running;) {
            switch (go) {
            case GO_HIT: {                                   // the camera ray or a BSDF-sampled ray arrived
                if (hit.tri < 0) { go = GO_DONE; break; }
                Vtx prev;
                if (!(c.flags & F_PT_FIRST)) rec_load(prev, M.lm.vt + lane);
                const R3 o = (c.flags & F_PT_FIRST) ? cam_pos(sc.cam) : prev.p;
                const R3 d = c.d;
                Real tHit;
                fill_vertex(sc, hit, o, d, v, tHit);
                if (!(c.flags & F_PT_FIRST)) {
                    // emitter hit by the BSDF-sampled ray: MIS against direct sampling (path.cpp:242-290)
                    if (v.emitter >= 0) {
                        const R3 value = dot(v.ns, -d) > 0. ? emitter_radiance(sc, v.emitter) : r3(0.);
                        Real lumPdf = 0.;
                        // pdfEmitterDirect (scene.cpp:1057-1060, area.cpp:172-180, shape.cpp:116-126) with dRec.setQuery(ray, its)
                        if (!(c.flags & F_DELTA) && dot(d, px.refN) >= 0. && dot(d, v.ns) < 0.) {
                            const DevEmitter &em = sc.emitters[v.emitter];
                            lumPdf = em.invArea * (tHit * tHit) / absdot(d, v.ns) * em.pdfDiscrete;
                        }
                        if ((c.flags & F_PT_DIRECT) && (c.flags & F_PT_NONSPEC))
                            px.Li += c.weight * value * ((px.bsPdf * px.bsPdf) / (px.bsPdf * px.bsPdf + lumPdf * lumPdf));
                    }
                    c.flags = (c.flags & ~F_PT_EMITTED) | F_PT_DIRECT;   // rRec.type = ERadianceNoEmission
                    if (c.j++ >= pc.rrDepth) {               // Russian roulette (path.cpp:297-306)
                        const Real q = fmin(max3(c.weight) * px.eta * px.eta, (Real) 0.95f);
                        if (rd.next1D(SMP_SENSOR) >= q) { go = GO_DONE; break; }
                        c.weight = c.weight / q;
                    }
                }
                c.flags &= ~F_PT_FIRST;
                go = GO_SHADE;
                break;
            }
            case GO_SHADE: {                                 // top of the loop body for the vertex v (in registers)
                const R3 d = c.d;
                m = load_material(sc, v.mat, v.uv);
                if (v.emitter >= 0 && (c.flags & F_PT_EMITTED) && (c.flags & F_PT_NONSPEC) && dot(v.ns, -d) > 0.)
                    px.Li += c.weight * emitter_radiance(sc, v.emitter);
                if (c.j >= pc.maxDepth && pc.maxDepth > 0) { go = GO_DONE; break; }
                const R3 wi = to_local(v, -d);
                px.refN = mat_transmissive_or_backside(m) ? r3(0.) : v.ns;    // records.inl:160-164
                rec_store(M.lm.vt + lane, v);
                px.dIn = d;
                // ---- direct illumination (scene.cpp:879-904, area.cpp:156-170, shape.cpp:102-114)
                if ((c.flags & F_PT_DIRECT) && mat_has_smooth(m.type) && sc.nEmitters > 0) {
                    const R2 u = rd.next2D(SMP_SENSOR);
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
                            px.pending = c.weight * value * bsdfVal * ((pdf * pdf) / (pdf * pdf + bp * bp));
                        }
                        haveShadow = true; shO = v.p; shD = dd; shMax = dist * (1. - sc.shadowEpsilon);
                    }
                }
                go = GO_BSDF;
                break;
            }
            case GO_BSDF: {                                  // BSDF sampling (path.cpp:222-240); always entered from GO_SHADE: m is this vertex's
                const R3 d = c.d;
                const R3 wi = to_local(v, -d);
                const R2 ub = rd.next2D(SMP_SENSOR);
                const Real uz = mat_uses_sampler(m.type) ? rd.next1D(SMP_SENSOR) : 0.5;      // bRec.sampler->next1D() (roughdielectric.cpp:555)
                BsdfSample bs;
                bsdf_sample(m, wi, MODE_RADIANCE, ub.x, ub.y, uz, sc.epsilon, bs);
                if (is_zero(bs.weight)) { go = GO_DONE; break; }
                if (!(bs.sampledType & BT_DELTA)) c.flags |= F_PT_NONSPEC;
                c.flags = (bs.sampledType & BT_DELTA) ? (c.flags | F_DELTA) : (c.flags & ~F_DELTA);
                // throughput and eta are only read again if the ray hits something (path.cpp:268-274)
                c.weight *= bs.weight;
                px.eta *= bs.eta;
                px.bsPdf = bs.pdf;
                c.pstate = PS_PT_HIT;
                emit_ray(M, lane, c, v.p, to_world(v, bs.wo), sc.epsilon, INFINITY, ray);
                dest = Q_RAYC + (M.parity ^ 1);
                if (haveShadow) {                            // ... and the shadow ray beside it
                    Real tmin = sc.epsilon;
                    tmin *= fmax(fmax(fmax(fabs(shO.x), fabs(shO.y)), fabs(shO.z)), (Real) sc.epsilon);       // as emit_ray
                    RayF sh;
                    sh.a = make_float4((float) shO.x, (float) shO.y, (float) shO.z, (float) tmin);
                    sh.b = make_float4((float) shD.x, (float) shD.y, (float) shD.z, (float) shMax);
                    double2 *r2 = reinterpret_cast<double2 *>(M.lm.rayd2 + 8 * (size_t) lane);
                    r2[0] = make_double2(shO.x, shO.y); r2[1] = make_double2(shO.z, shD.x); r2[2] = make_double2(shD.y, shD.z); r2[3] = make_double2(tmin, shMax);
                    ++c.nrays;
                    c.flags |= F_PT_NEEPEND;
                    q_push_ray(M.q, Q_RAYS + (M.parity ^ 1), (uint32_t) lane | Q_DEFERRED, sh);
                    haveShadow = false;
                }
                running = false;
                break;
            }
            default:
                if (haveShadow) {                            // the path ends here, but its last direct-illumination sample is still to be resolved
                    c.pstate = PS_PT_NEE;
                    emit_ray(M, lane, c, shO, shD, sc.epsilon, shMax, ray);
                    dest = Q_RAYS + (M.parity ^ 1);
                } else c.pstate = PS_PT_DONE;
                running = false;
                break;
            }
        }
        reader_close(rd, c);
        rec_store(reinterpret_cast<PtExtra *>(M.lm.vs + lane), px);
        rec_store(M.lm.core + lane, c);
        q_push_ray(M.q, dest, (uint32_t) lane, ray);
    }
}

void launch_pt(const Machine &M, const LaunchCfg &lc) {
    const unsigned g = stage_grid(lc.nLanes, 128);
    k_pt<<<g, 128, 0, lc.stream>>>(M);
}
