// k_chain.cu -- the chain kernel of the wavefront machine (machine.cuh) and the lane set-up kernels.
//
// k_chain consumes the lanes whose path has just ended (or which have no path yet).  In ONE thread per
// lane it (1) turns the finished path into a result (Path::miWeight, src/libbidir/path.cpp:763-1028, and the
// single MMLT splat, pathsampler.cpp:288-313), (2) runs the chain-level step: delayed-rejection acceptance
// (DRMLTRenderer::process, src/integrators/drmlt/drmlt_proc.cpp:539-769; processMixture :161-380;
// PSSMLTRenderer::process, src/integrators/pssmlt/pssmlt_proc.cpp:175-272), expectation-weighted film splats,
// commit of the accepted primary-sample vector and statistics, and (3) hands the lane to the k_begin class that mutates its
// next proposal and emits the first ray of the next path (begin.cuh).
#include "begin.cuh"

namespace {

DR_D bool invalid_strict(Real x) { return isnan(x) || isinf(x) || x <= 0.; }   // drmlt_proc.cpp:428
DR_D bool invalid_loose(Real x) { return isnan(x) || isinf(x) || x < 0.; }     // drmlt_proc.cpp:181
DR_D Real metropolis_clamp(Real x) { return x < 1.0 ? x : 1.0; }               // std::min(1, x): NaN -> 1

DR_D void result_clear(PathResult &r) { r.lum = 0.; r.n = 0; r.nl = 0; r.val = r3(0.); r.pos = r2(0., 0.); r.mis = 0.; r.s = r.t = -1; }

DR_D float3 normalized_value(const PathResult &r) {          // SplatList::normalize (pathsampler.cpp:1001-1028)
    const Real inv = r.lum > 0. ? 1.0 / r.lum : 1.0;
    return to_f3(r.val * inv);
}

// ---- light-image splat lists of technique=bdpt (list 0 = x, 1 = y, 2 = z, 3 = the path that just ended)
DR_D float4 *splat_list(const Machine &M, int lane, int list) { return M.lm.bsplat + ((size_t) lane * 4 + list) * BD_MAXS * 2; }
// copy list `src` to `dst`, scaling the values (SplatList::normalize when src = 3)
DR_D void splat_list_copy(const Machine &M, int lane, int dst, int src, int n, float scale) {
    const float4 *s = splat_list(M, lane, src);
    float4 *d = splat_list(M, lane, dst);
    for (int i = 0; i < n; ++i) {
        const float4 v = s[2 * i + 1];
        d[2 * i] = s[2 * i];
        d[2 * i + 1] = make_float4(v.x * scale, v.y * scale, v.z * scale, 0.f);
    }
}
DR_D void splat_list_put(const Machine &M, int lane, int list, int n, float weight) {
    const float4 *s = splat_list(M, lane, list);
    for (int i = 0; i < n; ++i) {
        const float4 p = s[2 * i], v = s[2 * i + 1];
        film_put(M.film, M.fp, make_float2(p.x, p.y), make_float3(v.x * weight, v.y * weight, v.z * weight));
    }
}
DR_D float inv_lum(const PathResult &r) { return r.lum > 0. ? (float) (1.0 / r.lum) : 1.f; }

// MiraDRMLTSampler::getTransitionRatio over the three samplers (drmlt_sampler.cpp:400-414).
// dimStage holds the largest INDEX touched, so the last used coordinate is skipped.
DR_D Real mira_transition_ratio(const Machine &M, const MutCtx &mc, const double *ub, const uint8_t posY[3], const int posZ[3]) {
    const ChainParams &cp = M.cp;
    const int nU = M.lm.nU;
    Real ratio = 1.0;
    for (int s = 0; s < 3; ++s) {
        if (mc.identity1(s)) continue;
        Real num = 0., den = 0.;
        const int dimStage = max(max((int) posY[s] - 1, 0), max(posZ[s] - 1, 0));
        for (int i = 0; i < dimStage; i += 2) {
            const int slot = M.pp.off[s] + i;
            const R2 x = ub_load(ub + UB_X * nU, slot), y = ub_load(ub + UB_Y * nU, slot), z = ub_load(ub + UB_Z * nU, slot);
            num += kelemen_logpdf(z.x - y.x, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
            den += kelemen_logpdf(x.x - y.x, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
            if (i + 1 < dimStage) {
                num += kelemen_logpdf(z.y - y.y, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
                den += kelemen_logpdf(x.y - y.y, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
            }
        }
        ratio *= exp(num - den);
    }
    return ratio;
}

// write the accepted proposal back as the new current state (DRMLTSampler::accept, drmlt_sampler.cpp:189-199)
DR_D void commit_state(const Machine &M, const MutCtx &mc, double *ub, const int dims[3], int from, bool first, int s_, int t_, bool drmlt) {
    const int nU = M.lm.nU;
    int ext[3];
    pair_extent(M, dims, s_, t_, ext);
    for (int s = 0; s < 3; ++s) {
        if (!mc.largeStep && (first ? mc.identity1(s) : mc.identity2(s))) continue;
        for (int p = 0; p < ext[s]; ++p) {
            const int slot = M.pp.off[s] + 2 * p;
            R2 v = ub_load(ub + from * nU, slot);
            if (drmlt) { v.x = wrap_reflect(v.x); v.y = wrap_reflect(v.y); }
            ub_store(ub + UB_X * nU, slot, v);
        }
    }
}

// ------------------------------------------------------------------ path end -> result
DR_D void path_result(const Machine &M, int lane, int tri, Core &c, PathResult &out) {
    result_clear(out);
    if (M.pc.technique == DR_TECH_MMLT) { out.s = c.s; out.t = c.t; }
    if (c.pstate == PS_CONNECT_SHADOW) {
        if (tri >= 0) return;                                 // occluded
        if (M.pc.excludeDirect && c.depth <= 2) return;       // pathsampler.cpp:279-284
        if (c.flags & F_SPOS_FAIL) return;
        c.pstate = PS_FINISH;
    }
    if (c.pstate == PS_FINISH) {                              // pathsampler.cpp:288-313
        const int depth = c.depth, k = depth + 2, s = c.s, t = c.t;
        // assemble pdfImp / pdfRad / conv of the full path (vertex i <= s: emitter subpath vertex i, else sensor subpath
        // vertex k - i) from the per-step records of the two walks and the four densities of the connection
        MisArrays A;
        A.connectable = c.connectable;
#pragma unroll
        for (int i = 0; i <= DR_MAXK; ++i) { A.pdfImp[i] = 0.; A.pdfRad[i] = 0.; A.conv[i] = 0.; }
        A.pdfImp[0] = 1.0; A.pdfRad[k] = 1.0; A.pdfRad[k - 1] = 1.0;       // supernodes; sensor sample (perspective.cpp:305)
        for (int j = 0; j < s; ++j) {                         // emitter walk step j produced vertex j + 1
            const double4 r = *reinterpret_cast<const double4 *>(misrec_slot(M, lane, SIDE_E, j));
            A.pdfImp[j + 1] = r.x;
            if (j >= 1) { A.pdfRad[j - 1] = r.y; A.conv[j] = r.z; }
        }
        for (int j = 1; j < t; ++j) {                         // sensor walk step j produced vertex j + 1 (global index k - j - 1)
            const double4 r = *reinterpret_cast<const double4 *>(misrec_slot(M, lane, SIDE_S, j));
            A.pdfRad[k - j - 1] = r.x;
            A.pdfImp[k - j + 1] = r.y;
            A.conv[k - j - 1] = r.z;
        }
        {
            const double4 q = *reinterpret_cast<const double4 *>(M.lm.conn + 4 * (size_t) lane);
            A.pdfImp[s + 1] = q.x;
            if (s >= 1) A.pdfRad[s - 1] = q.y;
            if (s >= 1) A.pdfRad[s] = q.z;                     // (s = 0: pdfRad[0] is not used by the sweep)
            if (s + 2 <= k) A.pdfImp[s + 2] = q.w;
        }
        const Real w = mis_weight(A, s, t, M.pc.lightImage != 0);
        const int nStrats = M.pc.lightImage ? depth + 1 : depth;
        const R3 value = c.weight * (w * (Real) nStrats);
        out.mis = w; out.n = 1; out.pos = r2(c.spos.x, c.spos.y); out.val = value; out.lum = luminance(value);
    } else if (c.pstate == PS_BD_DONE) {                      // splat 0 + light-image splats (the lane's in-flight list)
        BdAcc acc;
        rec_load(acc, M.lm.bacc + lane);
        out.n = acc.has0; out.nl = acc.nl;
        out.s = acc.ns; out.t = acc.nt;                       // diagnostic: index of the last vertex of each subpath
        out.pos = r2(acc.pos0.x, acc.pos0.y); out.val = acc.val0; out.lum = acc.lum;
    } else if (c.pstate == PS_PT_DONE) {
        PtExtra px;
        rec_load(px, reinterpret_cast<const PtExtra *>(M.lm.vs + lane));
        out.n = 1; out.pos = r2(c.spos.x, c.spos.y); out.val = px.Li; out.lum = luminance(px.Li);
    }
}

// SplatList::normalize(importanceMap), first half (pathsampler.cpp:1001-1020): two-stage MLT divides every non-zero
// splat by the importance of its pixel and re-derives the list's luminance from the re-weighted splats.  The chain then
// runs on that luminance; develop multiplies the importance back (drmlt_proc.cpp:823-849).
DR_D Real importance_at(const Machine &M, Real px, Real py) {
    const int x = min(max(0, (int) px), M.fp.w - 1), y = min(max(0, (int) py), M.fp.h - 1);
    return (Real) __ldg(M.cp.importance + (size_t) y * M.fp.w + x);
}
DR_D void importance_apply(const Machine &M, int lane, PathResult &r) {
    Real lum = 0.;
    if (r.n && !is_zero(r.val)) {
        r.val = r.val * (1.0 / importance_at(M, r.pos.x, r.pos.y));
        lum += luminance(r.val);
    }
    if (r.nl) {                                                   // bdpt: the light-image splats of the path in flight
        float4 *sl = splat_list(M, lane, 3);
        for (int i = 0; i < r.nl; ++i) {
            const float4 p = sl[2 * i], v = sl[2 * i + 1];
            if (v.x == 0.f && v.y == 0.f && v.z == 0.f) continue;
            const R3 w = r3((Real) v.x, (Real) v.y, (Real) v.z) * (1.0 / importance_at(M, (Real) p.x, (Real) p.y));
            sl[2 * i + 1] = make_float4((float) w.x, (float) w.y, (float) w.z, 0.f);
            lum += luminance(w);
        }
    }
    r.lum = lum;
}

// A lane takes up chain `idx` of the job: seed replay + fillReplay (drmlt_proc.cpp:467-504) -- current = the seed's
// bootstrap vector; the lane then evaluates it (PH_INIT) before its first mutation.
DR_D void chain_init(const Machine &M, int lane, Core &c, const int *depthIn, const unsigned long long *chainIdIn, const unsigned long long *seedIdxIn, size_t idx) {
    memset(&c, 0, sizeof(c));
    c.tx = -1; c.large = 2u;
    const unsigned long long sidx = seedIdxIn[idx];
    c.chainId = chainIdIn[idx]; c.seedIdx = sidx;
    c.depth = M.pc.technique == DR_TECH_MMLT ? (uint8_t) depthIn[idx] : 0;
    double *ub = M.lm.ubuf + (size_t) lane * M.lm.ubCount * M.lm.nU;
    const int alloc[3] = { M.cp.dimS, M.cp.dimE, M.cp.dimD };
    for (int k = 0; k < M.lm.nU; ++k) ub[k] = 0.0;
    for (int s = 0; s < 3; ++s)
        for (int k = 0; k < alloc[s]; ++k)
            if (M.job.replay) {          // the recorded chain's seed state (a coordinate it never had reads 0.5)
                const double v = k < M.job.replayDim ? M.job.replay[(size_t) lane * M.job.replayStride + (size_t) s * M.job.replayDim + k] : NAN;
                ub[M.pp.off[s] + k] = isnan(v) ? 0.5 : v;
            } else ub[M.pp.off[s] + k] = (double) keyed_uniform(M.pp.seed, S_BOOT, sidx, (uint32_t) s, (uint32_t) k);
    c.phase = PH_INIT;
    c.pstate = PS_START;
}

} // namespace

// ------------------------------------------------------------------ the chain kernel

// FUSED: the lane's next path is started here instead of by k_begin -- one launch (and its ~20 us latency floor) less per round.  For
// small PSSMLT groups only (one proposal class): C1 +6 %; with drmlt's three proposal classes diverging inside a warp it loses 2-3 % on
// C2 / C3 / C4 and more at C5 size (164 registers).
template <bool FUSED>
__global__ void __launch_bounds__(128, CHAIN_MINB)
k_chain(const __grid_constant__ Machine M) {
    const JobParams &job = M.job;
    const PssParams &pp = M.pp;
    const ChainParams &cp = M.cp;
    const FilmParams &fp = M.fp;
    float4 *film = M.film;
    const bool drmlt = pp.integrator == DR_INTEGRATOR_DRMLT;
    uint32_t st[ST_COUNT];
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) st[i] = 0;
    if (FUSED) queues_recycle(M.q);                           // (the last kernel of the round empties the in-round queues)
    const uint32_t cnt = M.q.count[Q_CHAIN + M.parity];
    const uint32_t *items = M.q.items + (size_t) (Q_CHAIN + M.parity) * M.q.n;
    for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < cnt; qi += gridDim.x * blockDim.x) {
        const int lane = (int) items[qi];
        Core c;
        rec_load(c, M.lm.core + lane);
        int dest = -1;
        for (;;) {
            long long item = 0;
            MutCtx mc;
            mc.pp = &pp; mc.chain = c.chainId; mc.mut = c.mut; mc.largeStep = false; mc.lightTracing = false; mc.table = replay_table(M, lane);
            if (job.type != JOB_CHAIN) item = (long long) lane + (long long) c.mut * M.lm.n;
            else {
                mc.largeStep = c.phase != PH_INIT && c.large == 1u;
                mc.lightTracing = c.phase == PH_STAGE2 && cp.fixEmitterPath && c.tx == 1;   // nextStage(current->t == 1)
            }
            if (c.pstate != PS_START) {
                // ================= a path is complete: job-level step =================
                PathResult r;
                path_result(M, lane, __ldcs(M.q.aux + (size_t) (Q_CHAIN + M.parity) * M.q.n + qi), c, r);
                ++st[ST_PATHS];
                if (job.type == JOB_BOOT) {
                    job.lumOut[item] = (float) r.lum;
                    if (job.lumTargetOut) {                   // the density the chains will actually sample (importance_apply)
                        if (cp.importance) importance_apply(M, lane, r);
                        job.lumTargetOut[item] = (float) r.lum;
                    }
                    ++c.mut;
                    c.pstate = ((long long) lane + (long long) c.mut * M.lm.n < job.nItems) ? PS_START : PS_IDLE;
                } else if (job.type == JOB_EVAL) {
                    dr_path_result o;
                    memset(&o, 0, sizeof(o));
                    o.luminance = (float) r.lum; o.n_splats = r.n; o.s = r.s; o.t = r.t; o.mis_weight = (float) r.mis;
                    o.pos[0][0] = (float) r.pos.x; o.pos[0][1] = (float) r.pos.y;
                    o.value[0][0] = (float) r.val.x; o.value[0][1] = (float) r.val.y; o.value[0][2] = (float) r.val.z;
                    if (r.nl) {                               // bdpt: light-image splats follow splat 0 (if it exists)
                        const float4 *sl = splat_list(M, lane, 3);
                        for (int i = 0; i < r.nl && r.n + i < DR_MAX_SPLATS; ++i) {
                            const float4 sp = sl[2 * i], sv = sl[2 * i + 1];
                            o.pos[r.n + i][0] = sp.x; o.pos[r.n + i][1] = sp.y;
                            o.value[r.n + i][0] = sv.x; o.value[r.n + i][1] = sv.y; o.value[r.n + i][2] = sv.z;
                        }
                        o.n_splats = r.n + r.nl;
                    }
                    o.n_rays = c.nrays;
                    job.out[item] = o;
                    ++c.mut;
                    c.pstate = ((long long) lane + (long long) c.mut * M.lm.n < job.nItems) ? PS_START : PS_IDLE;
                } else {
                    if (cp.importance) importance_apply(M, lane, r);
                    ChainCore cc;
                    rec_load(cc, M.lm.chain + lane);
                    double *ub = M.lm.ubuf + (size_t) lane * M.lm.ubCount * M.lm.nU;
                    int dims[3];
                    chain_dims(M.pc, cp, c.depth, dims);
                    bool mutationDone = false;
                    Real a2 = 0.; bool acc2 = false;
                    const bool largeStep = c.large == 1u;
                    const float2 rpos = make_float2((float) r.pos.x, (float) r.pos.y);
                    if (c.phase == PH_INIT) {
                        // seed replay (drmlt_proc.cpp:467-512): the bootstrap vector becomes the current state
                        cc.Lx = r.lum; cc.posx = rpos; cc.valx = normalized_value(r);
                        cc.xl = (uint8_t) r.nl;
                        if (r.nl) splat_list_copy(M, lane, 0, 3, r.nl, inv_lum(r));
                        cc.cumW = 0.; c.tx = (int8_t) r.t;
                        c.phase = PH_STAGE1; c.large = 2u;
                        c.pstate = c.mut < mut_target(job, c.depth) ? PS_START : PS_IDLE;
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
                            accept = (a == 1.) || (chain_coin(M, lane, c, 1) < a);
                        } else {
                            currentWeight = cp.kelemenWeights ? cc.Lx / (cc.Lx / cp.b + cp.pLarge) : 1.;
                            proposedWeight = 0.; accept = false;
                        }
                        cc.cumW += currentWeight;
                        if (job.records) {
                            dr_step_record &rec = job.records[(size_t) lane * job.recordStride + (c.mut - job.mut0)];
                            rec.L_x = (float) cc.Lx; rec.L_y = (float) yL; rec.L_z = 0.f; rec.a1 = (float) a; rec.a2 = 0.f;
                            rec.large_step = largeStep; rec.accept1 = accept; rec.did_second = 0; rec.accept2 = 0;
                        }
                        ++st[ST_ACC_B];
                        if (largeStep) ++st[ST_LARGE_B]; else ++st[ST_BOLD_B];
                        const float3 yval = normalized_value(r);
                        if (accept) {
                            const float3 v = cc.valx * (float) cc.cumW;
                            if (film && !is_zero(v)) film_put(film, fp, cc.posx, v);
                            if (film && cc.xl) splat_list_put(M, lane, 0, cc.xl, (float) cc.cumW);
                            cc.cumW = proposedWeight;
                            commit_state(M, mc, ub, dims, UB_Y, true, r.s, r.t, false);
                            cc.Lx = yL; cc.posx = rpos; cc.valx = yval; c.tx = (int8_t) r.t;
                            cc.xl = (uint8_t) r.nl;
                            if (r.nl) splat_list_copy(M, lane, 0, 3, r.nl, inv_lum(r));
                            ++st[ST_ACC_A];
                            if (largeStep) ++st[ST_LARGE_A]; else ++st[ST_BOLD_A];
                        } else {
                            const float3 v = yval * (float) proposedWeight;
                            if (film && r.n && !is_zero(v)) film_put(film, fp, rpos, v);
                            if (film && r.nl && proposedWeight > 0.) splat_list_put(M, lane, 3, r.nl, (float) proposedWeight * inv_lum(r));
                        }
                        mutationDone = true;
                    } else if (c.phase == PH_STAGE1) {
                        // ---------------- DRMLT first stage (drmlt_proc.cpp:539-558; mixture :284-299)
                        ++st[ST_MUT];
                        cc.yL = r.lum; cc.ypos = rpos; cc.yval = normalized_value(r); cc.yn = (uint8_t) r.n; cc.yt = (int8_t) r.t; cc.ys = (int8_t) r.s;
                        cc.yl = (uint8_t) r.nl; cc.zl = 0;
                        if (r.nl) splat_list_copy(M, lane, 1, 3, r.nl, inv_lum(r));
                        cc.posY[0] = c.pos0; cc.posY[1] = c.pos1; cc.posY[2] = c.pos2;
                        cc.a1 = 0.; cc.acc1 = 0; cc.zL = 0.; cc.zn = 0;
                        bool doSecond;
                        if (cp.useMixture) {
                            if (!invalid_loose(cc.yL)) {
                                cc.a1 = metropolis_clamp(cc.yL / cc.Lx);
                                cc.acc1 = (cc.a1 >= 1.) || (chain_coin(M, lane, c, 1) < cc.a1);
                            }
                            doSecond = !largeStep && (chain_coin(M, lane, c, 3) < 0.5);
                        } else {
                            if (!invalid_strict(cc.yL)) {
                                cc.a1 = metropolis_clamp(cc.yL / cc.Lx);
                                cc.acc1 = (cc.a1 >= 1.) || (chain_coin(M, lane, c, 1) < cc.a1);
                            }
                            doSecond = !cc.acc1 && (cp.timidAfterLarge || !largeStep);
                        }
                        if (doSecond) { c.phase = PH_STAGE2; c.pstate = PS_START; }
                        else mutationDone = true;
                    } else if (c.phase == PH_STAGE2) {
                        cc.zL = r.lum; cc.zpos = rpos; cc.zval = normalized_value(r); cc.zn = (uint8_t) r.n; cc.zt = (int8_t) r.t; cc.zs = (int8_t) r.s;
                        cc.zl = (uint8_t) r.nl;
                        if (r.nl) splat_list_copy(M, lane, 2, 3, r.nl, inv_lum(r));
                        mutationDone = true;
                        if (cp.useMixture) {   // drmlt_proc.cpp:317-324: plain MH on the replaced proposal
                            cc.a1 = 0.; cc.acc1 = 0;
                            if (!invalid_loose(cc.zL)) {
                                a2 = metropolis_clamp(cc.zL / cc.Lx);
                                acc2 = (a2 >= 1.) || (chain_coin(M, lane, c, 2) < a2);
                            }
                        } else if (!invalid_strict(cc.zL)) {
                            if (pp.type == DR_TYPE_GREEN) { c.phase = PH_REVERSE; c.pstate = PS_START; mutationDone = false; }
                            else if (pp.type == DR_TYPE_MIRA) {   // drmlt_proc.cpp:625-650
                                const Real aReverse = metropolis_clamp(cc.yL / cc.zL);
                                if (!(aReverse >= 1.)) {
                                    const int posZ[3] = { c.pos0, c.pos1, c.pos2 };
                                    const Real T = largeStep ? 1.0 : mira_transition_ratio(M, mc, ub, cc.posY, posZ);
                                    if (!invalid_strict(T)) {
                                        a2 = metropolis_clamp((cc.zL / cc.Lx) * T * (1.0 - aReverse) / (1.0 - cc.a1));
                                        acc2 = (a2 >= 1.) || (chain_coin(M, lane, c, 2) < a2);
                                    }
                                }
                            } else {                               // orbital, drmlt_proc.cpp:655-669
                                if (cc.zL < cc.yL) { a2 = 0.; }
                                else if (cc.zL >= cc.Lx) { a2 = 1.0; acc2 = true; }
                                else {
                                    a2 = (cc.zL - cc.yL) / (cc.Lx - cc.yL);
                                    acc2 = (a2 >= 1.) || (chain_coin(M, lane, c, 2) < a2);
                                }
                            }
                        }
                    } else {                   // Green's reverse path (drmlt_proc.cpp:588-621)
                        const Real Lr = r.lum;
                        const Real aReverse = invalid_strict(Lr) ? 0. : metropolis_clamp(Lr / cc.zL);
                        if (aReverse != 1.) {
                            a2 = metropolis_clamp((cc.zL / cc.Lx) * (1. - aReverse) / (1. - cc.a1));
                            acc2 = (a2 >= 1.) || (chain_coin(M, lane, c, 2) < a2);
                        }
                        mutationDone = true;
                    }

                    if (mutationDone && drmlt) {
                        // ---- splat with expectation weights (drmlt_proc.cpp:676-688; mixture :327-333)
                        const bool did2 = c.phase != PH_STAGE1;
                        const bool acc1 = cc.acc1 != 0;
                        const Real a1 = cc.a1;
                        if (job.records) {
                            dr_step_record &rec = job.records[(size_t) lane * job.recordStride + (c.mut - job.mut0)];
                            rec.L_x = (float) cc.Lx; rec.L_y = (float) cc.yL; rec.L_z = did2 ? (float) cc.zL : 0.f; rec.a1 = (float) a1; rec.a2 = (float) a2;
                            rec.large_step = largeStep; rec.accept1 = acc1; rec.did_second = did2; rec.accept2 = acc2;
                        }
                        if (film && !cp.acceptanceMap) {
                            Real wy, wz, wx;
                            if (cp.useMixture) { wy = did2 ? 0. : a1; wz = did2 ? a2 : 0.; wx = 1.0 - (did2 ? a2 : a1); }
                            else { wy = a1; wz = (1.0 - a1) * a2; wx = 1.0 - wy - wz; }
                            if (wx > 0.) { film_put(film, fp, cc.posx, cc.valx * (float) wx); if (cc.xl) splat_list_put(M, lane, 0, cc.xl, (float) wx); }
                            if (wy > 0.) { if (cc.yn) film_put(film, fp, cc.ypos, cc.yval * (float) wy); if (cc.yl) splat_list_put(M, lane, 1, cc.yl, (float) wy); }
                            if (wz > 0.) { if (cc.zn) film_put(film, fp, cc.zpos, cc.zval * (float) wz); if (cc.zl) splat_list_put(M, lane, 2, cc.zl, (float) wz); }
                        }
                        // the stage-2 kernels of the commit see the stage-2 context
                        MutCtx mc2 = mc;
                        mc2.lightTracing = cp.fixEmitterPath && c.tx == 1;
                        // ---- accept / reject, statistics (drmlt_proc.cpp:691-769; mixture :335-378)
                        if (cp.useMixture) {
                            const bool accept = did2 ? acc2 : acc1;
                            ++st[ST_ACC_B];
                            if (!did2) { ++st[ST_FIRST_B]; if (largeStep) ++st[ST_LARGE_B]; else ++st[ST_BOLD_B]; } else ++st[ST_SECOND_B];
                            if (accept) {
                                ++st[ST_ACC_A];
                                if (!did2) { ++st[ST_FIRST_A]; if (largeStep) ++st[ST_LARGE_A]; else ++st[ST_BOLD_A]; } else ++st[ST_SECOND_A];
                                if (did2) {
                                    commit_state(M, mc2, ub, dims, UB_Z, false, cc.zs, cc.zt, true);
                                    cc.Lx = cc.zL; cc.posx = cc.zpos; cc.valx = cc.zval; c.tx = cc.zt; cc.xl = cc.zl; if (cc.zl) splat_list_copy(M, lane, 0, 2, cc.zl, 1.f);
                                } else {
                                    commit_state(M, mc, ub, dims, UB_Y, true, cc.ys, cc.yt, true);
                                    cc.Lx = cc.yL; cc.posx = cc.ypos; cc.valx = cc.yval; c.tx = cc.yt; cc.xl = cc.yl; if (cc.yl) splat_list_copy(M, lane, 0, 1, cc.yl, 1.f);
                                }
                            }
                        } else if (acc1 || acc2) {
                            // acceptance map: binned at the state that is being LEFT (proposed.* after the swap, :695-708)
                            if (film && cp.acceptanceMap && (acc1 ? !largeStep : true))
                                film_put(film, fp, cc.posx, acc1 ? make_float3(1.f, 0.f, 0.f) : make_float3(0.f, 1.f, 0.f));
                            if (acc1) {
                                commit_state(M, mc, ub, dims, UB_Y, true, cc.ys, cc.yt, true);
                                cc.Lx = cc.yL; cc.posx = cc.ypos; cc.valx = cc.yval; c.tx = cc.yt; cc.xl = cc.yl; if (cc.yl) splat_list_copy(M, lane, 0, 1, cc.yl, 1.f);
                            } else {
                                commit_state(M, mc2, ub, dims, UB_Z, false, cc.zs, cc.zt, true);
                                cc.Lx = cc.zL; cc.posx = cc.zpos; cc.valx = cc.zval; c.tx = cc.zt; cc.xl = cc.zl; if (cc.zl) splat_list_copy(M, lane, 0, 2, cc.zl, 1.f);
                            }
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
                        ++c.mut;
                        c.phase = PH_STAGE1; c.large = 2u;
                        c.pstate = c.mut < mut_target(job, c.depth) ? PS_START : PS_IDLE;
                    }
                    if (c.pstate == PS_IDLE && job.chainCursor) {
                        // work-unit queue: this chain is complete; the lane takes the next one (generateWork, drmlt_proc.cpp:869-883)
                        if (!drmlt) {                              // PSSMLT's last splat of the accumulated state (pssmlt_proc.cpp:274-279)
                            const float3 v = cc.valx * (float) cc.cumW;
                            if (film && !is_zero(v)) film_put(film, fp, cc.posx, v);
                            if (film && cc.xl && cc.cumW > 0.) splat_list_put(M, lane, 0, cc.xl, (float) cc.cumW);
                            cc.cumW = 0.;
                        }
                        const unsigned int idx = atomicAdd(job.chainCursor, 1u);
                        if (idx < job.chainEnd) chain_init(M, lane, c, job.qDepth, job.qChainId, job.qSeedIdx, idx);
                    }
                    rec_store(M.lm.chain + lane, cc);
                }
                if (c.pstate == PS_IDLE) break;
            }
            break;
        }
        // ================= start the lane's next path: mutate the proposal, emit its first ray (begin.cuh) =================
        if (FUSED && c.pstate == PS_START) {
            RayF ray;
            const int dest = begin_path(M, lane, c, ray);
            rec_store(M.lm.core + lane, c);
            q_push_ray(M.q, dest, (uint32_t) lane, ray);
            continue;
        }
        rec_store(M.lm.core + lane, c);
        if (c.pstate == PS_START)
            q_push(M.q, Q_BEGIN + (job.type != JOB_CHAIN ? BEGIN_OTHER : (c.phase == PH_STAGE1 ? BEGIN_STAGE1 : (c.phase == PH_STAGE2 ? BEGIN_STAGE2 : BEGIN_OTHER))), (uint32_t) lane);
    }
    stats_flush(st, M.counters);
}

// ------------------------------------------------------------------ lane set-up
// JOB_CHAIN: seed replay + fillReplay (drmlt_proc.cpp:467-504): current = the seed's bootstrap vector; the lane then
// evaluates it (PH_INIT) before its first mutation.  JOB_BOOT / JOB_EVAL: lane l starts at item l.
__global__ void k_setup_lanes(const __grid_constant__ Machine M, const int *depthIn, const unsigned long long *chainIdIn, const unsigned long long *seedIdxIn) {
    const int lane = M.laneBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= M.laneEnd) return;
    Core c;
    memset(&c, 0, sizeof(c));
    c.tx = -1; c.large = 2u; c.phase = PH_STAGE1;
    bool queued;
    if (M.job.type == JOB_CHAIN) {
        // resident chains: chain = lane; work-unit queue: the lanes start with the first chains of the range
        const size_t idx = M.job.chainCursor ? (size_t) M.job.chainFirst + lane : (size_t) lane;
        queued = !M.job.chainCursor || idx < M.job.chainEnd;
        if (queued) chain_init(M, lane, c, depthIn, chainIdIn, seedIdxIn, idx);
        else c.pstate = PS_IDLE;
    } else {
        queued = lane < M.job.nItems;
        c.pstate = queued ? PS_START : PS_IDLE;
    }
    rec_store(M.lm.core + lane, c);
    if (queued) q_push(M.q, Q_CHAIN + M.parity, (uint32_t) lane);
}

// dr_job_run raised the mutation target: idle chains start their next mutation
__global__ void k_resume_lanes(const __grid_constant__ Machine M) {
    const int lane = M.laneBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= M.laneEnd) return;
    Core *c = M.lm.core + lane;
    if (c->pstate == PS_IDLE && c->mut < mut_target(M.job, c->depth)) {
        c->pstate = PS_START; c->phase = PH_STAGE1; c->large = 2u;
        q_push(M.q, Q_CHAIN + M.parity, (uint32_t) lane);
    }
}

// PSSMLT's "last splat" of the accumulated current state (pssmlt_proc.cpp:274-279); resets the weight
__global__ void k_flush_pssmlt(const __grid_constant__ Machine M) {
    const int lane = M.laneBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= M.laneEnd) return;
    ChainCore cc;
    rec_load(cc, M.lm.chain + lane);
    const float3 v = cc.valx * (float) cc.cumW;
    if (!is_zero(v)) film_put(M.film, M.fp, cc.posx, v);
    if (cc.xl && cc.cumW > 0.) splat_list_put(M, lane, 0, cc.xl, (float) cc.cumW);
    cc.cumW = 0.;
    rec_store(M.lm.chain + lane, cc);
}

// dr_splat_points: the film alone -- ImageBlock::put of explicit (position, RGB) pairs (parity against the reference's own block)
__global__ void k_splat_points(const __grid_constant__ FilmParams fp, float4 *film, const float2 *pos, const float3 *rgb, long long n) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i < n) film_put(film, fp, pos[i], rgb[i]);
}
void launch_splat_points(const FilmParams &fp, float4 *film, const float *pos, const float *rgb, long long n, cudaStream_t s) {
    k_splat_points<<<(unsigned) ((n + 127) / 128), 128, 0, s>>>(fp, film, reinterpret_cast<const float2 *>(pos), reinterpret_cast<const float3 *>(rgb), n);
}

static unsigned grid_for(int n, int threads) { return stage_grid(n, threads); }

template <int CLS>
DR_D void begin_lane(const Machine &M, uint32_t qi) {
    const int lane = (int) M.q.items[(size_t) (Q_BEGIN + CLS) * M.q.n + qi];
    Core c;
    rec_load(c, M.lm.core + lane);
    if (CLS == BEGIN_STAGE1) c.phase = PH_STAGE1;             // compile-time phase for the proposal switch
    if (CLS == BEGIN_STAGE2) c.phase = PH_STAGE2;
    RayF ray;
    const int dest = begin_path(M, lane, c, ray);
    rec_store(M.lm.core + lane, c);
    q_push_ray(M.q, dest, (uint32_t) lane, ray);
}
__global__ void __launch_bounds__(128, BEGIN_MINB)
k_begin(const __grid_constant__ Machine M) {
    queues_recycle(M.q);
    const uint32_t cnt[3] = { M.q.count[Q_BEGIN + BEGIN_STAGE1], M.q.count[Q_BEGIN + BEGIN_STAGE2], M.q.count[Q_BEGIN + BEGIN_OTHER] };
    const uint32_t nWarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;; w += nWarps) {
        int cls; uint32_t qi;
        if (!multiq_locate<3>(cnt, w, cls, qi)) break;
        if (qi >= cnt[cls]) continue;
        if (cls == BEGIN_STAGE1) begin_lane<BEGIN_STAGE1>(M, qi);
        else if (cls == BEGIN_STAGE2) begin_lane<BEGIN_STAGE2>(M, qi);
        else begin_lane<BEGIN_OTHER>(M, qi);
    }
}

// groups of up to this many lanes start the next path inside k_chain (DRMLT_FUSE_BEGIN_LANES overrides; 0 = never)
static int fuse_begin_lanes() {
    static const int v = getenv("DRMLT_FUSE_BEGIN_LANES") ? atoi(getenv("DRMLT_FUSE_BEGIN_LANES")) : 131072;
    return v;
}
bool chain_begin_fused(int nLanes, int integrator) { return integrator == DR_INTEGRATOR_PSSMLT && nLanes <= fuse_begin_lanes(); }
void launch_chain(const Machine &M, const LaunchCfg &lc) {
    if (chain_begin_fused(lc.nLanes, M.pp.integrator)) { k_chain<true><<<grid_for(lc.nLanes, 128), 128, 0, lc.stream>>>(M); return; }
    k_chain<false><<<grid_for(lc.nLanes, 128), 128, 0, lc.stream>>>(M);
    k_begin<<<grid_for(lc.nLanes + 3 * 128, 128), 128, 0, lc.stream>>>(M);
}
void launch_setup(const Machine &M, const LaunchCfg &lc, const int *depth, const unsigned long long *chainId, const unsigned long long *seedIdx) {
    k_setup_lanes<<<(unsigned) ((lc.nLanes + 127) / 128), 128, 0, lc.stream>>>(M, depth, chainId, seedIdx);
}
void launch_resume(const Machine &M, const LaunchCfg &lc) { k_resume_lanes<<<(unsigned) ((lc.nLanes + 127) / 128), 128, 0, lc.stream>>>(M); }
void launch_flush_pssmlt(const Machine &M, const LaunchCfg &lc) { k_flush_pssmlt<<<(unsigned) ((lc.nLanes + 127) / 128), 128, 0, lc.stream>>>(M); }

#ifdef DR_FILM_MATCH_STATS
void film_match_stats_print() {
    unsigned long long h[3] = { 0, 0, 0 };
    cudaMemcpyFromSymbol(h, g_filmMatch, sizeof(h));
    fprintf(stderr, "[film_put] %llu splats from k_chain, %llu (%.4f %%) share their filter footprint with another thread splatting at the same time; "
            "%.1f of 32 threads active per splat\n", h[0], h[1], h[0] ? 100.0 * (double) h[1] / (double) h[0] : 0.0, h[0] ? (double) h[2] / (double) h[0] : 0.0);
}
#endif
