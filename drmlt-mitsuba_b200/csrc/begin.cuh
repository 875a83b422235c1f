// begin.cuh -- start of a path (k_begin, k_chain.cu): the proposal is mutated into the lane's coordinate buffer (pss.cuh) and the
// first ray of its path is emitted.  k_begin runs one class of this work per warp, fed by its own queue (machine.cuh Q_BEGIN).
// (Fusing it into the chain kernel -- one Core round trip and one launch per round less -- was measured in round 2: the fused
// kernel needs 164 registers instead of 126 and runs the three proposal classes divergently, 178 us against 75 + 76 us per
// 120 k paths under ncu and 145 against 152 M mutations/s in the bench; profiles/r02_*.)
//   stage 1   first-stage proposal  (large step: uniforms | Kelemen | orbital: radial Kelemen + angle)
//   stage 2   second-stage proposal (Gaussian | orbital rotation by a wrapped-Cauchy angle)
//   other     seed replay, Green's reverse state, bootstrap samples, replayed vectors
// Behavioural parity: DRMLTSampler::fillSpace / OrbitalDRMLTSampler (src/integrators/drmlt/drmlt_sampler.cpp:313-394),
// PSSMLTSampler::primarySample (src/integrators/pssmlt/pssmlt_sampler.cpp:124-166), the first steps of
// PathSampler::sampleSplats (src/libbidir/pathsampler.cpp:84-146, 529-567).
#pragma once
#include "bdpt.cuh"


// ------------------------------------------------------------------ proposals
// Fill the lane's coordinate buffer for the path that is about to start and select it (Core::ubuf).
DR_D void fill_proposal(const Machine &M, int lane, Core &c, const MutCtx &mc, long long item) {
    const PssParams &pp = M.pp;
    const int nU = M.lm.nU;
    double *ub = M.lm.ubuf + (size_t) lane * M.lm.ubCount * nU;
    int dims[3];
    chain_dims(M.pc, M.cp, c.depth, dims);
    const bool mmlt = M.pc.technique == DR_TECH_MMLT;
    if (M.job.type == JOB_EVAL) {                  // replayed host vectors (float) -> X
        c.ubuf = UB_X;
        const float *src[3] = { M.job.us + item * M.job.ds, M.job.ue + item * M.job.de, M.job.ud + item * M.job.dd };
        const int n[3] = { M.job.ds, M.job.de, M.job.dd };
        for (int s = 0; s < 3; ++s)
            for (int k = 0; k < n[s]; ++k) ub[pp.off[s] + k] = (double) src[s][k];
        return;
    }
    if (M.job.type == JOB_BOOT) {                  // bootstrap sample `index`: keyed uniforms -> X
        c.ubuf = UB_X;
        const unsigned long long index = M.job.first + (unsigned long long) item;
        auto boot_pair = [&](int s, int p) {
            const float4 u = keyed_uniform4(pp.seed, S_BOOT, index, (uint32_t) s, (uint32_t) (p >> 1));
            return (p & 1) ? r2(u.z, u.w) : r2(u.x, u.y);
        };
        int ext[3] = { (dims[0] + 1) >> 1, (dims[1] + 1) >> 1, (dims[2] + 1) >> 1 };
        if (mmlt) {                                // only what strategy (s, t) can consume
            const R2 d = boot_pair(SMP_DIRECT, 0);
            ub_store(ub, pp.off[SMP_DIRECT], d);
            int s_, t_;
            mmlt_strategy(M.pc, c.depth, d.x, s_, t_);
            ext[0] = min(ext[0], subset_pairs(M, t_)); ext[1] = min(ext[1], subset_pairs(M, s_)); ext[2] = 0;
        }
        for (int s = 0; s < 3; ++s)
            for (int p = 0; p < ext[s]; ++p) ub_store(ub, pp.off[s] + 2 * p, boot_pair(s, p));
        return;
    }
    // ---- Markov chain
    if (c.phase == PH_INIT) { c.ubuf = UB_X; return; }        // seed replay: the bootstrap vector is already in X
    const int dst = c.phase == PH_STAGE1 ? UB_Y : (c.phase == PH_STAGE2 ? UB_Z : UB_R);
    c.ubuf = (uint8_t) dst;
    auto make_pair = [&](int s, int p) {
        const int slot = pp.off[s] + 2 * p;
        const R2 x = ub_load(ub + UB_X * nU, slot);
        R2 v;
        if (c.phase == PH_STAGE1) v = propose_stage1(mc, s, p, x);
        else {
            const R2 y = ub_load(ub + UB_Y * nU, slot);
            if (c.phase == PH_STAGE2) v = propose_stage2(mc, s, p, x, y);
            else { const R2 z = ub_load(ub + UB_Z * nU, slot); v = r2(z.x - (y.x - x.x), z.y - (y.y - x.y)); }   // y* = z - (y - x), drmlt_sampler.cpp:293-296
        }
        ub_store(ub + dst * nU, slot, v);
        return v;
    };
    int ext[3] = { (dims[0] + 1) >> 1, (dims[1] + 1) >> 1, (dims[2] + 1) >> 1 };
    if (pp.subset) {
        const R2 d = make_pair(SMP_DIRECT, 0);
        int s_, t_;
        mmlt_strategy(M.pc, c.depth, wrap_reflect(d.x), s_, t_);
        ext[0] = subset_pairs(M, t_); ext[1] = subset_pairs(M, s_); ext[2] = 0;
    }
    for (int s = 0; s < 3; ++s)
        for (int p = 0; p < ext[s]; ++p) make_pair(s, p);
}

// ------------------------------------------------------------------ path start
// returns the queue the lane goes to (Q_RAYC, Q_CONNECT) or -1 when the path is already over (empty result)
DR_D int path_start(const Machine &M, int lane, Core &c, RayF &ray) {
    const DevScene &sc = M.sc;
    UReader rd;
    c.pos0 = c.pos1 = c.pos2 = 0; c.nrays = 0;
    reader_open(M, c, lane, rd);
    int dest;
    if (M.pc.technique == DR_TECH_MMLT) {                     // pathsampler.cpp:84-159
        const int depth = c.depth, k = depth + 2;
        int s, t;
        mmlt_strategy(M.pc, depth, rd.next1D(SMP_DIRECT), s, t);
        c.s = (uint8_t) s; c.t = (uint8_t) t;
        if (depth == 1) { reader_close(rd, c); c.pstate = PS_EMPTY; return -1; }
        c.connectable = 0; c.flags = 0; c.weight = r3(1.);
        if (!mmlt_emitter_sample(M, lane, c, rd)) { reader_close(rd, c); c.pstate = PS_EMPTY; return -1; }
        (void) rd.next2D(SMP_SENSOR);                         // sampleSensorPosition consumes 2 (vertex.cpp:79)
        // (pdfRad of the sensor supernode and of the sensor sample are 1: perspective.cpp:305; path_result knows)
        Vtx vt;
        vt.p = cam_pos(sc.cam); vt.ng = vt.ns = cam_dir(sc.cam); vt.ss = r3(0.); vt.type = V_SENSOR_SAMPLE; vt.degenerate = 0; vt.mat = -1; vt.emitter = -1;
        c.connectable |= 1u << (k - 1);                       // sensor sample: never discrete, not degenerate
        rec_store(M.lm.vt + lane, vt);
        if (t >= 2) {
            PredRec g1;
            g1.p = vt.p; g1.ng = vt.ng; g1.pad[0] = g1.pad[1] = 0.;
            rec_store(geo_slot(M, lane, SIDE_S, 1), g1);
        }
        c.j = 1;
        if (t >= 2) {                                         // vertex.cpp:126-151, perspective.cpp:318-345
            const R2 u = rd.next2D(SMP_SENSOR);
            const R3 dl = cam_sample_to_dir(sc.cam, u.x, u.y);
            c.pdfFwd = sc.cam.normalization / (dl.z * dl.z * dl.z);
            c.pdfBwd = 1.0;
            c.pstate = PS_SENSOR_HIT;
            emit_ray(M, lane, c, vt.p, cam_xform_dir(sc.cam, dl), sc.epsilon, INFINITY, ray);
            dest = Q_RAYC;
        } else {
            dest = mmlt_emitter_launch(M, lane, c, ray);
        }
    } else if (M.pc.technique == DR_TECH_BDPT) {              // pathsampler.cpp:321-341
        dest = bdpt_path_start(M, lane, c, rd, ray);
    } else {                                                  // PathSampler EUnidirectional (pathsampler.cpp:529-567)
        const R2 u0 = rd.next2D(SMP_SENSOR);
        const R2 samplePos = r2(u0.x * sc.cam.resX, u0.y * sc.cam.resY);
        c.spos = make_float2((float) samplePos.x, (float) samplePos.y);
        const R3 dl = cam_sample_to_dir(sc.cam, samplePos.x / sc.cam.resX, samplePos.y / sc.cam.resY);
        const Real invZ = 1.0 / dl.z;
        PtExtra px;
        px.Li = r3(0.); px.pending = r3(0.); px.refN = r3(0.); px.eta = 1.0; px.bsPdf = 0.; px.dIn = r3(0.);
        c.weight = r3(1.);
        c.flags = F_PT_FIRST | (M.pc.excludeDirect ? 0u : (F_PT_EMITTED | F_PT_DIRECT));
        c.j = 1; c.s = c.t = 0;
        c.pstate = PS_PT_HIT;
        rec_store(reinterpret_cast<PtExtra *>(M.lm.vs + lane), px);
        emit_ray(M, lane, c, cam_pos(sc.cam), cam_xform_dir(sc.cam, dl), sc.cam.nearClip * invZ, sc.cam.farClip * invZ, ray);
        dest = Q_RAYC;
    }
    reader_close(rd, c);
    return dest;
}



// Start the next path of a lane whose state says PS_START: returns the queue the lane goes to with `ray` (a ray queue of the
// next round, or the next round's chain queue for a path that is over before its first ray -- MMLT depth 1, no emitter).
DR_D int begin_path(const Machine &M, int lane, Core &c, RayF &ray) {
    const JobParams &job = M.job;
    long long item = 0;
    MutCtx mc;
    mc.pp = &M.pp; mc.chain = c.chainId; mc.mut = c.mut; mc.largeStep = false; mc.lightTracing = false; mc.table = replay_table(M, lane);
    if (job.type == JOB_CHAIN) {
        if (c.phase == PH_STAGE1 && c.large == 2u)                // new mutation: draw the large-step coin (drmlt_proc.cpp:533)
            c.large = chain_coin(M, lane, c, 0) < M.cp.pLarge ? 1u : 0u;
        mc.largeStep = c.phase != PH_INIT && c.large == 1u;
        mc.lightTracing = c.phase == PH_STAGE2 && M.cp.fixEmitterPath && c.tx == 1;   // nextStage(current->t == 1)
    } else {
        item = (long long) lane + (long long) c.mut * M.lm.n;
        if (job.type == JOB_BOOT) {
            const unsigned long long index = job.first + (unsigned long long) item;
            c.depth = M.pc.technique == DR_TECH_MMLT ? (uint8_t) ((index % (unsigned long long) M.pc.maxDepth) + 1) : 0;
        } else c.depth = job.depthIn ? (uint8_t) job.depthIn[item] : 0;
    }
    fill_proposal(M, lane, c, mc, item);
    const int dest = path_start(M, lane, c, ray);
    if (dest != Q_RAYC) c.pstate = PS_EMPTY;                      // (a connection without any ray cannot occur for depth >= 2)
    return dest == Q_RAYC ? Q_RAYC + (M.parity ^ 1) : Q_CHAIN + (M.parity ^ 1);
}
