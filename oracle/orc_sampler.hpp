// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
//
// orc_sampler.hpp: PathSampler::sampleSplats (src/libbidir/pathsampler.cpp:79-571).
#pragma once
#include "orc_path.hpp"

namespace orc {

struct PathSampler {
    PathCtx ctx;
    PathSamplerConfig cfg;
    Sampler *emitterSampler, *sensorSampler, *directSampler;
    int emitterDepth, sensorDepth;
    Path emitterSubpath, sensorSubpath;

    PathSampler(const Scene *scene, const PathSamplerConfig &c, Sampler *em, Sampler *se, Sampler *di)
        : cfg(c), emitterSampler(em), sensorSampler(se), directSampler(di) {
        ctx.scene = scene;
        // pathsampler.cpp:53-71: pinhole = degenerate sensor (no extra emitter step),
        // area lights are intersectable (one extra sensor step)
        emitterDepth = sensorDepth = cfg.maxDepth;
        if (sensorDepth != -1) ++sensorDepth;
    }

    void sampleSplats(SplatList &list, int depth) {
        const Scene &sc = *ctx.scene;
        list.clear();
        list.s = list.t = -1;
        switch (cfg.technique) {
        case DR_TECH_MMLT: sampleMMLT(list, depth); break;
        case DR_TECH_BDPT: sampleBDPT(list); break;
        case DR_TECH_PATH: {   // pathsampler.cpp:529-567 (pinhole: no aperture/time sample)
            Vec2 samplePos = sensorSampler->next2D();
            samplePos.x *= sc.cam.resX; samplePos.y *= sc.cam.resY;
            // perspective.cpp:271-298 sampleRayDifferential
            Vec3 dl = sc.cam.sampleToDir(samplePos.x / sc.cam.resX, samplePos.y / sc.cam.resY);
            Float invZ = 1.0 / dl.z;
            Ray ray; ray.o = sc.cam.pos; ray.d = sc.cam.xformDir(dl);
            ray.mint = sc.cam.nearClip * invZ; ray.maxt = sc.cam.farClip * invZ;
            RGB value = pathTracerLi(ctx, sensorSampler, ray, cfg.maxDepth, cfg.rrDepth, cfg.excludeDirectIllum);
            list.append(samplePos, value);
            break;
        }
        }
    }

    // pathsampler.cpp:84-320
    void sampleMMLT(SplatList &list, int depth) {
        const Scene &sc = *ctx.scene;
        emitterSubpath.initialize(EImportance);
        sensorSubpath.initialize(ERadiance);
        int s, t, nStrats;
        Float random_decision = directSampler->next1D();
        if (cfg.lightImage) {
            nStrats = depth + 1;
            s = std::min(int(nStrats * random_decision), nStrats - 1);
            t = nStrats - s;
        } else {
            nStrats = depth;
            s = std::min(int(nStrats * random_decision), nStrats - 1);
            t = 1 + (nStrats - s);
        }
        list.s = s; list.t = t;
        if (depth == 1) return;

        int t_sampled = randomWalk(ctx, sensorSubpath, sensorSampler, t, t + 1, ERadiance);
        int s_sampled = randomWalk(ctx, emitterSubpath, emitterSampler, s, s + 1, EImportance);
        if (t_sampled != t) return;
        if (s_sampled != s) return;

        bool unconnectable = true;
        for (size_t i = 2; i < emitterSubpath.vertexCount(); ++i) unconnectable &= !emitterSubpath.v[i].isConnectable();
        for (size_t i = 2; i < sensorSubpath.vertexCount(); ++i) unconnectable &= !sensorSubpath.v[i].isConnectable();
        if (unconnectable) return;

        RGB weight(1.0);
        for (size_t i = 1; i < emitterSubpath.vertexCount(); ++i)
            weight *= emitterSubpath.v[i - 1].weight[EImportance] * emitterSubpath.v[i - 1].rrWeight * emitterSubpath.e[i - 1].weight[EImportance];
        for (size_t i = 1; i < sensorSubpath.vertexCount(); ++i)
            weight *= sensorSubpath.v[i - 1].weight[ERadiance] * sensorSubpath.v[i - 1].rrWeight * sensorSubpath.e[i - 1].weight[ERadiance];

        PathVertex *vsPred = emitterSubpath.vertexOrNull(s - 1), *vtPred = sensorSubpath.vertexOrNull(t - 1),
                   *vs = &emitterSubpath.v[s], *vt = &sensorSubpath.v[t];
        RGB value;
        Vec2 samplePos(0.0, 0.0);
        PathEdge connectionEdge;
        if (vs->isEmitterSupernode()) {
            if (!vertexCastEmitter(*vt) || vt->isDegenerate()) return;
            value = weight * vertexEval(sc, *vs, vsPred, vt, EImportance) * vertexEval(sc, *vt, vtPred, vs, ERadiance);
        } else if (vt->isSensorSupernode()) {
            if (!vertexCastSensor(*vs) || vs->isDegenerate()) return;
            if (!vertexGetSamplePosition(sc, *vs, *vsPred, samplePos)) return;
            value = weight * vertexEval(sc, *vs, vsPred, vt, EImportance) * vertexEval(sc, *vt, vtPred, vs, ERadiance);
        } else {
            if (vs->isDegenerate() || vt->isDegenerate()) return;
            value = weight * vertexEval(sc, *vs, vsPred, vt, EImportance) * vertexEval(sc, *vt, vtPred, vs, ERadiance);
            vs->measure = vt->measure = EArea;
        }
        int interactions = cfg.maxDepth - depth;
        if (value.isZero() || !pathConnectAndCollapse(ctx, connectionEdge, *vs, *vt, interactions)) return;
        if (cfg.excludeDirectIllum && depth <= 2) return;
        value *= edgeEvalCachedGG(connectionEdge, *vs, *vt);
        Float mis = miWeight(sc, emitterSubpath, &connectionEdge, sensorSubpath, s, t, cfg.lightImage);
        list.misWeight = mis;
        value *= mis;
        value *= (Float) nStrats;
        if (vt->isSensorSample() && !vertexGetSamplePosition(sc, *vt, *vs, samplePos)) return;
        if (t < 2) {
            list.append(samplePos, value);
        } else {
            Vec2 sensorSamplePos(0.0, 0.0);
            vertexGetSamplePosition(sc, sensorSubpath.v[1], sensorSubpath.v[2], sensorSamplePos);
            list.append(sensorSamplePos, value);
        }
    }

    // pathsampler.cpp:321-527 with m_sampleDirect = false
    void sampleBDPT(SplatList &list) {
        const Scene &sc = *ctx.scene;
        emitterSubpath.initialize(EImportance);
        sensorSubpath.initialize(ERadiance);
        randomWalk(ctx, emitterSubpath, emitterSampler, emitterDepth, cfg.rrDepth, EImportance);
        randomWalk(ctx, sensorSubpath, sensorSampler, sensorDepth, cfg.rrDepth, ERadiance);
        list.s = (int) emitterSubpath.vertexCount() - 1;   // diagnostic: index of the last vertex of each subpath
        list.t = (int) sensorSubpath.vertexCount() - 1;

        std::vector<RGB> importanceWeights(emitterSubpath.vertexCount()), radianceWeights(sensorSubpath.vertexCount());
        importanceWeights[0] = radianceWeights[0] = RGB(1.0);
        for (size_t i = 1; i < emitterSubpath.vertexCount(); ++i)
            importanceWeights[i] = importanceWeights[i - 1] * emitterSubpath.v[i - 1].weight[EImportance] *
                                   emitterSubpath.v[i - 1].rrWeight * emitterSubpath.e[i - 1].weight[EImportance];
        for (size_t i = 1; i < sensorSubpath.vertexCount(); ++i)
            radianceWeights[i] = radianceWeights[i - 1] * sensorSubpath.v[i - 1].weight[ERadiance] *
                                 sensorSubpath.v[i - 1].rrWeight * sensorSubpath.e[i - 1].weight[ERadiance];
        if (sensorSubpath.vertexCount() > 2) {
            Vec2 samplePos(0.0, 0.0);
            vertexGetSamplePosition(sc, sensorSubpath.v[1], sensorSubpath.v[2], samplePos);
            list.append(samplePos, RGB(0.0));
        }
        Vec2 samplePos(0.0, 0.0);
        for (int s = (int) emitterSubpath.vertexCount() - 1; s >= 0; --s) {
            int minT = std::max(2 - s, cfg.lightImage ? 0 : 2), maxT = (int) sensorSubpath.vertexCount() - 1;
            if (cfg.maxDepth != -1) maxT = std::min(maxT, cfg.maxDepth + 1 - s);
            for (int t = maxT; t >= minT; --t) {
                PathVertex *vsPred = emitterSubpath.vertexOrNull(s - 1), *vtPred = sensorSubpath.vertexOrNull(t - 1),
                           *vs = &emitterSubpath.v[s], *vt = &sensorSubpath.v[t];
                // RestoreMeasureHelper + cast() side effects: work on the stored vertices but restore afterwards
                PathVertex vsSaved = *vs, vtSaved = *vt;
                struct Restore { PathVertex *a, *b; PathVertex sa, sb; ~Restore() { a->measure = sa.measure; b->measure = sb.measure; } }
                    restore{ vs, vt, vsSaved, vtSaved };
                int depth = s + t - 1;
                int remaining = cfg.maxDepth - depth;
                RGB value;
                PathEdge connectionEdge;
                if (vs->isEmitterSupernode()) {
                    if (!vertexCastEmitter(*vt) || vt->isDegenerate()) continue;
                    value = radianceWeights[t] * vertexEval(sc, *vs, vsPred, vt, EImportance) * vertexEval(sc, *vt, vtPred, vs, ERadiance);
                } else if (vt->isSensorSupernode()) {
                    if (!vertexCastSensor(*vs) || vs->isDegenerate()) continue;
                    if (!vertexGetSamplePosition(sc, *vs, *vsPred, samplePos)) continue;
                    value = importanceWeights[s] * vertexEval(sc, *vs, vsPred, vt, EImportance) * vertexEval(sc, *vt, vtPred, vs, ERadiance);
                } else {
                    if (vs->isDegenerate() || vt->isDegenerate()) continue;
                    value = importanceWeights[s] * radianceWeights[t] * vertexEval(sc, *vs, vsPred, vt, EImportance) *
                            vertexEval(sc, *vt, vtPred, vs, ERadiance);
                    vs->measure = vt->measure = EArea;
                }
                int interactions = remaining;
                if (value.isZero() || !pathConnectAndCollapse(ctx, connectionEdge, *vs, *vt, interactions)) continue;
                depth += interactions;
                if (cfg.excludeDirectIllum && depth <= 2) continue;
                value *= edgeEvalCachedGG(connectionEdge, *vs, *vt);
                value *= miWeight(sc, emitterSubpath, &connectionEdge, sensorSubpath, s, t, cfg.lightImage);
                if (vt->isSensorSample() && !vertexGetSamplePosition(sc, *vt, *vs, samplePos)) continue;
                if (t < 2) list.append(samplePos, value);
                else list.accum(0, value);
            }
        }
    }
};

} // namespace orc
