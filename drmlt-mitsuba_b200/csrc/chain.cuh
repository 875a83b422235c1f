// chain.cuh -- the Markov-chain step: mutation, path evaluation, two-stage delayed-rejection
// acceptance and film splatting fused in one kernel; plus bootstrap, seed resampling and develop.
//
// One thread owns one chain and runs a small state machine whose every iteration evaluates ONE
// path (stage-1 proposal, stage-2 proposal, or Green's reverse path).  Threads of a warp therefore
// never idle while a neighbour runs its second stage: a chain that needs no second stage simply
// starts its next mutation.  Behavioural parity targets:
//   DRMLTRenderer::process / processMixture  src/integrators/drmlt/drmlt_proc.cpp:161-380, 386-771
//   PSSMLTRenderer::process                  src/integrators/pssmlt/pssmlt_proc.cpp:110-285
//   PathSampler::generateSeeds               src/libbidir/pathsampler.cpp:859-960
//   ImageBlock::put                          include/mitsuba/render/imageblock.h:149-196
//   DRMLTProcess::develop                    src/integrators/drmlt/drmlt_proc.cpp:813-854
#pragma once
#include "path.cuh"

struct FilmParams {
    int w, h;
    float radius, scaleFactor;
    float values[32];          // rfilter.cpp:37-55 discretised filter (MTS_FILTER_RESOLUTION = 31)
};

struct ChainParams {
    float pLarge;
    float b;                   // m_config.luminance
    int acceptanceMap, timidAfterLarge, fixEmitterPath, useMixture, kelemenWeights;
    float kel_s1, kel_s2, kel_logRatio;     // un-scaled Kelemen bounds for Mira's transition ratio
};

struct ChainArrays {
    float *X;                  // [totalDim][n] current primary-sample vectors (SoA)
    float *L;                  // [n] current luminance
    float2 *pos;               // [n] current splat position
    float4 *val;               // [n] current normalised splat RGB, .w = PSSMLT cumulative weight
    int *tcur;                 // [n] current MMLT strategy t (fixEmitterPath)
    int *depth;                // [n] MMLT depth (or -1)
    unsigned long long *chainId;   // [n] RNG key of the chain
    unsigned long long *seedIdx;   // [n] bootstrap sample the chain starts from
    unsigned int *mutDone;     // [n] mutations already performed
    int n;
    int dimS, dimE, dimD;      // allocated coordinates per sampler (maxDepth worst case)
};

enum { ST_MUT = 0, ST_FIRST_A, ST_FIRST_B, ST_LARGE_A, ST_LARGE_B, ST_BOLD_A, ST_BOLD_B, ST_SECOND_A, ST_SECOND_B,
       ST_SECOND_LARGE_A, ST_SECOND_LARGE_B, ST_SECOND_BOLD_A, ST_SECOND_BOLD_B, ST_ACC_A, ST_ACC_B, ST_PATHS, ST_RAYS, ST_COUNT };

// ---------------------------------------------------------------- film
// Warp-aggregated splat: lanes that hit the same pixel with the same footprint origin are merged
// by __match_any_sync before the (vector) atomic reaches L2.
DR_D void film_put(float4 *film, const FilmParams &fp, float2 pos, float3 value) {
    if (!rgb_valid(value)) return;
    const float px = pos.x - 0.5f, py = pos.y - 0.5f;
    const int minx = max((int) ceilf(px - fp.radius), 0), miny = max((int) ceilf(py - fp.radius), 0);
    const int maxx = min((int) floorf(px + fp.radius), fp.w - 1), maxy = min((int) floorf(py + fp.radius), fp.h - 1);
    for (int y = miny; y <= maxy; ++y) {
        const float wy = fp.values[min((int) fabsf((y - py) * fp.scaleFactor), 31)];
        for (int x = minx; x <= maxx; ++x) {
            const float w = fp.values[min((int) fabsf((x - px) * fp.scaleFactor), 31)] * wy;
            if (w == 0.f) continue;
            float4 *dst = film + (size_t) y * fp.w + x;
#if __CUDA_ARCH__ >= 900
            atomicAdd(dst, make_float4(w * value.x, w * value.y, w * value.z, 0.f));
#else
            atomicAdd(&dst->x, w * value.x); atomicAdd(&dst->y, w * value.y); atomicAdd(&dst->z, w * value.z);
#endif
        }
    }
}

DR_D void evaluate(const DevScene &sc, const PathCfg &pc, Pss &pss, int depth, PathResult &r, uint32_t &rays) {
    if (pc.technique == DR_TECH_MMLT) eval_mmlt(sc, pc, pss, depth, r, rays);
    else eval_pt(sc, pc, pss, r, rays);
}

DR_D void chain_dims(const PathCfg &pc, int depth, int dimS, int dimE, int dimD, int dims[3]) {
    // findMaxDimensions (pssmlt_utils.h:27-77): MMLT vectors depend on the chain's depth
    if (pc.technique == DR_TECH_MMLT) {
        int m = (depth + 2) * 3; if (m & 1) m++;
        dims[0] = m; dims[1] = m; dims[2] = 1;
    } else { dims[0] = dimS; dims[1] = dimE; dims[2] = dimD; }
}

DR_D bool invalid_strict(float x) { return isnan(x) || isinf(x) || x <= 0.f; }   // drmlt_proc.cpp:428
DR_D bool invalid_loose(float x) { return isnan(x) || isinf(x) || x < 0.f; }     // drmlt_proc.cpp:181
DR_D float metropolis_clamp(float x) { return x < 1.0f ? x : 1.0f; }             // std::min(1, x): NaN -> 1

// ---------------------------------------------------------------- bootstrap (generateSeeds body)
__global__ void __launch_bounds__(128)
k_bootstrap(const __grid_constant__ DevScene sc, const __grid_constant__ PathCfg pc, const __grid_constant__ PssParams pp, unsigned long long first, long long n, float *lum, unsigned long long *counters) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    uint32_t rays = 0;
    if (i < n) {
        Pss pss;
        pss.pp = &pp; pss.xs[0] = pss.xs[1] = pss.xs[2] = nullptr; pss.stride = 0;
        const unsigned long long index = first + (unsigned long long) i;
        const int depth = pc.technique == DR_TECH_MMLT ? (int) (index % (unsigned long long) pc.maxDepth) + 1 : -1;
        int dims[3];
        chain_dims(pc, depth < 0 ? pc.maxDepth : depth, 1 << 20, 1 << 20, 1 << 20, dims);
        pss.dim[0] = dims[0]; pss.dim[1] = dims[1]; pss.dim[2] = dims[2];
        pss.chain = index; pss.mut = 0; pss.largeStep = false; pss.lightTracing = false;
        pss.begin(PSS_BOOT);
        PathResult r;
        evaluate(sc, pc, pss, depth, r, rays);
        lum[i] = r.lum;
    }
    // rays -> counters[1], paths -> counters[0]
    unsigned int total = rays;
    for (int o = 16; o > 0; o >>= 1) total += __shfl_down_sync(0xffffffffu, total, o);
    if ((threadIdx.x & 31) == 0 && total) atomicAdd(&counters[1], (unsigned long long) total);
}

// ---------------------------------------------------------------- b and the luminance CDF
// sum / count of the non-NaN luminances: warp shuffle reduce -> one atomic per warp (double)
__global__ void k_lum_reduce(const float *lum, long long n, double *out /* [2]: sum, count */) {
    double s = 0.0, c = 0.0;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const float v = lum[i];
        if (!isnan(v)) { s += (double) v; c += 1.0; }
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); c += __shfl_down_sync(0xffffffffu, c, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], s); atomicAdd(&out[1], c); }
}

// Inclusive scan of max(lum,0) in double, three passes (block scan, scan of block totals, add).
#define SCAN_BLOCK 1024
#define SCAN_ITEMS 4
__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_blocks(const float *lum, long long n, double *cdf /* n+1, cdf[0] = 0 */, double *blockSums) {
    __shared__ double warpSums[32];
    const long long base = (long long) blockIdx.x * SCAN_BLOCK * SCAN_ITEMS + (long long) threadIdx.x * SCAN_ITEMS;
    double v[SCAN_ITEMS], run = 0.0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + k;
        float x = i < n ? lum[i] : 0.f;
        if (!(x > 0.f) || isinf(x)) x = 0.f;     // NaN / negative / inf samples carry no seed
        run += (double) x; v[k] = run;
    }
    double incl = run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warpSums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double w = warpSums[lane];
        for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        warpSums[lane] = w;
    }
    __syncthreads();
    const double offset = (incl - run) + (warp ? warpSums[warp - 1] : 0.0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { const long long i = base + k; if (i < n) cdf[i + 1] = v[k] + offset; }
    if (threadIdx.x == SCAN_BLOCK - 1) blockSums[blockIdx.x] = offset + run;
    if (blockIdx.x == 0 && threadIdx.x == 0) cdf[0] = 0.0;
}
__global__ void k_scan_sums(double *blockSums, int nb) {   // single block, serial over <= a few thousand entries per lane chunk
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    for (int start = 0; start < nb; start += blockDim.x) {
        const int i = start + threadIdx.x;
        double v = i < nb ? blockSums[i] : 0.0;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        __shared__ double ws[32];
        double incl = v;
        for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            double w = lane < (blockDim.x >> 5) ? ws[lane] : 0.0;
            for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            ws[lane] = w;
        }
        __syncthreads();
        const double excl = incl - v + (warp ? ws[warp - 1] : 0.0) + carry;
        if (i < nb) blockSums[i] = excl;           // exclusive prefix of block totals
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
}
__global__ void k_scan_add(double *cdf, long long n, const double *blockSums) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i < n) cdf[i + 1] += blockSums[i / (SCAN_BLOCK * SCAN_ITEMS)];
}

// seedPDF.sample(next1D()) (pathsampler.cpp:946-954) on the un-normalised CDF
__global__ void k_resample(const double *cdf, long long n, unsigned long long seed, unsigned long long firstChain, int nChains,
                           unsigned long long bootFirst, int maxDepth, int technique,
                           unsigned long long *seedIdx, unsigned long long *chainId, int *depth) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nChains) return;
    const unsigned long long id = firstChain + (unsigned long long) c;
    const double total = cdf[n];
    const double v = (double) keyed_uniform(seed, S_RESAMPLE, id, 0u, 0u) * total;
    long long lo = 0, hi = n + 1;
    while (lo < hi) { long long mid = (lo + hi) >> 1; if (cdf[mid] < v) lo = mid + 1; else hi = mid; }
    long long index = min(n - 1, max(0ll, lo - 1));
    while (index < n - 1 && cdf[index + 1] - cdf[index] == 0.0) ++index;
    const unsigned long long s = bootFirst + (unsigned long long) index;
    seedIdx[c] = s; chainId[c] = id;
    depth[c] = technique == DR_TECH_MMLT ? (int) (s % (unsigned long long) maxDepth) + 1 : -1;
}

// ---------------------------------------------------------------- chain initialisation (seed replay)
__global__ void __launch_bounds__(128)
k_init_chains(const __grid_constant__ DevScene sc, const __grid_constant__ PathCfg pc, const __grid_constant__ PssParams pp, const __grid_constant__ ChainArrays ca, unsigned long long *counters) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ca.n) return;
    const unsigned long long sidx = ca.seedIdx[j];
    const int depth = ca.depth[j];
    int dims[3];
    chain_dims(pc, depth, ca.dimS, ca.dimE, ca.dimD, dims);
    // seed replay + fillReplay (drmlt_proc.cpp:467-504): current = the seed's bootstrap vector
    const int offs[3] = { 0, ca.dimS, ca.dimS + ca.dimE }, alloc[3] = { ca.dimS, ca.dimE, ca.dimD };
    for (int s = 0; s < 3; ++s)
        for (int k = 0; k < alloc[s]; ++k)
            ca.X[(size_t) (offs[s] + k) * ca.n + j] = keyed_uniform(pp.seed, S_BOOT, sidx, (uint32_t) s, (uint32_t) k);
    Pss pss;
    pss.pp = &pp; pss.stride = ca.n;
    for (int s = 0; s < 3; ++s) { pss.xs[s] = ca.X + (size_t) offs[s] * ca.n + j; pss.dim[s] = dims[s]; }
    pss.chain = sidx; pss.mut = 0; pss.largeStep = false; pss.lightTracing = false;
    pss.begin(PSS_ARRAY);
    PathResult r;
    uint32_t rays = 0;
    evaluate(sc, pc, pss, depth, r, rays);
    const float inv = r.lum > 0.f ? 1.0f / r.lum : 1.0f;      // SplatList::normalize
    ca.L[j] = r.lum;
    ca.pos[j] = r.pos;
    ca.val[j] = make_float4(r.val.x * inv, r.val.y * inv, r.val.z * inv, 0.f);
    ca.tcur[j] = r.t;
    ca.mutDone[j] = 0u;
    atomicAdd(&counters[ST_PATHS], 1ull);
    atomicAdd(&counters[ST_RAYS], (unsigned long long) rays);
}

// ---------------------------------------------------------------- the fused chain step
struct Proposal { float L; float2 pos; float3 val; int n; int t; };

DR_D void set_proposal(Proposal &p, const PathResult &r) {
    p.L = r.lum; p.pos = r.pos; p.n = r.n; p.t = r.t;
    const float inv = r.lum > 0.f ? 1.0f / r.lum : 1.0f;
    p.val = r.val * inv;
}

// MiraDRMLTSampler::getTransitionRatio over the three samplers (drmlt_sampler.cpp:400-414)
DR_D float mira_transition_ratio(const Pss &pss, const ChainParams &cp, const int maxIdx1[3], const int maxIdx2[3]) {
    float num = 0.f, den = 0.f;
    for (int s = 0; s < 3; ++s) {
        if (pss.identity1(s)) continue;
        const int dimStage = max(maxIdx1[s], maxIdx2[s]);
        for (int i = 0; i < dimStage; i += 2) {
            const float2 y = pss.prop1(s, i >> 1), z = pss.prop2(s, i >> 1);
            num += kelemen_logpdf(z.x - y.x, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
            den += kelemen_logpdf(pss.xat(s, i) - y.x, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
            if (i + 1 < dimStage) {
                num += kelemen_logpdf(z.y - y.y, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
                den += kelemen_logpdf(pss.xat(s, i + 1) - y.y, cp.kel_s1, cp.kel_s2, cp.kel_logRatio);
            }
        }
    }
    return expf(num - den);
}

// write the accepted proposal back as the new current state (DRMLTSampler::accept, drmlt_sampler.cpp:189-199)
DR_D void commit_state(const Pss &pss, float *Xbase, size_t n, const int offs[3], int slot, bool first, bool drmlt) {
    for (int s = 0; s < 3; ++s) {
        if (!pss.largeStep && (first ? pss.identity1(s) : pss.identity2(s))) continue;
        float *xs = Xbase + (size_t) offs[s] * n + slot;
        for (int k = 0; k < pss.dim[s]; k += 2) {
            float2 v = first ? pss.prop1(s, k >> 1) : pss.prop2(s, k >> 1);
            if (drmlt) { v.x = wrap_reflect(v.x); v.y = wrap_reflect(v.y); }
            xs[(size_t) k * n] = v.x;
            if (k + 1 < pss.dim[s]) xs[(size_t) (k + 1) * n] = v.y;
        }
    }
}

__global__ void __launch_bounds__(128)
k_chain_step(const __grid_constant__ DevScene sc, const __grid_constant__ PathCfg pc, const __grid_constant__ PssParams pp,
             const __grid_constant__ ChainParams cp, const __grid_constant__ FilmParams fp, const __grid_constant__ ChainArrays ca, float4 *film,
             unsigned long long *counters, dr_step_record *records, int recordStride, int steps) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t st[ST_COUNT];
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) st[i] = 0;
    if (j < ca.n) {
        const bool drmlt = pp.integrator == DR_INTEGRATOR_DRMLT;
        const int depth = ca.depth[j];
        int dims[3];
        chain_dims(pc, depth, ca.dimS, ca.dimE, ca.dimD, dims);
        const int offs[3] = { 0, ca.dimS, ca.dimS + ca.dimE };
        Pss pss;
        pss.pp = &pp; pss.stride = ca.n;
        for (int s = 0; s < 3; ++s) { pss.xs[s] = ca.X + (size_t) offs[s] * ca.n + j; pss.dim[s] = dims[s]; }
        pss.chain = ca.chainId[j];
        pss.largeStep = false; pss.lightTracing = false;
        // current state
        float Lx = ca.L[j];
        float2 posx = ca.pos[j];
        const float4 v4 = ca.val[j];
        float3 valx = f3(v4.x, v4.y, v4.z);
        float cumW = v4.w;
        int tx = ca.tcur[j];
        const uint32_t mut0 = ca.mutDone[j];
        uint32_t mut = mut0;
        const uint32_t mutEnd = mut0 + (uint32_t) steps;
        // per-mutation scratch
        int phase = 0;
        bool largeStep = false, acc1 = false, doSecond = false;
        float a1 = 0.f;
        Proposal y, z;
        int maxIdx1[3] = { 0, 0, 0 };
        y.L = 0.f; y.n = 0; z.L = 0.f; z.n = 0; z.val = f3(0.f); z.pos = make_float2(0.f, 0.f); z.t = -1;
        y.val = f3(0.f); y.pos = make_float2(0.f, 0.f); y.t = -1;
        uint32_t rays = 0;

        while (mut < mutEnd) {
            pss.mut = mut;
            if (phase == 0) {
                largeStep = keyed_uniform(pp.seed, S_COIN, pss.chain, mut, 0u) < cp.pLarge;
                pss.largeStep = largeStep;
                pss.lightTracing = false;
                pss.begin(PSS_STAGE1);
            } else if (phase == 1) {
                pss.lightTracing = cp.fixEmitterPath && tx == 1;     // nextStage(current->t == 1)
                pss.begin(PSS_STAGE2);
            } else {
                pss.begin(PSS_REVERSE);
            }
            PathResult r;
            evaluate(sc, pc, pss, depth, r, rays);
            ++st[ST_PATHS];

            float a2 = 0.f; bool acc2 = false;
            bool finish = true;
            if (!drmlt) {
                // ---------------- PSSMLT (pssmlt_proc.cpp:175-272)
                set_proposal(y, r);
                ++st[ST_MUT];
                float a = fminf(1.0f, y.L / Lx);
                if (isnan(y.L) || y.L < 0.f) a = 0.f;
                a = isnan(a) ? 1.0f : a;                               // std::min(1, NaN) = 1
                bool accept; float currentWeight, proposedWeight;
                if (a > 0.f) {
                    if (cp.kelemenWeights) {
                        currentWeight = (1.f - a) * Lx / (Lx / cp.b + cp.pLarge);
                        proposedWeight = (a + (largeStep ? 1.f : 0.f)) * y.L / (y.L / cp.b + cp.pLarge);
                    } else { currentWeight = 1.f - a; proposedWeight = a; }
                    accept = (a == 1.f) || (keyed_uniform(pp.seed, S_COIN, pss.chain, mut, 1u) < a);
                } else {
                    currentWeight = cp.kelemenWeights ? Lx / (Lx / cp.b + cp.pLarge) : 1.f;
                    proposedWeight = 0.f; accept = false;
                }
                cumW += currentWeight;
                if (records) {
                    dr_step_record &rec = records[(size_t) j * recordStride + (mut - mut0)];
                    rec.L_x = Lx; rec.L_y = y.L; rec.L_z = 0.f; rec.a1 = a; rec.a2 = 0.f;
                    rec.large_step = largeStep; rec.accept1 = accept; rec.did_second = 0; rec.accept2 = 0;
                }
                ++st[ST_ACC_B];
                if (largeStep) ++st[ST_LARGE_B]; else ++st[ST_BOLD_B];
                if (accept) {
                    const float3 v = valx * cumW;
                    if (film && !is_zero(v)) film_put(film, fp, posx, v);
                    cumW = proposedWeight;
                    commit_state(pss, ca.X, ca.n, offs, j, true, false);
                    Lx = y.L; posx = y.pos; valx = y.val; tx = y.t;
                    ++st[ST_ACC_A];
                    if (largeStep) ++st[ST_LARGE_A]; else ++st[ST_BOLD_A];
                } else {
                    const float3 v = y.val * proposedWeight;
                    if (film && y.n && !is_zero(v)) film_put(film, fp, y.pos, v);
                }
                ++mut;
                continue;
            }

            // ---------------- DRMLT
            if (phase == 0) {
                set_proposal(y, r);
                ++st[ST_MUT];
                maxIdx1[0] = pss.maxIdx[0]; maxIdx1[1] = pss.maxIdx[1]; maxIdx1[2] = pss.maxIdx[2];
                a1 = 0.f; acc1 = false;
                z.L = 0.f; z.n = 0;
                if (cp.useMixture) {   // processMixture (drmlt_proc.cpp:284-299)
                    if (!invalid_loose(y.L)) {
                        a1 = metropolis_clamp(y.L / Lx);
                        acc1 = (a1 >= 1.f) || (keyed_uniform(pp.seed, S_COIN, pss.chain, mut, 1u) < a1);
                    }
                    doSecond = !largeStep && (keyed_uniform(pp.seed, S_COIN, pss.chain, mut, 3u) < 0.5f);
                } else {               // drmlt_proc.cpp:543-558
                    if (!invalid_strict(y.L)) {
                        a1 = metropolis_clamp(y.L / Lx);
                        acc1 = (a1 >= 1.f) || (keyed_uniform(pp.seed, S_COIN, pss.chain, mut, 1u) < a1);
                    }
                    doSecond = !acc1 && (cp.timidAfterLarge || !largeStep);
                }
                if (doSecond) { phase = 1; finish = false; }
            } else if (phase == 1) {
                set_proposal(z, r);
                if (cp.useMixture) {   // drmlt_proc.cpp:317-324: plain MH on the replaced proposal
                    a1 = 0.f; acc1 = false;
                    if (!invalid_loose(z.L)) {
                        a2 = metropolis_clamp(z.L / Lx);
                        acc2 = (a2 >= 1.f) || (keyed_uniform(pp.seed, S_COIN, pss.chain, mut, 2u) < a2);
                    }
                } else if (!invalid_strict(z.L)) {
                    if (pp.type == DR_TYPE_GREEN) { phase = 2; finish = false; }
                    else if (pp.type == DR_TYPE_MIRA) {   // drmlt_proc.cpp:625-650
                        const float aReverse = metropolis_clamp(y.L / z.L);
                        if (!(aReverse >= 1.f)) {
                            const float T = largeStep ? 1.0f : mira_transition_ratio(pss, cp, maxIdx1, pss.maxIdx);
                            if (!invalid_strict(T)) {
                                a2 = metropolis_clamp((z.L / Lx) * T * (1.0f - aReverse) / (1.0f - a1));
                                acc2 = (a2 >= 1.f) || (keyed_uniform(pp.seed, S_COIN, pss.chain, mut, 2u) < a2);
                            }
                        }
                    } else {                               // orbital, drmlt_proc.cpp:655-669
                        if (z.L < y.L) { a2 = 0.f; }
                        else if (z.L >= Lx) { a2 = 1.0f; acc2 = true; }
                        else {
                            a2 = (z.L - y.L) / (Lx - y.L);
                            acc2 = (a2 >= 1.f) || (keyed_uniform(pp.seed, S_COIN, pss.chain, mut, 2u) < a2);
                        }
                    }
                }
            } else {                   // Green & Mira reverse path (drmlt_proc.cpp:588-621)
                const float Lr = r.lum;
                const float aReverse = invalid_strict(Lr) ? 0.f : metropolis_clamp(Lr / z.L);
                if (aReverse != 1.f) {
                    a2 = metropolis_clamp((z.L / Lx) * (1.f - aReverse) / (1.f - a1));
                    acc2 = (a2 >= 1.f) || (keyed_uniform(pp.seed, S_COIN, pss.chain, mut, 2u) < a2);
                }
            }
            if (!finish) continue;

            // ---- splat with expectation weights (drmlt_proc.cpp:676-688; mixture :327-333)
            const bool did2 = phase >= 1;
            if (records) {
                dr_step_record &rec = records[(size_t) j * recordStride + (mut - mut0)];
                rec.L_x = Lx; rec.L_y = y.L; rec.L_z = did2 ? z.L : 0.f; rec.a1 = a1; rec.a2 = a2;
                rec.large_step = largeStep; rec.accept1 = acc1; rec.did_second = did2; rec.accept2 = acc2;
            }
            if (film && !cp.acceptanceMap) {
                float wy, wz, wx;
                if (cp.useMixture) { wy = did2 ? 0.f : a1; wz = did2 ? a2 : 0.f; wx = 1.0f - (did2 ? a2 : a1); }
                else { wy = a1; wz = (1.0f - a1) * a2; wx = 1.0f - wy - wz; }
                if (wx > 0.f) film_put(film, fp, posx, valx * wx);
                if (wy > 0.f && y.n) film_put(film, fp, y.pos, y.val * wy);
                if (wz > 0.f && z.n) film_put(film, fp, z.pos, z.val * wz);
            }
            // ---- accept / reject, statistics (drmlt_proc.cpp:691-769; mixture :335-378)
            if (cp.useMixture) {
                const bool accept = did2 ? acc2 : acc1;
                ++st[ST_ACC_B];
                if (!did2) { ++st[ST_FIRST_B]; if (largeStep) ++st[ST_LARGE_B]; else ++st[ST_BOLD_B]; } else ++st[ST_SECOND_B];
                if (accept) {
                    ++st[ST_ACC_A];
                    if (!did2) { ++st[ST_FIRST_A]; if (largeStep) ++st[ST_LARGE_A]; else ++st[ST_BOLD_A]; } else ++st[ST_SECOND_A];
                    commit_state(pss, ca.X, ca.n, offs, j, !did2, true);
                    const Proposal &p = did2 ? z : y;
                    Lx = p.L; posx = p.pos; valx = p.val; tx = p.t;
                }
            } else if (acc1 || acc2) {
                // acceptance map: binned at the state that is being LEFT (proposed.* after the swap)
                if (film && cp.acceptanceMap && (acc1 ? !largeStep : true))
                    film_put(film, fp, posx, acc1 ? f3(1.f, 0.f, 0.f) : f3(0.f, 1.f, 0.f));
                commit_state(pss, ca.X, ca.n, offs, j, acc1, true);
                const Proposal &p = acc1 ? y : z;
                Lx = p.L; posx = p.pos; valx = p.val; tx = p.t;
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
            phase = 0;
            ++mut;
        }
        st[ST_RAYS] = rays;
        // persist the chain
        ca.L[j] = Lx; ca.pos[j] = posx; ca.val[j] = make_float4(valx.x, valx.y, valx.z, cumW);
        ca.tcur[j] = tx; ca.mutDone[j] = mut;
    }
    // per-block counters -> global (the reference's StatsCounter, statistics.h:80-110)
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) {
        unsigned int v = st[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counters[i], (unsigned long long) v);
    }
}

// PSSMLT's "last splat" of the accumulated current state (pssmlt_proc.cpp:274-279); resets the weight
__global__ void k_flush_pssmlt(const __grid_constant__ ChainArrays ca, const __grid_constant__ FilmParams fp, float4 *film) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ca.n) return;
    float4 v = ca.val[j];
    const float3 c = f3(v.x, v.y, v.z) * v.w;
    if (!is_zero(c)) film_put(film, fp, ca.pos[j], c);
    v.w = 0.f;
    ca.val[j] = v;
}

// ---------------------------------------------------------------- develop (drmlt_proc.cpp:813-854)
__global__ void k_film_luminance(const float4 *film, long long n, double *out) {
    double s = 0.0;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const float4 p = film[i];
        s += (double) luminance(f3(p.x, p.y, p.z));
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}
__global__ void k_develop(const float4 *film, long long n, float factor, float *rgb) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = film[i];
    rgb[3 * i] = p.x * factor; rgb[3 * i + 1] = p.y * factor; rgb[3 * i + 2] = p.z * factor;
}

// ---------------------------------------------------------------- replay kernels
__global__ void k_trace(const __grid_constant__ DevScene sc, const dr_ray *rays, long long n, int shadow, const unsigned int *order, dr_hit *hits) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const dr_ray r = rays[i];
    Hit h;
    const float3 o = f3(r.o[0], r.o[1], r.o[2]), d = f3(r.d[0], r.d[1], r.d[2]);
    bool hit = shadow ? traverse<true>(sc, o, d, r.mint, r.maxt, h) : traverse<false>(sc, o, d, r.mint, r.maxt, h);
    dr_hit out;
    out.prim = hit ? (int) order[h.tri] : -1;
    out.t = hit ? h.t : 0.f; out.u = hit ? h.u : 0.f; out.v = hit ? h.v : 0.f;
    hits[i] = out;
}

__global__ void __launch_bounds__(128)
k_eval_paths(const __grid_constant__ DevScene sc, const __grid_constant__ PathCfg pc, const __grid_constant__ PssParams pp, const float *us, int ds, const float *ue, int de, const float *ud, int dd,
             const int *depth, long long n, dr_path_result *out) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pss pss;
    pss.pp = &pp; pss.stride = 1;
    pss.xs[0] = us + i * ds; pss.xs[1] = ue + i * de; pss.xs[2] = ud + i * dd;
    pss.dim[0] = ds; pss.dim[1] = de; pss.dim[2] = dd;
    pss.chain = 0; pss.mut = 0; pss.largeStep = false; pss.lightTracing = false;
    pss.begin(PSS_ARRAY);
    PathResult r;
    uint32_t rays = 0;
    evaluate(sc, pc, pss, depth ? depth[i] : -1, r, rays);
    dr_path_result o;
    memset(&o, 0, sizeof(o));
    o.luminance = r.lum; o.n_splats = r.n; o.s = r.s; o.t = r.t; o.mis_weight = r.mis;
    o.pos[0][0] = r.pos.x; o.pos[0][1] = r.pos.y;
    o.value[0][0] = r.val.x; o.value[0][1] = r.val.y; o.value[0][2] = r.val.z;
    o.n_rays = (int) rays;
    out[i] = o;
}
