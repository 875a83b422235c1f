// machine.cuh -- the staged wavefront machine that runs the Markov chains: lane records, work
// queues, job description.  (Kernels: k_chain.cu, k_walk.cu, k_pt.cu, k_trace.cu.)
//
// Every chain ("lane") has at most ONE ray in flight.  Lanes move between small, specialised kernels
// through device-side work queues of lane indices, so that each kernel runs convergent code on full
// warps, with a register budget that fits its own stage only:
//
//   k_trace<closest|shadow>  float32 BVH traversal of the queued rays; a closest hit is routed to the
//                            walk queue of the BSDF model of the triangle it found (sort by material)
//   k_walk<bsdf>             hit -> path vertex (double re-intersection), MIS bookkeeping, BSDF sample of
//                            the next direction (sensor or emitter subpath)            -> next ray / connect
//   k_connect                MMLT connection of the two subpath ends (PathVertex::eval, geometric term,
//                            the four densities next to the connection)                 -> shadow ray
//   k_pt                     technique=path: MIPathTracer::Li as a resumable state machine
//   k_chain                  a path has ended: MIS weight, then the chain-level step fused in the same
//                            thread -- delayed-rejection acceptance (mira / green / orbital / mixture, or
//                            PSSMLT), expectation-weighted film splats, commit of the accepted
//                            primary-sample vector, statistics
//   k_begin<class>           the NEXT proposal is mutated (one class of transition-kernel arithmetic per warp) and the
//                            first ray of its path is emitted
//
// One round = { trace (closest + shadow), walk x materials, connect, chain, begin }; every ray emitted in
// round r is traced in round r + 1.  Lanes progress at their own pace: a lane whose first stage was
// accepted starts its next mutation while its neighbour traces a second-stage path.  The same machine
// runs three kinds of jobs: Markov chains (JOB_CHAIN), the bootstrap (JOB_BOOT,
// PathSampler::generateSeeds) and replayed primary-sample vectors (JOB_EVAL, parity tests).
//
// Behavioural parity targets:
//   PathSampler::sampleSplats                src/libbidir/pathsampler.cpp:79-571 (MMLT :84-320, PT :529-567)
//   MIPathTracer::Li                         src/integrators/path/path.cpp:123-312
//   DRMLTRenderer::process / processMixture  src/integrators/drmlt/drmlt_proc.cpp:161-380, 386-771
//   PSSMLTRenderer::process                  src/integrators/pssmlt/pssmlt_proc.cpp:110-285
#pragma once
#include "path.cuh"

// Resident CTAs (of 128 threads) per SM that each stage kernel is compiled for.  Every stage is bound by the latency of
// dependent gathers (ncu: long-scoreboard stalls dominate), so registers are traded for resident warps -- up to the point
// where spills start to cost (measured at the end of round 1: 4 / 3 / 3 / 4 CTAs for walk / connect / chain / begin, each within 2 % of
// its neighbours; traversal with
// 4-wide nodes: 6 CTAs = 85 registers, no spills -- 8 CTAs spill 64 B and are 3 % slower, 10 / 12 CTAs 18 / 37 % slower,
// 4 CTAs 6 % slower).
#ifdef DR_NO_STREAMING
#define DR_REC_LD(p) (*(p))
#define DR_REC_ST(p, v) (*(p) = (v))
#else
#define DR_REC_LD(p) __ldcs(p)
#define DR_REC_ST(p, v) __stcs(p, v)
#endif
#ifndef TRACE_MINB
#define TRACE_MINB 6
#endif
#ifndef WALK_MINB
#define WALK_MINB 4
#endif
#ifndef CONNECT_MINB
#define CONNECT_MINB 3
#endif
#ifndef CHAIN_MINB
#define CHAIN_MINB 3
#endif
#ifndef BEGIN_MINB
#define BEGIN_MINB 4
#endif

// ------------------------------------------------------------------ lane records (AoS: one record per lane per array,
// every record a multiple of 16 bytes, so a lane picked from a queue is fetched with a few 128-bit loads)
enum { PS_IDLE = 0, PS_START, PS_SENSOR_HIT, PS_EMITTER_HIT, PS_CONNECT, PS_CONNECT_SHADOW, PS_FINISH, PS_EMPTY,
       PS_PT_HIT, PS_PT_NEE, PS_PT_DONE, PS_BD_EHIT, PS_BD_SHIT, PS_BD_SHADOW, PS_BD_DONE, PS_BD_BATCH };
enum { PH_STAGE1 = 0, PH_STAGE2 = 1, PH_REVERSE = 2, PH_INIT = 3 };
enum { F_DELTA = 1u, F_ANYCONN = 2u, F_SPOS_FAIL = 4u, F_PT_FIRST = 8u, F_PT_EMITTED = 16u, F_PT_DIRECT = 32u, F_PT_NONSPEC = 64u,
       F_PT_NEEPEND = 128u /* the shadow ray of the previous vertex's direct-illumination sample travels beside the BSDF-sampled ray */ };
#define Q_DEFERRED 0x40000000u        /* shadow-queue item: lane | Q_DEFERRED = the result goes to LaneMem::neeOcc, the lane is not queued (k_pt.cu) */

struct alignas(16) Core {     // 128 bytes
    uint8_t pstate, s, t, j;          // path in flight: state, strategy, vertices walked on the current side
    uint8_t depth, phase, large, ubuf;   // MMLT depth | chain phase | large-step coin (2 = not drawn) | active coordinate buffer
    int8_t tx; uint8_t nrays, pos0, pos1;   // t of the current state | rays cast | reader positions (sensor, emitter)
    uint8_t pos2, pad0, pad1, pad2;   // reader position (direct)
    uint32_t flags, connectable;
    uint32_t mut;                     // JOB_CHAIN: mutation counter | JOB_BOOT / JOB_EVAL: items done by this lane
    uint32_t pad3;
    float2 spos;                      // pixel of the path in flight
    uint64_t chainId;                 // RNG key of the chain
    uint64_t seedIdx;                 // bootstrap sample the chain started from
    uint64_t pad4;
    R3 weight;                        // MMLT: product of walk weights, then the connection value | PT: throughput
    R3 d;                             // direction of the ray in flight (double; the traversal gets its float cast)
    Real pdfFwd, pdfBwd;              // densities of the step in flight (solid angle or discrete)
};
static_assert(sizeof(Core) == 128, "Core layout");

struct alignas(16) PredRec { R3 p, ng; Real pad[2]; };     // 64 bytes (two 32-byte sectors: no partial-sector writes)
struct alignas(16) PtExtra { R3 Li, pending, refN; Real eta, bsPdf; R3 dIn; Real pad[2]; };   // 128 bytes (aliases the emitter-side vertex)
struct alignas(16) ChainCore {        // 128 bytes
    Real Lx, a1, cumW;
    Real yL, zL;
    float2 posx, ypos, zpos;
    float3 valx, yval, zval;
    uint8_t acc1, yn, zn, xl;         // yn / zn: splat 0 of y / z exists | xl, yl, zl: light-image splats of x / y / z (bdpt)
    int8_t yt, zt, ys, zs;
    uint8_t posY[3], yl;              // reader positions at the end of the stage-1 path (m_dimStage1, drmlt_sampler.cpp:237-238)
    uint8_t zl, pad0[3];
    uint32_t pad2[3];
};
static_assert(sizeof(ChainCore) == 128 && sizeof(Vtx) == 128 && sizeof(PtExtra) == sizeof(Vtx) && sizeof(PredRec) == 64, "lane layout");

// ---- technique=bdpt: both subpaths are kept (every vertex can be a connection end point)
#define BD_MAXV (DR_MAXK)             // vertices per subpath, supernode = index 0 (maxDepth + 2 <= BD_MAXV)
#define BD_MAXS DR_MAX_SPLATS         // splats of one path: splat 0 + light-image splats
enum { BD_E = 0, BD_S = 1 };          // emitter / sensor subpath
struct alignas(16) BExtra {           // 64 bytes per subpath vertex v
    R3 prefix;                        // product of the sampling weights (and Russian-roulette weights) of vertices 0 .. v-1
    Real fwd;                         // area density of v when generated from v-1 (pdf[mode] of v-1)
    Real bwd;                         // area density of v when generated from v+1 (pdf[1-mode] of v+1)
    Real conv;                        // len^2 / |cos cos| of edge (v, v+1), geometric normals (path.cpp:875-899)
    uint32_t discrete, pad[3];        // v sampled a delta direction (measure == EDiscrete)
};
struct alignas(16) BdAcc {            // 128 bytes: the splat list being built (aliases nothing; own array)
    R3 val0;                          // splat 0: accumulated value of all t >= 2 strategies
    Real lum;                         // SplatList::luminance
    float2 pos0;
    int has0, nl;                     // splat 0 exists (sensor subpath has >= 2 vertices) | light-image splats so far
    int ns, nt;                       // last vertex index of the emitter / sensor subpath
    int pad[2];
    Real pdfs[4];                     // the four densities next to the connection in flight (path.cpp:835-859)
    Real pad2[4];
};
// Batched connections (PathCfg::bdBatch): all (s, t) pairs of a path are evaluated in ONE round -- their shadow rays travel together
// through the Q_BDS queue, tagged lane * BD_MAXC + connection index -- instead of one pair per round.
#define BD_MAXC 64                    // connections per path: (maxDepth + 1)(maxDepth + 2) / 2 <= BD_MAXC (one visibility bit each)
struct alignas(16) BdConn {           // 80 bytes: one evaluated connection waiting for its shadow ray
    R3 value;                         // prefix_s prefix_t f_s f_t G, MIS weight not yet applied
    Real pdfs[4];                     // the four densities next to the connection (path.cpp:835-859)
    float2 spos;                      // t = 1: pixel of the light-image splat
    uint8_t s, t, needsRay, pad[5];
    Real pad2;
};
static_assert(sizeof(BExtra) == 64 && sizeof(BdAcc) == 128 && sizeof(BdConn) == 80, "bdpt records");

// MIS bookkeeping of the MMLT walks: ONE 32-byte record per walk step, written as a whole sector.
//   misrec[side][j] = { fwdNext, bwdPrev, conv, - }: written when vertex j of `side` (0 sensor, 1 emitter) has produced
//   vertex j + 1:  fwdNext = area density of vertex j+1 generated from j, bwdPrev = area density of vertex j-1 generated
//   from j (reverse direction), conv = len^2 / |cos cos| of edge (j, j+1) with geometric normals (path.cpp:875-899).
//   conn[0..3] = the four densities next to the connection, recomputed by k_connect (path.cpp:835-859):
//   pdfImp[s+1], pdfRad[s-1], pdfRad[s], pdfImp[s+2].
enum { MR_FWD = 0, MR_BWD = 1, MR_CONV = 2, MR_WORDS = 4, MR_MAXV = 16, SIDE_S = 0, SIDE_E = 1 };
static_assert(DR_MAXK + 1 <= MR_MAXV, "MIS records");

struct LaneMem {
    Core *core;               // [n]
    Vtx *vt, *vs;             // [n] last vertex of the sensor / emitter subpath
    PredRec *geo;             // [n][2 sides][2]: position + geometric normal of the last two vertices of each subpath;
                              //   vertex j lives in slot j & 1 (the walk overwrites j-1 with j+1: no copies)
    ChainCore *chain;         // [n]
    double *misrec;           // [n][2 sides][mrSlots][MR_WORDS]: per-step MIS records (see MR_*); mrSlots = maxDepth + 2 <= MR_MAXV
    double *conn;             // [n][4]: densities next to the connection
    double *ubuf;             // [n][ubCount][nU] coordinate buffers X, Y (pssmlt), Z (drmlt), R (green): only the buffers the
                              //   integrator can touch are allocated
    double *rayd;             // [n][8] o, d, tmin, tmax of the same ray un-rounded (deciding triangle tests)
    // technique=path only (else null): the deferred direct-illumination shadow ray of a lane (k_pt.cu)
    double *rayd2;            // [n][8] its un-rounded ray
    int *neeOcc;              // [n] 1: it found an occluder
    // technique=bdpt only (else null)
    Vtx *bv;                  // [n][2][BD_MAXV] subpath vertices
    BExtra *bx;               // [n][2][BD_MAXV]
    BdAcc *bacc;              // [n]
    float4 *bsplat;           // [n][4][BD_MAXS][2]: light-image splats (pos.xy | rgb) of the lists x, y, z and of the path in flight
    // ... with batched connections only (else null)
    BdConn *bconn;            // [n][bdStride] evaluated connections of the path in flight, in the reference's (s, t) order
    double *brayd;            // [n][bdStride][8] their un-rounded shadow rays
    unsigned long long *bvis; // [n] bit i: shadow ray of connection i found no occluder
    uint32_t *bpend;          // [n] shadow rays still in flight
    uint32_t *bcount;         // [n] connections recorded
    int bdStride;             // (maxDepth + 1)(maxDepth + 2) / 2
    int n, nU;
    int ubCount, mrSlots;
};

// whole-record copies through 128-bit accesses.  Lane records are touched once per round and are far larger than L2
// in total, so they use the streaming (evict-first) cache operators: the BVH and triangle lines that every ray shares
// are what should stay resident in L2.
template <class T> DR_D void rec_load(T &dst, const T *src) {
    static_assert(sizeof(T) % 16 == 0, "record size");
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(&dst);
#pragma unroll
    for (int i = 0; i < (int) (sizeof(T) / 16); ++i) d[i] = DR_REC_LD(s + i);
}
template <class T> DR_D void rec_store(T *dst, const T &src) {
    static_assert(sizeof(T) % 16 == 0, "record size");
    const uint4 *s = reinterpret_cast<const uint4 *>(&src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < (int) (sizeof(T) / 16); ++i) DR_REC_ST(d + i, s[i]);
}

// ------------------------------------------------------------------ work queues
// Q_RAYC / Q_RAYS / Q_CHAIN are double-buffered by round parity: kernels of round r consume [r & 1] and
// produce into [(r + 1) & 1] (Q_CHAIN is also fed in-round by trace / walk / connect).  Q_WALK, Q_CONNECT, Q_PT and
// Q_BEGIN are produced and consumed inside one round.
enum { Q_RAYC = 0, Q_RAYS = 2, Q_CHAIN = 4, Q_BDS = 6 /* x2 parities: batched BDPT shadow rays */, Q_WALK = 8 /* + bsdf type, 7 */, Q_CONNECT = 15, Q_PT = 16,
       Q_BEGIN = 17 /* + class, 3 */, Q_COUNT = 20 };
#define N_WALK_CLASSES 7
// classes of "start the next path" work: each runs one kind of proposal arithmetic on full warps
enum { BEGIN_STAGE1 = 0, BEGIN_STAGE2 = 1, BEGIN_OTHER = 2 };
struct RayF { float4 a, b; };                               // (o, tmin), (d, tmax): float32 cast of a ray, for the traversal
struct Queues {
    uint32_t *items;          // [Q_COUNT][n]
    uint32_t *count;          // [Q_COUNT] (+ 2 head counters of the ray queues)
    float4 *rays;             // [4][n][2]: the float32 rays of the four ray queues (Q_RAYC x2, Q_RAYS x2), parallel to
                              // items -- the traversal kernels read their input with coalesced, independent loads
    uint32_t *bitems;         // [2][bn] batched BDPT shadow rays (Q_BDS, by parity): lane * BD_MAXC + connection index
    float4 *brays;            // [2][bn][2] their float32 rays
    int bn;                   // capacity per parity: n * bdStride
    int *aux;                 // [Q_COUNT][n]: result of the traversal that queued the lane (leaf-order triangle, -1 = miss),
                              // parallel to items: the consumer reads it coalesced instead of gathering a per-lane record
    int n;
};
// Opportunistic warp-aggregated append: the threads of the warp that push to the same queue at the same
// time share one atomic.
DR_D uint32_t q_push(const Queues &q, int which, uint32_t lane) {
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, which);
    const int leader = __ffs(peers) - 1;
    const unsigned self = threadIdx.x & 31u;
    uint32_t base = 0;
    if ((int) self == leader) base = atomicAdd(&q.count[which], (uint32_t) __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t slot = base + __popc(peers & ((1u << self) - 1u));
    q.items[(size_t) which * q.n + slot] = lane;
    return slot;
}
// append a lane together with the traversal result that sends it there
DR_D void q_push_hit(const Queues &q, int which, uint32_t lane, int tri) {
    const uint32_t slot = q_push(q, which, lane);
    q.aux[(size_t) which * q.n + slot] = tri;
}
// append a lane and, when the destination is a ray queue, its ray
DR_D void q_push_ray(const Queues &q, int which, uint32_t lane, const RayF &ray) {
    const uint32_t slot = q_push(q, which, lane);
    if (which < Q_CHAIN) {
        float4 *r = q.rays + 2 * ((size_t) which * q.n + slot);
        r[0] = ray.a; r[1] = ray.b;
    }
}

// The queue counters are emptied by the kernels of the round themselves (no launch of its own).  k_begin, the last kernel
// of round r, empties the in-round walk / connect / path-tracer queues (their consumers ran before it) and the head
// counters of the dynamic ray fetch; k_trace, the first kernel of round r + 1, empties the ray / chain queues that round r
// consumed -- the ones round r + 1 produces into once the traversal is done -- and the begin-class queues k_chain fills later
// in its round (trace_recycle).
DR_D void queues_recycle(const Queues &q) {
    const int t = threadIdx.x;
    if (blockIdx.x == 0 && t >= Q_WALK && t < Q_COUNT + 2 && !(t >= Q_BEGIN && t < Q_BEGIN + 3)) q.count[t] = 0;
}
DR_D void trace_recycle(const Queues &q, int parity) {
    const int t = threadIdx.x;
    if (blockIdx.x == 0 && t < Q_WALK && (t & 1) == (parity ^ 1)) q.count[t] = 0;
    if (blockIdx.x == 0 && t >= Q_BEGIN && t < Q_BEGIN + 3) q.count[t] = 0;
}

// Several queues in ONE launch: the global warp index space is the concatenation of the queues, each rounded up to
// whole warps, so that every warp works on one queue (one kind of work) and the classes run side by side instead of
// each paying for a launch and a latency-bound tail of its own.  Returns false when `warp` is past the end.
template <int K>
DR_D bool multiq_locate(const uint32_t (&cnt)[K], uint32_t warp, int &cls, uint32_t &qi) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t w = (cnt[k] + 31u) >> 5;
        if (warp < w) { cls = k; qi = (warp << 5) + (threadIdx.x & 31u); return true; }
        warp -= w;
    }
    return false;
}

// ------------------------------------------------------------------ jobs
enum { JOB_CHAIN = 0, JOB_BOOT = 1, JOB_EVAL = 2 };
struct JobParams {
    int type;
    uint32_t mutTarget;                 // JOB_CHAIN: lanes run until they have done this many mutations
    long long nItems;                   // JOB_BOOT / JOB_EVAL: lane l evaluates items l, l + n, l + 2n, ...
    unsigned long long first;           // JOB_BOOT: bootstrap sample index of item 0
    float *lumOut;                      // JOB_BOOT: luminance of every item
    float *lumTargetOut;                // JOB_BOOT, two-stage MLT: luminance after the importance re-weighting (the chains' target)
    const float *us, *ue, *ud;          // JOB_EVAL: replayed primary-sample vectors [nItems][d*]
    int ds, de, dd;
    const int *depthIn;                 // JOB_EVAL: MMLT depth per item
    dr_path_result *out;                // JOB_EVAL
    dr_step_record *records;            // JOB_CHAIN (parity): one record per mutation
    int recordStride;
    uint32_t mut0;
    // JOB_CHAIN, work-unit queue (null cursor: chains are resident, one per lane): a lane whose chain has done its
    // `mutTarget` mutations takes the next chain of [.., chainEnd) from the cursor -- the reference's work units handed
    // out by DRMLTProcess::generateWork (drmlt_proc.cpp:869-883), one seed + nMutations mutations each
    unsigned int *chainCursor;
    unsigned int chainFirst, chainEnd;
    const int *qDepth;                  // [chains] MMLT depth | chain id (RNG key) | bootstrap sample of every chain
    const unsigned long long *qChainId, *qSeedIdx;
    // JOB_CHAIN, replayed chains of the reference (dr_chain_replay): lane l reads its uniforms from replay + l * replayStride
    const double *replay;
    long long replayStride;
    int replayDim;
    // JOB_CHAIN, depth-balanced MMLT chains (dr_config.depth_balance): 16.16 fixed-point factor per path depth; a chain of
    // depth d runs (mutTarget * mutScale[d]) >> 16 mutations.  Null: every chain runs mutTarget mutations.
    const uint32_t *mutScale;
};
// the number of mutations a chain of path depth `depth` has to reach
__device__ __forceinline__ uint32_t mut_target(const JobParams &job, int depth) {
    if (!job.mutScale || depth < 0) return job.mutTarget;
    return (uint32_t) (((unsigned long long) job.mutTarget * __ldg(job.mutScale + depth)) >> 16);
}

struct FilmParams {
    int w, h;
    float radius, scaleFactor;
    float values[32];          // rfilter.cpp:37-55 discretised filter (MTS_FILTER_RESOLUTION = 31)
};

struct ChainParams {
    Real pLarge;
    Real b;                    // m_config.luminance
    int acceptanceMap, timidAfterLarge, fixEmitterPath, useMixture, kelemenWeights;
    Real kel_s1, kel_s2, kel_logRatio;     // un-scaled Kelemen bounds for Mira's transition ratio
    int dimS, dimE, dimD;      // allocated coordinates per sampler (maxDepth worst case)
    const float *importance;   // two-stage MLT: m_config.importanceMap at film resolution (null: single stage)
};

enum { ST_MUT = 0, ST_FIRST_A, ST_FIRST_B, ST_LARGE_A, ST_LARGE_B, ST_BOLD_A, ST_BOLD_B, ST_SECOND_A, ST_SECOND_B,
       ST_SECOND_LARGE_A, ST_SECOND_LARGE_B, ST_SECOND_BOLD_A, ST_SECOND_BOLD_B, ST_ACC_A, ST_ACC_B, ST_PATHS, ST_RAYS, ST_COUNT };

// everything a stage kernel needs, passed as ONE __grid_constant__ argument
struct Machine {
    DevScene sc;
    PathCfg pc;
    PssParams pp;
    ChainParams cp;
    FilmParams fp;
    LaneMem lm;
    Queues q;
    JobParams job;
    float4 *film;
    unsigned long long *counters;
    int parity;               // round & 1
    int laneBegin, laneEnd;   // lanes of this wavefront group (set-up kernels; the stage kernels only see queue items)
    int traceRefill, traceDescend;   // k_trace tuning: refill below this many busy lanes / leave the descend phase below this many
};

// uniforms of a chain: the replay table of the lane (dr_chain_replay) or keyed Philox (S_COIN: 0 large step, 1 accept 1, 2 accept 2, 3 mixture)
DR_D ReplayTable replay_table(const Machine &M, int lane) {
    ReplayTable t;
    t.base = M.job.replay ? M.job.replay + (size_t) lane * M.job.replayStride : nullptr;
    t.dim = M.job.replayDim;
    return t;
}
DR_D Real chain_coin(const Machine &M, int lane, const Core &c, int which);

// ------------------------------------------------------------------ MIS arrays of a lane
DR_D PredRec *geo_slot(const Machine &M, int lane, int side, int j) { return M.lm.geo + ((size_t) lane * 2 + side) * 2 + (j & 1); }
DR_D double *misrec_slot(const Machine &M, int lane, int side, int j) { return M.lm.misrec + (((size_t) lane * 2 + side) * M.lm.mrSlots + j) * MR_WORDS; }
DR_D void misrec_store(const Machine &M, int lane, int side, int j, Real fwdNext, Real bwdPrev, Real conv) {
    double4 v = make_double4(fwdNext, bwdPrev, conv, 0.0);
    *reinterpret_cast<double4 *>(misrec_slot(M, lane, side, j)) = v;      // one aligned 32-byte sector
}

DR_D Real chain_coin(const Machine &M, int lane, const Core &c, int which) {
    if (M.job.replay) return replay_table(M, lane).coin(c.mut, which);
    return (Real) keyed_uniform(M.pp.seed, S_COIN, c.chainId, c.mut, (uint32_t) which);
}

// findMaxDimensions (pssmlt_utils.h:27-77): MMLT vectors depend on the chain's depth
DR_D void chain_dims(const PathCfg &pc, const ChainParams &cp, int depth, int dims[3]) {
    if (pc.technique == DR_TECH_MMLT) {
        int m = (depth + 2) * 3; if (m & 1) m++;
        dims[0] = m; dims[1] = m; dims[2] = 1;
    } else { dims[0] = cp.dimS; dims[1] = cp.dimE; dims[2] = cp.dimD; }
}

// MMLT strategy from the direct sampler's coordinate (pathsampler.cpp:104-129)
DR_D void mmlt_strategy(const PathCfg &pc, int depth, Real decision, int &s, int &t) {
    int nStrats;
    if (pc.lightImage) { nStrats = depth + 1; s = min((int) (nStrats * decision), nStrats - 1); t = nStrats - s; }
    else { nStrats = depth; s = min((int) (nStrats * decision), nStrats - 1); t = 1 + (nStrats - s); }
}

// coordinate pairs a subpath of n vertices can consume: 2n coordinates, or up to 3n - 2 when BSDF samples may draw a
// third number (rough dielectrics)
DR_D int subset_pairs(const Machine &M, int n) { return M.pc.hasRoughDielectric ? (3 * n) / 2 : n; }
// number of coordinate PAIRS of each sampler that a proposal with strategy (s, t) must carry
DR_D void pair_extent(const Machine &M, const int dims[3], int s, int t, int ext[3]) {
    if (M.pp.subset) { ext[0] = subset_pairs(M, t); ext[1] = subset_pairs(M, s); ext[2] = 1; }
    else { ext[0] = (dims[0] + 1) >> 1; ext[1] = (dims[1] + 1) >> 1; ext[2] = (dims[2] + 1) >> 1; }
}

// coordinate reader over the lane's active buffer
DR_D void reader_open(const Machine &M, const Core &c, int lane, UReader &rd) {
    rd.buf = M.lm.ubuf + ((size_t) lane * M.lm.ubCount + c.ubuf) * M.lm.nU;
    rd.off[0] = M.pp.off[0]; rd.off[1] = M.pp.off[1]; rd.off[2] = M.pp.off[2];
    if (M.job.type == JOB_EVAL) { rd.lim[0] = M.job.ds; rd.lim[1] = M.job.de; rd.lim[2] = M.job.dd; }
    else {
        int dims[3];
        chain_dims(M.pc, M.cp, c.depth, dims);
        rd.lim[0] = dims[0] + (dims[0] & 1); rd.lim[1] = dims[1] + (dims[1] & 1); rd.lim[2] = dims[2] + (dims[2] & 1);
    }
    rd.pos[0] = c.pos0; rd.pos[1] = c.pos1; rd.pos[2] = c.pos2;
    rd.reflect = M.job.type == JOB_CHAIN && M.pp.integrator == DR_INTEGRATOR_DRMLT;
}
DR_D void reader_close(const UReader &rd, Core &c) { c.pos0 = (uint8_t) rd.pos[0]; c.pos1 = (uint8_t) rd.pos[1]; c.pos2 = (uint8_t) rd.pos[2]; }

// the next ray of a lane (un-rounded copy -> lane record, float32 cast -> `ray`, queued by q_push_ray); mint == epsilon gets the adaptive scaling of skdtree.cpp:126-129
DR_D void emit_ray(const Machine &M, int lane, Core &c, R3 o, R3 d, Real tmin, Real tmax, RayF &ray) {
    if (tmin == (Real) M.sc.epsilon) tmin *= fmax(fmax(fmax(fabs(o.x), fabs(o.y)), fabs(o.z)), (Real) M.sc.epsilon);
    ray.a = make_float4((float) o.x, (float) o.y, (float) o.z, (float) tmin);
    ray.b = make_float4((float) d.x, (float) d.y, (float) d.z, (float) tmax);
    double2 *rd = reinterpret_cast<double2 *>(M.lm.rayd + 8 * (size_t) lane);
    rd[0] = make_double2(o.x, o.y); rd[1] = make_double2(o.z, d.x); rd[2] = make_double2(d.y, d.z); rd[3] = make_double2(tmin, tmax);
    c.d = d;
    ++c.nrays;
}

// ------------------------------------------------------------------ film
// Splat of one (position, RGB) pair through the tabulated reconstruction filter
// (ImageBlock::put, include/mitsuba/render/imageblock.h:149-196): one 16-byte vector atomic per touched pixel.
#ifdef DR_FILM_MATCH_STATS
// measurement build (tools/build_variant.sh stats -DDR_FILM_MATCH_STATS): how often do two threads that splat at the same time share
// a filter footprint -- i.e. what a warp-aggregated splat (north_star (e)) could merge?  Printed by dr_job_destroy.
static __device__ unsigned long long g_filmMatch[3];      // film_put calls | calls whose footprint another active thread shares | threads active per call
#endif
static __device__ __noinline__ void film_put(float4 *film, const FilmParams &fp, float2 pos, float3 value) {
    if (!rgb_valid(value)) return;
    const float px = pos.x - 0.5f, py = pos.y - 0.5f;
    const int minx = max((int) ceilf(px - fp.radius), 0), miny = max((int) ceilf(py - fp.radius), 0);
    const int maxx = min((int) floorf(px + fp.radius), fp.w - 1), maxy = min((int) floorf(py + fp.radius), fp.h - 1);
#ifdef DR_FILM_MATCH_STATS
    {
        const unsigned act = __activemask();
        const unsigned peers = __match_any_sync(act, miny * fp.w + minx);
        atomicAdd(&g_filmMatch[0], 1ull);
        if (__popc(peers) > 1) atomicAdd(&g_filmMatch[1], 1ull);
        atomicAdd(&g_filmMatch[2], (unsigned long long) __popc(act));
    }
#endif
    for (int y = miny; y <= maxy; ++y) {
        const float wy = fp.values[min((int) fabsf((y - py) * fp.scaleFactor), 31)];
        for (int x = minx; x <= maxx; ++x) {
            const float w = fp.values[min((int) fabsf((x - px) * fp.scaleFactor), 31)] * wy;
            if (w == 0.f) continue;
            atomicAdd(film + (size_t) y * fp.w + x, make_float4(w * value.x, w * value.y, w * value.z, 0.f));
        }
    }
}

// per-warp statistics -> global counters (the reference's StatsCounter, statistics.h:80-110)
DR_D void stats_flush(const uint32_t *st, unsigned long long *counters) {
    unsigned int any = 0;
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) any |= st[i];
    if (__any_sync(0xffffffffu, any != 0)) {
#pragma unroll
        for (int i = 0; i < ST_COUNT; ++i) {
            unsigned int v = st[i];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counters[i], (unsigned long long) v);
        }
    }
}

// MMLT emitter end of the path, part 1 (begin_path, when the path starts): sample emitter subpath vertex 1 and, for
// s >= 2, the emission direction (Scene::sampleEmitterPosition scene.cpp:1066-1082, vertex.cpp:99-124,
// area.cpp:130-138).  Nothing here depends on the sensor subpath, so it is done up front, where every lane of the
// warp does the same thing; the results wait in the lane's emitter-side records:
//   vs = the emitter sample (vs.ss holds the position-sampling weight m_power / emPdf), its position / normal also in
//   geo[emitter][1];  geo[emitter][0] = { emission direction, its solid-angle density } until vertex 2 overwrites it;
//   misrec[emitter][0].fwdNext = area density of the emitter sample.
// Returns false when there is nothing to sample (the path is dead).
DR_D bool mmlt_emitter_sample(const Machine &M, int lane, Core &c, UReader &rd) {
    const DevScene &sc = M.sc;
    if (c.s == 0) return true;
    if (sc.nEmitters == 0) return false;
    EmitterPoint ep;
    const R2 u0 = rd.next2D(SMP_EMITTER);
    sample_emitter_point(sc, u0.x, u0.y, ep);
    const DevEmitter &em = sc.emitters[ep.emitter];
    misrec_store(M, lane, SIDE_E, 0, ep.pdfArea, 1.0, 0.0);
    Vtx vs;
    vs.p = ep.p; vs.ng = vs.ns = ep.n; vs.type = V_EMITTER_SAMPLE; vs.degenerate = 0; vs.emitter = ep.emitter; vs.mat = -1;
    vs.ss = emitter_radiance(sc, ep.emitter) * (R_PI * em.area / ep.emPdf);   // m_power / emPdf
    rec_store(M.lm.vs + lane, vs);
    if (c.s >= 2) {
        PredRec g1;
        g1.p = vs.p; g1.ng = vs.ng; g1.pad[0] = g1.pad[1] = 0.;
        rec_store(geo_slot(M, lane, SIDE_E, 1), g1);
        const R2 u = rd.next2D(SMP_EMITTER);
        const R3 local = square_to_cosine_hemisphere(u.x, u.y);
        R3 fs, ft;
        coordinate_system(vs.ns, fs, ft);
        PredRec e;
        e.p = fs * local.x + ft * local.y + vs.ns * local.z;
        e.ng = r3(R_INV_PI * local.z, 0., 0.); e.pad[0] = e.pad[1] = 0.;
        rec_store(geo_slot(M, lane, SIDE_E, 0), e);
    }
    return true;
}
// Part 2 (when the sensor subpath is complete): switch the lane to its emitter subpath -- launch the emission ray
// (s >= 2) or go straight to the connection.  Returns the queue the lane goes to (Q_RAYC or Q_CONNECT).
DR_D int mmlt_emitter_launch(const Machine &M, int lane, Core &c, RayF &ray) {
    c.flags &= ~F_DELTA;
    c.connectable |= 1u;                          // area lights: supernode not degenerate, never discrete
    if (c.s >= 1) {
        Vtx vs;
        rec_load(vs, M.lm.vs + lane);
        c.weight *= vs.ss;
        c.connectable |= 1u << 1;
        c.j = 1;
        if (c.s >= 2) {
            PredRec e;
            rec_load(e, geo_slot(M, lane, SIDE_E, 0));
            c.pdfFwd = e.ng.x; c.pdfBwd = 1.0;
            c.pstate = PS_EMITTER_HIT;
            emit_ray(M, lane, c, vs.p, e.p, M.sc.epsilon, INFINITY, ray);
            return Q_RAYC;
        }
    }
    c.pstate = PS_CONNECT;
    return Q_CONNECT;
}

// ------------------------------------------------------------------ host-callable launchers (one per translation unit)
struct LaunchCfg { cudaStream_t stream; int nLanes; };
// Grid of a stage kernel (grid-stride loops over a queue): enough CTAs for `n` items, capped at `stageCtasPerSm` CTAs per SM so
// that the stage kernels of different wavefront groups share the SMs instead of each flooding the device (DRMLT_STAGE_CTAS).
int stage_ctas_per_sm();                                                  // drmlt_b200.cu
int trace_ctas_per_sm();                                                  // drmlt_b200.cu: 0 = the occupancy limit
inline unsigned stage_grid(int n, int threads) { return (unsigned) std::max(1, std::min((n + threads - 1) / threads, 148 * stage_ctas_per_sm())); }
void trace_init();                                                        // k_trace.cu: occupancy query, once, outside stream capture
void launch_trace(const Machine &M, const LaunchCfg &lc);                 // k_trace.cu: closest + shadow queues
void launch_walk(const Machine &M, const LaunchCfg &lc, unsigned typeMask);   // k_walk.cu: walk queues of the BSDF types present, then connect
void launch_pt(const Machine &M, const LaunchCfg &lc);                    // k_pt.cu
void launch_bdpt(const Machine &M, const LaunchCfg &lc);                  // k_bdpt.cu: (k_bd_shadow: the batched shadow rays,) k_bdpt
bool chain_begin_fused(int nLanes, int integrator);                       // k_chain.cu: small pssmlt groups start the next path inside k_chain (no k_begin launch)
void launch_chain(const Machine &M, const LaunchCfg &lc);                 // k_chain.cu: k_chain (chain step), then k_begin (start of the next path, one class per warp)
void launch_setup(const Machine &M, const LaunchCfg &lc, const int *depth, const unsigned long long *chainId,
                  const unsigned long long *seedIdx);                     // k_chain.cu: initialise lanes for M.job and queue them
void launch_resume(const Machine &M, const LaunchCfg &lc);                // k_chain.cu: re-queue idle chains whose target was raised
void launch_flush_pssmlt(const Machine &M, const LaunchCfg &lc);          // k_chain.cu
void launch_direct(const DevScene &sc, const FilmParams &fp, unsigned long long seed, int pixelSamples, int shadingSamples, float4 *film, float *rgb,
                   double *li, cudaStream_t stream);                       // k_direct.cu: weighted film -> normalised rgb
void launch_splat_points(const FilmParams &fp, float4 *film, const float *pos, const float *rgb, long long n, cudaStream_t s);   // k_chain.cu
void launch_trace_rays(const DevScene &sc, const dr_ray *rays, long long n, int shadow, const unsigned int *order, dr_hit *hits, cudaStream_t stream);
void launch_texture_eval(const DevScene &sc, uint32_t texture, const double *uv, long long n, double *rgb, cudaStream_t stream);   // k_direct.cu
