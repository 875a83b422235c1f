// k_trace.cu -- ray traversal stages of the wavefront machine (machine.cuh): one thread per queued ray,
// float32 BVH traversal only (traverse.cuh); small register footprint, so many warps are resident to hide the
// node / triangle gather latency.  A closest hit is routed to the walk queue of the BSDF model of the triangle it
// found (the model is packed into the triangle record at scene creation), a miss or a shadow-ray result to the
// chain queue (MMLT), everything to the path-tracer queue for technique=path.
#include "machine.cuh"
#include <algorithm>

// Persistent warps with dynamic ray fetch: incoherent rays of one warp finish after very different numbers of node
// visits, so lanes whose ray is done pull the next ray from the queue (one warp-aggregated atomic on the queue's head
// counter) as soon as fewer than REFILL lanes of the warp are still traversing.

// One kernel traces both ray queues of the round (closest-hit rays first, then the shadow rays of the connections):
// the handful of shadow rays no longer pays for a launch and a latency-bound tail of its own.
__global__ void __launch_bounds__(128, TRACE_MINB)
k_trace(const __grid_constant__ Machine M) {
    const uint32_t cntC = M.q.count[Q_RAYC + M.parity], cntS = M.q.count[Q_RAYS + M.parity], cnt = cntC + cntS;
    const uint32_t *itemsC = M.q.items + (size_t) (Q_RAYC + M.parity) * M.q.n, *itemsS = M.q.items + (size_t) (Q_RAYS + M.parity) * M.q.n;
    const float4 *raysC = M.q.rays + 2 * (size_t) (Q_RAYC + M.parity) * M.q.n, *raysS = M.q.rays + 2 * (size_t) (Q_RAYS + M.parity) * M.q.n;
    uint32_t *head = M.q.count + Q_COUNT;
    if (blockIdx.x == 0 && threadIdx.x == 0 && cnt) atomicAdd(&M.counters[ST_RAYS], (unsigned long long) cnt);
    trace_recycle(M.q, M.parity);
    const bool pt = M.pc.technique != DR_TECH_MMLT;
    const unsigned self = threadIdx.x & 31u;
    int stack[DR_STACK];
    Traversal tr;
    tr.done = true; tr.anyhit = false;
    int lane = -1;
    bool exhausted = false;                                  // warp-uniform: the queue has no rays left
    for (;;) {
        // ---- refill the idle lanes of this warp
        if (!exhausted) {
            const unsigned idle = __ballot_sync(0xffffffffu, lane < 0);
            if (idle) {
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if ((int) self == leader) base = atomicAdd(head, (uint32_t) __popc(idle));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (lane < 0) {
                    const uint32_t qi = base + __popc(idle & ((1u << self) - 1u));
                    if (qi < cnt) {                          // three independent, coalesced loads: no dependent gather
                        const bool shadow = qi >= cntC;
                        const uint32_t k = shadow ? qi - cntC : qi;
                        const float4 *rays = shadow ? raysS : raysC;
                        const float4 a = __ldcs(rays + 2 * (size_t) k), b = __ldcs(rays + 2 * (size_t) k + 1);
                        lane = (int) __ldcs((shadow ? itemsS : itemsC) + k);      // (shadow queue of technique=path: may carry Q_DEFERRED)
                        const double *rayd = (lane & Q_DEFERRED) ? M.lm.rayd2 + 8 * (size_t) (lane & ~Q_DEFERRED) : M.lm.rayd + 8 * (size_t) lane;
                        tr.begin(stack, shadow, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, b.w, rayd);
                    }
                }
                exhausted = base + (uint32_t) __popc(idle) >= cnt;
            }
        }
        if (!__any_sync(0xffffffffu, lane >= 0)) break;
        // ---- while-while traversal: (A) every busy lane descends inner nodes until it holds a leaf -- lanes that found
        //      one wait, and the phase ends once fewer than TRACE_DESCEND lanes are still descending; (B) the lanes that
        //      hold a leaf intersect its triangles together.  Repeat until too few lanes are busy (refill) or all are done.
        for (;;) {
            for (;;) {
                const bool descending = lane >= 0 && !tr.done && tr.cur >= 0;
                if (descending) tr.node_step(M.sc);
                const unsigned still = __ballot_sync(0xffffffffu, lane >= 0 && !tr.done && tr.cur >= 0);
                const unsigned leafy = __ballot_sync(0xffffffffu, lane >= 0 && !tr.done && tr.cur < 0);
                if (still == 0 || (leafy != 0 && __popc(still) < M.traceDescend)) break;
            }
            if (lane >= 0 && !tr.done && tr.cur < 0) tr.leaf_step(M.sc);
            if (lane >= 0 && tr.done) {
                const bool found = tr.hit.tri >= 0;
                int dest;
                if (pt) dest = Q_PT;
                else if (tr.anyhit || !found) dest = Q_CHAIN + M.parity;
                else {
                    const uint32_t mf = (uint32_t) __float_as_int(__ldg(&M.sc.tris[3 * (size_t) tr.hit.tri + 2].z));
                    dest = Q_WALK + (int) ((mf >> 24) & 7u);
                }
                if (lane & Q_DEFERRED) M.lm.neeOcc[lane & ~Q_DEFERRED] = found ? 1 : 0;   // the lane comes back with its BSDF-sampled ray
                else q_push_hit(M.q, dest, (uint32_t) lane, found ? tr.hit.tri : -1);
                lane = -1;
            }
            const int busy = __popc(__ballot_sync(0xffffffffu, lane >= 0));
            if (busy == 0 || (!exhausted && busy < M.traceRefill)) break;
        }
    }
}

// ---------------------------------------------------------------- replay kernel of dr_trace_rays
__global__ void k_trace_rays(const __grid_constant__ DevScene sc, const dr_ray *rays, long long n, int shadow, const unsigned int *order, dr_hit *hits) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const dr_ray r = rays[i];
    Hit h;
    const float3 o = f3(r.o[0], r.o[1], r.o[2]), d = f3(r.d[0], r.d[1], r.d[2]);
    const double rd[8] = { r.o[0], r.o[1], r.o[2], r.d[0], r.d[1], r.d[2], r.mint, r.maxt };   // the caller's ray, exactly
    bool hit = shadow ? traverse<true>(sc, o, d, r.mint, r.maxt, rd, h) : traverse<false>(sc, o, d, r.mint, r.maxt, rd, h);
    dr_hit out;
    out.prim = hit ? (int) order[h.tri] : -1;
    out.t = hit ? h.t : 0.f; out.u = hit ? h.u : 0.f; out.v = hit ? h.v : 0.f;
    hits[i] = out;
}

// persistent kernels: exactly as many CTAs as are resident at once (SMs x occupancy)
static int gridC = 0;
void trace_init() {                                          // outside any stream capture
    int dev = 0, sms = 148, bc = 4;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bc, k_trace, 128, 0);
    if (trace_ctas_per_sm() > 0) bc = std::min(bc, trace_ctas_per_sm());
    gridC = sms * std::max(bc, 1);
}

void launch_trace(const Machine &M, const LaunchCfg &lc) {
    const int need = std::max(1, (lc.nLanes + 127) / 128);
    k_trace<<<(unsigned) std::min(gridC, need), 128, 0, lc.stream>>>(M);
}
void launch_trace_rays(const DevScene &sc, const dr_ray *rays, long long n, int shadow, const unsigned int *order, dr_hit *hits, cudaStream_t stream) {
    k_trace_rays<<<(unsigned) ((n + 127) / 128), 128, 0, stream>>>(sc, rays, n, shadow, order, hits);
}
