// k_trace.cu -- ray traversal stages of the wavefront machine (machine.cuh): one thread per queued ray,
// float32 BVH traversal only (traverse.cuh); small register footprint, so many warps are resident to hide the
// node / triangle gather latency.  A closest hit is routed to the walk queue of the BSDF model of the triangle it
// found (the model is packed into the triangle record at scene creation), a miss or a shadow-ray result to the
// chain queue (MMLT), everything to the path-tracer queue for technique=path.
#include "machine.cuh"

template <bool SHADOW>
__global__ void __launch_bounds__(256)
k_trace(const __grid_constant__ Machine M) {
    const int qid = (SHADOW ? Q_RAYS : Q_RAYC) + M.parity;
    const uint32_t cnt = M.q.count[qid];
    const uint32_t *items = M.q.items + (size_t) qid * M.q.n;
    if (blockIdx.x == 0 && threadIdx.x == 0 && cnt) atomicAdd(&M.counters[ST_RAYS], (unsigned long long) cnt);
    const bool pt = M.pc.technique != DR_TECH_MMLT;
    for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < cnt; qi += gridDim.x * blockDim.x) {
        const int lane = (int) items[qi];
        const float4 a = M.lm.ray[2 * (size_t) lane], b = M.lm.ray[2 * (size_t) lane + 1];
        Hit h;
        const bool found = traverse<SHADOW>(M.sc, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, b.w, M.lm.rayd + 8 * (size_t) lane, h);
        M.lm.hit[lane] = make_float4(h.t, h.u, h.v, __int_as_float(found ? h.tri : -1));
        int dest;
        if (pt) dest = Q_PT;
        else if (SHADOW || !found) dest = Q_CHAIN + M.parity;
        else {
            const uint32_t mf = (uint32_t) __float_as_int(__ldg(&M.sc.tris[3 * (size_t) h.tri + 2].z));
            dest = Q_WALK + (int) ((mf >> 24) & 3u);
        }
        q_push(M.q, dest, (uint32_t) lane);
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

// start of a round: empty the queues this round produces into (next-parity ray / chain queues, and the
// in-round walk / connect / path-tracer queues)
__global__ void k_round_begin(uint32_t *count, int parity) {
    const int t = threadIdx.x;
    if (t < Q_COUNT) {
        const bool nextParity = (t < Q_WALK) && ((t & 1) == (parity ^ 1));
        if (nextParity || t >= Q_WALK) count[t] = 0;
    }
}

void launch_trace(const Machine &M, const LaunchCfg &lc) {
    k_round_begin<<<1, 32, 0, lc.stream>>>(M.q.count, M.parity);
    const unsigned g = (unsigned) std::max(1, std::min((lc.nLanes + 255) / 256, 148 * 8));
    k_trace<false><<<g, 256, 0, lc.stream>>>(M);
    k_trace<true><<<g, 256, 0, lc.stream>>>(M);
}
void launch_trace_rays(const DevScene &sc, const dr_ray *rays, long long n, int shadow, const unsigned int *order, dr_hit *hits, cudaStream_t stream) {
    k_trace_rays<<<(unsigned) ((n + 127) / 128), 128, 0, stream>>>(sc, rays, n, shadow, order, hits);
}
