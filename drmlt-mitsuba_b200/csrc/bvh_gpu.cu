// bvh_gpu.cu -- BVH build on the device: 63-bit Morton codes of the triangle centroids, radix sort, the binary radix tree of
// the sorted codes (one thread per inner node), bottom-up refit, and a level-by-level collapse into the 4-wide nodes of
// scene.h.  ~4 ms for 1 M triangles against ~0.6 s for the host's binned-SAH build (bvh_build.cpp); the tree is ~15 % more
// expensive to traverse (SAH cost + 15 %, C5 job - 6 %, measured), so dr_scene_create keeps the host build and dr_scene_create_ex(DR_SCENE_BVH_GPU)
// -- what a one-shot `mitsuba scene.xml` wants -- selects this one.  Like the host build it replaces the reference's SAH
// kd-tree construction (include/mitsuba/render/gkdtree.h): what matters for parity is the set of triangles, not the tree.
#include "scene.h"
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>

#define CKB(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { dr_set_error("%s: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); cudaGetLastError(); return false; } } while (0)

namespace {

constexpr int LEAF = 2;                       // triangles per leaf (the leaf code holds count - 1 in two bits)
constexpr uint32_t LEAF_REF = 0x80000000u;    // child reference: sorted triangle position | LEAF_REF, or inner node index

// order-preserving float <-> uint (for atomicMin / atomicMax on floats of either sign)
__device__ __forceinline__ uint32_t f2o(float f) { const uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__host__ __device__ __forceinline__ float o2f(uint32_t o) {
    const uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

struct Bounds { uint32_t boxLo[3], boxHi[3], cenLo[3], cenHi[3]; };    // ordered-uint encoded

// triangle boxes + the scene's box and centroid box
__global__ void k_prim_boxes(const float *P, const uint32_t *I, uint32_t n, float4 *lo, float4 *hi, Bounds *bounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float l[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, h[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    if (i < n) {
        for (int v = 0; v < 3; ++v) {
            const float *p = P + 3 * (size_t) I[3 * (size_t) i + v];
            for (int a = 0; a < 3; ++a) { l[a] = fminf(l[a], p[a]); h[a] = fmaxf(h[a], p[a]); }
        }
        lo[i] = make_float4(l[0], l[1], l[2], 0.f);
        hi[i] = make_float4(h[0], h[1], h[2], 0.f);
    }
    for (int a = 0; a < 3; ++a) {
        float bl = l[a], bh = h[a], cl = i < n ? 0.5f * (l[a] + h[a]) : FLT_MAX, ch = i < n ? cl : -FLT_MAX;
        for (int o = 16; o > 0; o >>= 1) {
            bl = fminf(bl, __shfl_xor_sync(0xffffffffu, bl, o)); bh = fmaxf(bh, __shfl_xor_sync(0xffffffffu, bh, o));
            cl = fminf(cl, __shfl_xor_sync(0xffffffffu, cl, o)); ch = fmaxf(ch, __shfl_xor_sync(0xffffffffu, ch, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&bounds->boxLo[a], f2o(bl)); atomicMax(&bounds->boxHi[a], f2o(bh));
            atomicMin(&bounds->cenLo[a], f2o(cl)); atomicMax(&bounds->cenHi[a], f2o(ch));
        }
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {     // 21 bits -> every third bit
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
__global__ void k_morton(const float4 *lo, const float4 *hi, uint32_t n, const Bounds *bounds, unsigned long long *keys, uint32_t *vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 l = lo[i], h = hi[i];
    const float c[3] = { 0.5f * (l.x + h.x), 0.5f * (l.y + h.y), 0.5f * (l.z + h.z) };
    unsigned long long q[3];
    for (int a = 0; a < 3; ++a) {
        const float c0 = o2f(bounds->cenLo[a]), ext = o2f(bounds->cenHi[a]) - c0;
        const double t = ext > 0.f ? ((double) c[a] - (double) c0) / (double) ext : 0.0;
        q[a] = (unsigned long long) fmin(2097151.0, t * 2097152.0);
    }
    keys[i] = spread21(q[0]) << 2 | spread21(q[1]) << 1 | spread21(q[2]);
    vals[i] = i;
}

// length of the common prefix of the keys at sorted positions i and j (equal keys: the positions break the tie)
__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    return a == b ? 64 + __clz(i ^ j) : __clzll((long long) (a ^ b));
}
// binary radix tree over the sorted keys, one thread per inner node (n - 1 of them; node 0 is the root)
__global__ void k_radix_tree(const unsigned long long *keys, int n, uint32_t *left, uint32_t *right, uint32_t *first, uint32_t *last,
                             uint32_t *parentInner /* [n - 1] */, uint32_t *parentLeaf /* [n] */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t L = lo == gamma ? (uint32_t) gamma | LEAF_REF : (uint32_t) gamma;
    const uint32_t R = hi == gamma + 1 ? (uint32_t) (gamma + 1) | LEAF_REF : (uint32_t) (gamma + 1);
    left[i] = L; right[i] = R; first[i] = (uint32_t) lo; last[i] = (uint32_t) hi;
    if (L & LEAF_REF) parentLeaf[gamma] = (uint32_t) i; else parentInner[gamma] = (uint32_t) i;
    if (R & LEAF_REF) parentLeaf[gamma + 1] = (uint32_t) i; else parentInner[gamma + 1] = (uint32_t) i;
    if (i == 0) parentInner[0] = 0xffffffffu;
}

// bottom-up boxes: the second thread to arrive at a node joins its children's boxes and climbs on
__global__ void k_refit(const uint32_t *order, const float4 *primLo, const float4 *primHi, int n, const uint32_t *left, const uint32_t *right,
                        const uint32_t *parentInner, const uint32_t *parentLeaf, float4 *leafLo, float4 *leafHi, float4 *nodeLo, float4 *nodeHi,
                        uint32_t *arrived) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t prim = order[s];
    leafLo[s] = primLo[prim]; leafHi[s] = primHi[prim];
    __threadfence();
    uint32_t node = parentLeaf[s];
    while (node != 0xffffffffu) {
        if (atomicAdd(arrived + node, 1u) == 0u) return;
        const uint32_t L = left[node], R = right[node];
        // (L2 loads: the sibling's box was written by another SM, and a neighbour of it may already sit in this SM's L1)
        const float4 al = __ldcg((L & LEAF_REF) ? leafLo + (L & ~LEAF_REF) : nodeLo + L), ah = __ldcg((L & LEAF_REF) ? leafHi + (L & ~LEAF_REF) : nodeHi + L);
        const float4 bl = __ldcg((R & LEAF_REF) ? leafLo + (R & ~LEAF_REF) : nodeLo + R), bh = __ldcg((R & LEAF_REF) ? leafHi + (R & ~LEAF_REF) : nodeHi + R);
        nodeLo[node] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.f);
        nodeHi[node] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
        __threadfence();
        node = parentInner[node];
    }
}

// One level of the collapse.  An item is (binary inner node, index of the 4-wide node it becomes).  The children of the 4-wide
// node are the binary node's two children, the inner one of largest surface area replaced by ITS children until there are four
// (as collapse_bvh4, bvh_build.cpp); binary subtrees of <= LEAF triangles are leaves.  Inner children get the next free
// 4-wide node and go to the next level's items.
struct Child { float lo[3], hi[3]; int code; bool inner; };
struct TreeView {
    const uint32_t *left, *right, *first, *last;
    const float4 *leafLo, *leafHi, *nodeLo, *nodeHi;
};
__device__ __forceinline__ Child make_child(const TreeView &t, uint32_t ref) {
    Child c;
    float4 l, h;
    if (ref & LEAF_REF) {
        const uint32_t pos = ref & ~LEAF_REF;
        l = t.leafLo[pos]; h = t.leafHi[pos];
        c.code = ~(int) (pos << 2); c.inner = false;
    } else {
        l = t.nodeLo[ref]; h = t.nodeHi[ref];
        const uint32_t f = t.first[ref], cnt = t.last[ref] - f + 1u;
        if (cnt <= (uint32_t) LEAF) { c.code = ~(int) ((f << 2) | (cnt - 1u)); c.inner = false; }
        else { c.code = (int) ref; c.inner = true; }
    }
    c.lo[0] = l.x; c.lo[1] = l.y; c.lo[2] = l.z; c.hi[0] = h.x; c.hi[1] = h.y; c.hi[2] = h.z;
    return c;
}
__device__ __forceinline__ float half_area(const Child &c) {
    const float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
    return dx * dy + dy * dz + dz * dx;
}
__global__ void k_collapse_level(TreeView t, const uint2 *items, uint32_t nItems, uint2 *nextItems, uint32_t *nextCount, uint32_t *nodeCount,
                                 uint32_t maxNodes, float pad, float4 *nodes4) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nItems) return;
    const uint2 item = items[k];
    Child ch[4];
    int cnt = 2;
    ch[0] = make_child(t, t.left[item.x]);
    ch[1] = make_child(t, t.right[item.x]);
    while (cnt < 4) {
        int best = -1; float bestArea = -1.f;
        for (int i = 0; i < cnt; ++i)
            if (ch[i].inner) { const float a = half_area(ch[i]); if (a > bestArea) { bestArea = a; best = i; } }
        if (best < 0) break;
        const uint32_t node = (uint32_t) ch[best].code;
        ch[best] = make_child(t, t.left[node]);
        ch[cnt++] = make_child(t, t.right[node]);
    }
    int nInner = 0;
    for (int i = 0; i < cnt; ++i) nInner += ch[i].inner ? 1 : 0;
    uint32_t base = 0, slot = 0;
    if (nInner) { base = atomicAdd(nodeCount, (uint32_t) nInner); slot = atomicAdd(nextCount, (uint32_t) nInner); }
    float v[7][4];
    for (int i = 0; i < 4; ++i) {
        int code = DR_NO_CHILD;
        if (i < cnt) {
            code = ch[i].code;
            if (ch[i].inner) {
                const uint32_t idx = base++;
                if (idx < maxNodes) nextItems[slot++] = make_uint2((uint32_t) code, idx);
                code = (int) idx;
            }
        }
        for (int a = 0; a < 3; ++a) { v[a][i] = i < cnt ? ch[i].lo[a] - pad : 1e30f; v[3 + a][i] = i < cnt ? ch[i].hi[a] + pad : 1e30f; }
        v[6][i] = __int_as_float(code);
    }
    float4 *n = nodes4 + 8 * (size_t) item.y;
    for (int r = 0; r < 7; ++r) n[r] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
    n[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// temporary device memory of a build: one allocation, handed out in 256-byte aligned pieces
struct Arena {
    char *base = nullptr; size_t size = 0, used = 0;
    ~Arena() { if (base) cudaFree(base); }
    static size_t padded(size_t bytes) { return (std::max<size_t>(bytes, 16) + 255) & ~(size_t) 255; }
    template <class T> T *take(size_t count) { T *p = (T *) (base + used); used += padded(count * sizeof(T)); return p; }
};

// triangle records of scene.h in leaf order, from the caller's arrays (what dr_scene_create's host loop does for the host build)
__global__ void k_pack_triangles(const float *P, const float *N, const uint32_t *I, const uint32_t *order, const uint32_t *triMat, const int32_t *triEm,
                                 const uint32_t *triFlags, const int32_t *matType, uint32_t n, int anySmooth, int hasUV, float4 *tris, float4 *normals) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const uint32_t prim = order[slot];
    const uint32_t i0 = I[3 * (size_t) prim], i1 = I[3 * (size_t) prim + 1], i2 = I[3 * (size_t) prim + 2];
    const float *p0 = P + 3 * (size_t) i0, *p1 = P + 3 * (size_t) i1, *p2 = P + 3 * (size_t) i2;
    const bool smooth = anySmooth && triFlags && (triFlags[prim] & DR_TRI_SMOOTH);
    const uint32_t m = triMat[prim];
    const uint32_t mf = m | ((uint32_t) matType[m] << 24) | (smooth ? 0x80000000u : 0u) |
                        (hasUV && !(triFlags && (triFlags[prim] & DR_TRI_NO_TEXCOORDS)) ? DR_MF_HAS_UV | (triFlags && (triFlags[prim] & DR_TRI_UV_TANGENTS) ? DR_MF_UV_TANGENTS : 0u) : 0u);
    tris[3 * (size_t) slot] = make_float4(p0[0], p0[1], p0[2], p1[0]);
    tris[3 * (size_t) slot + 1] = make_float4(p1[1], p1[2], p2[0], p2[1]);
    tris[3 * (size_t) slot + 2] = make_float4(p2[2], __int_as_float((int) prim), __int_as_float((int) mf), __int_as_float(triEm[prim]));
    if (anySmooth) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
        if (smooth) {
            const float *n0 = N + 3 * (size_t) i0, *n1 = N + 3 * (size_t) i1, *n2 = N + 3 * (size_t) i2;
            a = make_float4(n0[0], n0[1], n0[2], n1[0]); b = make_float4(n1[1], n1[2], n2[0], n2[1]); c = make_float4(n2[2], 0.f, 0.f, 0.f);
        }
        normals[3 * (size_t) slot] = a; normals[3 * (size_t) slot + 1] = b; normals[3 * (size_t) slot + 2] = c;
    }
}

} // namespace

void GpuScene::release() {
    for (void *p : { (void *) nodes, (void *) order, (void *) tris, (void *) normals }) if (p) cudaFree(p);
    nodes = nullptr; order = nullptr; tris = nullptr; normals = nullptr;
}

// Builds the 4-wide BVH of the scene's triangles on the current device and packs the triangle / normal records in leaf order,
// all device-resident (the four arrays of `out` are cudaMalloc'ed; the caller owns them).  False: see dr_last_error -- or, with
// `tooDeep` set, the tree exceeds the traversal stack and the caller should fall back to the host build.
bool build_scene_gpu(const dr_scene_desc *d, bool anySmooth, int stackLimit, GpuScene &out, bool *tooDeep) {
    *tooDeep = false;
    const int n = (int) d->n_triangles;
    const size_t nVerts = d->n_vertices;
    const uint32_t maxNodes = (uint32_t) n;          // a 4-wide node per binary inner node at most
    size_t tempBytes = 0;
    CKB(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, (unsigned long long *) nullptr, (unsigned long long *) nullptr,
                                        (uint32_t *) nullptr, (uint32_t *) nullptr, n, 0, 63, 0));
    float *dP, *dN; uint32_t *dI;
    float4 *primLo, *primHi, *leafLo, *leafHi, *nodeLo, *nodeHi, *nodes4;
    unsigned long long *keys, *keysSorted;
    uint2 *itemsA, *itemsB;
    uint32_t *vals, *left, *right, *first, *last, *parentInner, *parentLeaf, *arrived, *triMat, *triFlags, *counters;
    int32_t *triEm, *matType;
    Bounds *dBounds;
    void *temp;
    const size_t N4 = (size_t) n * sizeof(float4), N1 = (size_t) n * sizeof(uint32_t);
    auto layout = [&](Arena &a) {
        dP = a.take<float>(3 * nVerts); dN = a.take<float>(anySmooth ? 3 * nVerts : 4); dI = a.take<uint32_t>(3 * (size_t) n);
        primLo = a.take<float4>(n); primHi = a.take<float4>(n); leafLo = a.take<float4>(n); leafHi = a.take<float4>(n);
        nodeLo = a.take<float4>(n); nodeHi = a.take<float4>(n);
        keys = a.take<unsigned long long>(n); keysSorted = a.take<unsigned long long>(n);
        itemsA = a.take<uint2>(n); itemsB = a.take<uint2>(n);
        vals = a.take<uint32_t>(n); left = a.take<uint32_t>(n); right = a.take<uint32_t>(n); first = a.take<uint32_t>(n); last = a.take<uint32_t>(n);
        parentInner = a.take<uint32_t>(n); parentLeaf = a.take<uint32_t>(n); arrived = a.take<uint32_t>(n);
        triMat = a.take<uint32_t>(n); triFlags = a.take<uint32_t>(n); counters = a.take<uint32_t>(2);
        triEm = a.take<int32_t>(n); matType = a.take<int32_t>(d->n_materials);
        dBounds = a.take<Bounds>(1);
        temp = a.take<char>(tempBytes);
        nodes4 = a.take<float4>(8 * (size_t) maxNodes);
    };
    Arena A;
    { Arena measure; layout(measure); A.size = measure.used; measure.base = nullptr; }
    if (cudaMalloc((void **) &A.base, A.size) != cudaSuccess) { dr_set_error("build_scene_gpu: cudaMalloc of %zu bytes failed", A.size); cudaGetLastError(); return false; }
    layout(A);
    out.release();
    struct Guard { GpuScene &g; bool keep = false; ~Guard() { if (!keep) g.release(); } } guard{ out };
    if (cudaMalloc((void **) &out.order, std::max<size_t>(N1, 16)) != cudaSuccess || cudaMalloc((void **) &out.tris, 3 * N4) != cudaSuccess ||
        cudaMalloc((void **) &out.normals, anySmooth ? 3 * N4 : 16) != cudaSuccess) {
        dr_set_error("build_scene_gpu: cudaMalloc of the triangle records failed"); cudaGetLastError(); return false;
    }
    out.normalsCount = anySmooth ? 3 * (size_t) n : 0;

    CKB(cudaMemcpyAsync(dP, d->positions, 3 * nVerts * sizeof(float), cudaMemcpyHostToDevice, 0));
    CKB(cudaMemcpyAsync(dI, d->indices, 3 * N1, cudaMemcpyHostToDevice, 0));
    Bounds init;
    for (int a = 0; a < 3; ++a) { init.boxLo[a] = init.cenLo[a] = 0xffffffffu; init.boxHi[a] = init.cenHi[a] = 0u; }
    CKB(cudaMemcpyAsync(dBounds, &init, sizeof(init), cudaMemcpyHostToDevice, 0));
    const unsigned blocks = (unsigned) ((n + 255) / 256);
    k_prim_boxes<<<blocks, 256>>>(dP, dI, (uint32_t) n, primLo, primHi, dBounds);
    k_morton<<<blocks, 256>>>(primLo, primHi, (uint32_t) n, dBounds, keys, vals);
    CKB(cub::DeviceRadixSort::SortPairs(temp, tempBytes, keys, keysSorted, vals, out.order, n, 0, 63, 0));
    k_radix_tree<<<blocks, 256>>>(keysSorted, n, left, right, first, last, parentInner, parentLeaf);
    CKB(cudaMemsetAsync(arrived, 0, N1, 0));
    k_refit<<<blocks, 256>>>(out.order, primLo, primHi, n, left, right, parentInner, parentLeaf, leafLo, leafHi, nodeLo, nodeHi, arrived);
    // the per-triangle attributes travel while the tree is built
    std::vector<int32_t> types(d->n_materials);
    for (uint32_t m = 0; m < d->n_materials; ++m) types[m] = d->materials[m].type;
    CKB(cudaMemcpyAsync(triMat, d->tri_material, N1, cudaMemcpyHostToDevice, 0));
    CKB(cudaMemcpyAsync(triEm, d->tri_emitter, N1, cudaMemcpyHostToDevice, 0));
    if (d->tri_flags) CKB(cudaMemcpyAsync(triFlags, d->tri_flags, N1, cudaMemcpyHostToDevice, 0));
    if (anySmooth) CKB(cudaMemcpyAsync(dN, d->normals, 3 * nVerts * sizeof(float), cudaMemcpyHostToDevice, 0));
    CKB(cudaMemcpyAsync(matType, types.data(), types.size() * sizeof(int32_t), cudaMemcpyHostToDevice, 0));
    k_pack_triangles<<<blocks, 256>>>(dP, anySmooth ? dN : nullptr, dI, out.order, triMat, triEm, d->tri_flags ? triFlags : nullptr, matType, (uint32_t) n,
                                      anySmooth ? 1 : 0, d->texcoords ? 1 : 0, out.tris, out.normals);
    Bounds b;
    CKB(cudaMemcpy(&b, dBounds, sizeof(b), cudaMemcpyDeviceToHost));
    // child boxes are padded by a few 1e-6 of the scene extent, as in the host build (the traversal tests them with the float-cast ray)
    const float pad = 4e-6f * std::max(std::max(o2f(b.boxHi[0]) - o2f(b.boxLo[0]), o2f(b.boxHi[1]) - o2f(b.boxLo[1])),
                                       std::max(o2f(b.boxHi[2]) - o2f(b.boxLo[2]), 1e-30f));
    // collapse, level by level from the root
    TreeView tv = { left, right, first, last, leafLo, leafHi, nodeLo, nodeHi };
    const uint2 root = make_uint2(0u, 0u);
    CKB(cudaMemcpyAsync(itemsA, &root, sizeof(root), cudaMemcpyHostToDevice, 0));
    uint32_t hostCounters[2] = { 0u, 1u };              // { items of the next level, 4-wide nodes allocated }
    CKB(cudaMemcpyAsync(counters, hostCounters, sizeof(hostCounters), cudaMemcpyHostToDevice, 0));
    uint32_t nItems = 1;
    int levels = 0;
    while (nItems > 0) {
        ++levels;
        if (3 * levels + 1 >= stackLimit) { *tooDeep = true; cudaDeviceSynchronize(); return false; }
        k_collapse_level<<<(nItems + 127) / 128, 128>>>(tv, itemsA, nItems, itemsB, counters, counters + 1, maxNodes, pad, nodes4);
        CKB(cudaMemcpy(hostCounters, counters, sizeof(hostCounters), cudaMemcpyDeviceToHost));
        nItems = hostCounters[0];
        if (hostCounters[1] > maxNodes) { dr_set_error("build_scene_gpu: node count %u exceeds the bound %u", hostCounters[1], maxNodes); return false; }
        CKB(cudaMemsetAsync(counters, 0, sizeof(uint32_t), 0));
        std::swap(itemsA, itemsB);
    }
    out.nNodes = hostCounters[1];
    const size_t nodeBytes = 8 * (size_t) out.nNodes * sizeof(float4);
    if (cudaMalloc((void **) &out.nodes, nodeBytes) != cudaSuccess) { dr_set_error("build_scene_gpu: cudaMalloc of %zu bytes failed", nodeBytes); cudaGetLastError(); return false; }
    CKB(cudaMemcpy(out.nodes, nodes4, nodeBytes, cudaMemcpyDeviceToDevice));
    CKB(cudaDeviceSynchronize());
    CKB(cudaGetLastError());
    out.stackBound = 3 * levels + 1;
    guard.keep = true;
    return true;
}
