// bvh_build.cpp -- host-side binned-SAH BVH build (binary tree, subtrees built by a thread pool, then collapsed to
// 4-wide nodes), flattened into the node layout of scene.h.  Replaces the reference's SAH kd-tree construction (include/mitsuba/render/gkdtree.h)
// for this path; what matters for parity is only the set of triangles, not the tree.
#include "scene.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <atomic>
#include <thread>

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = std::numeric_limits<float>::infinity(); hi[a] = -lo[a]; } }
    void grow(const Box &b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
};

int as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
float as_float(int i) { float f; memcpy(&f, &i, 4); return f; }

} // namespace

void build_bvh(const float *P, const uint32_t *I, uint32_t nTris, BuiltBVH &out) {
    const int NB = 16;
    const int LEAF = getenv("DRMLT_BVH_LEAF") ? std::max(1, std::min(4, atoi(getenv("DRMLT_BVH_LEAF")))) : 2;   // triangles per leaf (the leaf code holds count - 1 in two bits)
    std::vector<Box> box(nTris);
    std::vector<float> cen(3 * (size_t) nTris);
    out.order.resize(nTris);
    for (uint32_t i = 0; i < nTris; ++i) {
        out.order[i] = i;
        Box b; b.reset();
        for (int v = 0; v < 3; ++v)
            for (int a = 0; a < 3; ++a) {
                float x = P[3 * (size_t) I[3 * (size_t) i + v] + a];
                b.lo[a] = std::min(b.lo[a], x); b.hi[a] = std::max(b.hi[a], x);
            }
        box[i] = b;
        for (int a = 0; a < 3; ++a) cen[3 * (size_t) i + a] = 0.5f * (b.lo[a] + b.hi[a]);
    }
    out.nodes.clear();
    out.rootIsLeaf = nTris <= (uint32_t) LEAF;
    if (out.rootIsLeaf) {
        // one pseudo node whose child0 is the leaf and child1 an empty box
        Box b; b.reset();
        for (uint32_t i = 0; i < nTris; ++i) b.grow(box[i]);
        float inf = std::numeric_limits<float>::infinity();
        out.nodes.push_back(make_float4(b.lo[0], b.lo[1], b.lo[2], b.hi[0]));
        out.nodes.push_back(make_float4(b.hi[1], b.hi[2], inf, inf));
        out.nodes.push_back(make_float4(inf, -inf, -inf, -inf));
        out.nodes.push_back(make_float4(as_float(~(int) ((0u << 2) | (nTris ? nTris - 1 : 0))), as_float(~0), 0.f, 0.f));
        out.rootIsLeaf = 0;   // handled uniformly through the pseudo node
        return;
    }
    // Child boxes are padded by a few 1e-6 of the scene extent: the traversal tests them with the float-cast ray,
    // whose origin / direction differ from the double ray by one float ulp.
    Box sceneBox; sceneBox.reset();
    for (uint32_t i = 0; i < nTris; ++i) sceneBox.grow(box[i]);
    const float pad = 4e-6f * std::max(std::max(sceneBox.hi[0] - sceneBox.lo[0], sceneBox.hi[1] - sceneBox.lo[1]),
                                       std::max(sceneBox.hi[2] - sceneBox.lo[2], 1e-30f));
    struct Range { int node; uint32_t first, count; int depth; };
    // Binned-SAH split of one range: partitions out.order[first, first + count) in place, fills node `r.node` of `nodes` and
    // returns the child ranges that need nodes of their own (allocated at the end of `nodes`).  Ranges are disjoint, so
    // subtrees can be built concurrently as long as each thread appends to its own node array.
    auto split = [&](const Range &r, std::vector<float4> &nodes, std::vector<Range> &work, int &maxDepth) {
        float clo[3], chi[3];
        for (int a = 0; a < 3; ++a) { clo[a] = std::numeric_limits<float>::infinity(); chi[a] = -clo[a]; }
        for (uint32_t i = r.first; i < r.first + r.count; ++i) {
            const float *c = &cen[3 * (size_t) out.order[i]];
            for (int a = 0; a < 3; ++a) { clo[a] = std::min(clo[a], c[a]); chi[a] = std::max(chi[a], c[a]); }
        }
        int bestAxis = -1, bestSplit = -1; float bestCost = std::numeric_limits<float>::infinity();
        for (int a = 0; a < 3; ++a) {
            float ext = chi[a] - clo[a];
            if (!(ext > 0.f)) continue;
            uint32_t cnt[NB] = { 0 }; Box bb[NB];
            for (int b = 0; b < NB; ++b) bb[b].reset();
            float scale = NB / ext;
            for (uint32_t i = r.first; i < r.first + r.count; ++i) {
                uint32_t p = out.order[i];
                int b = std::min(NB - 1, (int) ((cen[3 * (size_t) p + a] - clo[a]) * scale));
                cnt[b]++; bb[b].grow(box[p]);
            }
            float rArea[NB]; uint32_t rCnt[NB];
            Box acc; acc.reset(); uint32_t c2 = 0;
            for (int b = NB - 1; b > 0; --b) { acc.grow(bb[b]); c2 += cnt[b]; rCnt[b] = c2; rArea[b] = c2 ? acc.area() : 0.f; }
            acc.reset(); uint32_t c1 = 0;
            for (int b = 0; b < NB - 1; ++b) {
                acc.grow(bb[b]); c1 += cnt[b];
                if (c1 == 0 || rCnt[b + 1] == 0) continue;
                float cost = c1 * acc.area() + rCnt[b + 1] * rArea[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = b; }
            }
        }
        uint32_t mid;
        if (bestAxis >= 0) {
            float ext = chi[bestAxis] - clo[bestAxis], scale = NB / ext;
            uint32_t *beg = out.order.data() + r.first, *end = beg + r.count;
            uint32_t *m = std::partition(beg, end, [&](uint32_t p) {
                int b = std::min(NB - 1, (int) ((cen[3 * (size_t) p + bestAxis] - clo[bestAxis]) * scale));
                return b <= bestSplit;
            });
            mid = (uint32_t) (m - out.order.data());
        } else {
            mid = r.first + r.count / 2;
        }
        if (mid == r.first || mid == r.first + r.count) mid = r.first + r.count / 2;
        uint32_t firsts[2] = { r.first, mid }, counts[2] = { mid - r.first, r.first + r.count - mid };
        Box cb[2]; int code[2];
        for (int c = 0; c < 2; ++c) {
            cb[c].reset();
            for (uint32_t i = firsts[c]; i < firsts[c] + counts[c]; ++i) cb[c].grow(box[out.order[i]]);
            if (counts[c] <= (uint32_t) LEAF) {
                code[c] = ~(int) ((firsts[c] << 2) | (counts[c] - 1));
            } else {
                code[c] = (int) (nodes.size() / 4);
                nodes.resize(nodes.size() + 4);
                work.push_back({ code[c], firsts[c], counts[c], r.depth + 1 });
                maxDepth = std::max(maxDepth, r.depth + 1);
            }
            for (int a = 0; a < 3; ++a) { cb[c].lo[a] -= pad; cb[c].hi[a] += pad; }
        }
        float4 *n = &nodes[4 * (size_t) r.node];
        n[0] = make_float4(cb[0].lo[0], cb[0].lo[1], cb[0].lo[2], cb[0].hi[0]);
        n[1] = make_float4(cb[0].hi[1], cb[0].hi[2], cb[1].lo[0], cb[1].lo[1]);
        n[2] = make_float4(cb[1].lo[2], cb[1].hi[0], cb[1].hi[1], cb[1].hi[2]);
        n[3] = make_float4(as_float(code[0]), as_float(code[1]), 0.f, 0.f);
    };
    // Phase A (serial): the top of the tree, until the open ranges are small enough to be handed out as subtree tasks.
    out.nodes.resize(4);
    out.maxDepth = 1;
    const uint32_t taskSize = std::max<uint32_t>(4096, nTris / 64);
    std::vector<Range> work, tasks;
    work.push_back({ 0, 0, nTris, 1 });
    while (!work.empty()) {
        Range r = work.back(); work.pop_back();
        if (r.count <= taskSize && r.node != 0) { tasks.push_back(r); continue; }
        split(r, out.nodes, work, out.maxDepth);
    }
    // Phase B (parallel): every task builds its subtree into a private node array whose node 0 is the subtree root.
    struct Sub { std::vector<float4> nodes; int maxDepth = 1; };
    std::vector<Sub> subs(tasks.size());
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= tasks.size()) break;
            Sub &sb = subs[k];
            sb.nodes.resize(4);
            std::vector<Range> w;
            w.push_back({ 0, tasks[k].first, tasks[k].count, tasks[k].depth });
            sb.maxDepth = tasks[k].depth;
            while (!w.empty()) { Range r = w.back(); w.pop_back(); split(r, sb.nodes, w, sb.maxDepth); }
        }
    };
    const unsigned nThreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nThreads && t < tasks.size(); ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    // Stitch: the subtree root goes into the node reserved for it, its other nodes are appended; inner child codes move along.
    for (size_t k = 0; k < tasks.size(); ++k) {
        const Sub &sb = subs[k];
        const int offset = (int) (out.nodes.size() / 4) - 1;      // private node i >= 1 -> global node offset + i
        const size_t nLocal = sb.nodes.size() / 4;
        out.nodes.resize(out.nodes.size() + 4 * (nLocal - 1));
        for (size_t i = 0; i < nLocal; ++i) {
            float4 n3 = sb.nodes[4 * i + 3];
            int c0 = as_int(n3.x), c1 = as_int(n3.y);
            if (c0 >= 0) c0 += offset;
            if (c1 >= 0) c1 += offset;
            n3.x = as_float(c0); n3.y = as_float(c1);
            float4 *dst = &out.nodes[4 * (i == 0 ? (size_t) tasks[k].node : (size_t) (offset + (int) i))];
            dst[0] = sb.nodes[4 * i]; dst[1] = sb.nodes[4 * i + 1]; dst[2] = sb.nodes[4 * i + 2]; dst[3] = n3;
        }
        out.maxDepth = std::max(out.maxDepth, sb.maxDepth);
    }
}

// ------------------------------------------------------------------ BVH2 -> BVH4
// Collapse: the children of a 4-wide node are the grandchildren of the binary node, expanding the inner child of largest
// surface area first (three quarters of the binary nodes disappear; a ray does half as many dependent node fetches).
// 128-byte node = 8 x float4, structure of arrays over the four children:
//   lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] child[4] (int bits: >= 0 inner node, < 0 leaf code, DR_NO_CHILD empty) -
void collapse_bvh4(BuiltBVH &bvh) {
    const std::vector<float4> n2 = bvh.nodes;
    std::vector<float4> n4;
    struct Child { float lo[3], hi[3]; int code; };
    auto children2 = [&](int node, Child out[2]) {
        const float4 *n = &n2[4 * (size_t) node];
        out[0] = { { n[0].x, n[0].y, n[0].z }, { n[0].w, n[1].x, n[1].y }, as_int(n[3].x) };
        out[1] = { { n[1].z, n[1].w, n[2].x }, { n[2].y, n[2].z, n[2].w }, as_int(n[3].y) };
    };
    auto area = [](const Child &c) {
        const float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
        return dx * dy + dy * dz + dz * dx;
    };
    struct Todo { int node2, node4, depth; };
    std::vector<Todo> work;
    n4.resize(8);
    work.push_back({ 0, 0, 1 });
    int maxDepth = 1;
    while (!work.empty()) {
        const Todo t = work.back(); work.pop_back();
        maxDepth = std::max(maxDepth, t.depth);
        Child ch[4];
        int k = 2;
        children2(t.node2, ch);
        while (k < 4) {
            int best = -1; float bestArea = -1.f;
            for (int i = 0; i < k; ++i)
                if (ch[i].code >= 0 && std::isfinite(ch[i].lo[0]) && area(ch[i]) > bestArea) { bestArea = area(ch[i]); best = i; }
            if (best < 0) break;
            Child two[2];
            children2(ch[best].code, two);
            ch[best] = two[0];
            ch[k++] = two[1];
        }
        float v[7][4];
        for (int i = 0; i < 4; ++i) {
            const bool empty = i >= k || (ch[i].code < 0 && ch[i].code == ~0 && !std::isfinite(ch[i].lo[0]));   // pseudo-root's empty child
            int code = empty ? DR_NO_CHILD : ch[i].code;
            if (!empty && code >= 0) {
                const int idx = (int) (n4.size() / 8);
                n4.resize(n4.size() + 8);
                work.push_back({ code, idx, t.depth + 1 });
                code = idx;
            }
            for (int a = 0; a < 3; ++a) { v[a][i] = empty ? 1e30f : ch[i].lo[a]; v[3 + a][i] = empty ? 1e30f : ch[i].hi[a]; }
            v[6][i] = as_float(code);
        }
        float4 *n = &n4[8 * (size_t) t.node4];
        for (int r = 0; r < 7; ++r) n[r] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
        n[7] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    bvh.nodes.swap(n4);
    bvh.maxDepth = 3 * maxDepth + 1;          // a step pushes up to three children
}
