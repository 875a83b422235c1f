// traverse.cuh -- BVH ray traversal (4-wide nodes; binary with -DDR_BVH2): closest hit and any hit.
//
// One ray per thread, ordered while-while traversal with a short per-thread stack; every node
// visit is seven 16-byte loads (six box planes x four children + the four child codes), every
// triangle test three (48 B), through the read-only path (ld.global.nc).  Replaces ShapeKDTree::rayIntersect + rayIntersectHavran + TriAccel
// (src/librender/skdtree.cpp:111-215, include/mitsuba/render/sahkdtree3.h:179-320,
// include/mitsuba/render/triaccel.h:91-157) for this path: same closest hit in [mint, maxt]
// (inclusive), same adaptive ray epsilon.  Float32 arithmetic only PRUNES: boxes are tested conservatively and
// a triangle that passes the float test within a tolerance band is decided by a double-precision test on the
// un-rounded ray, so hit / miss decisions do not depend on float rounding.
#pragma once
#include "scene.h"

struct Hit {
    float t, u, v;
    int tri;        // leaf-order triangle index, -1 = miss
};

#define DR_STACK 64
#ifdef DR_BVH4
#define DR_NODE_F4 8         /* float4 per node */
#else
#define DR_NODE_F4 4
#endif
#define DR_TRI_TOL 2e-3f     /* candidate band of the float32 triangle test (barycentric units) */

DR_D float4 ldg4(const float4 *p) { return __ldg(p); }

// skdtree.cpp:126-129: a ray whose mint is exactly Epsilon gets mint scaled by the origin's magnitude
DR_D float adaptive_mint(const DevScene &sc, float3 o, float mint) {
    if (mint == sc.epsilon)
        mint *= fmaxf(fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z)), sc.epsilon);
    return mint;
}

// Float32 triangle test with a tolerance band: 0 = miss, 1 = clear hit (inside the triangle and the ray interval by
// more than the band: float rounding cannot change the decision, nor the order against the best hit so far),
// 2 = borderline (decided in double by tri_verify).
DR_D int tri_classify(const float4 t0, const float4 t1, const float4 t2, float3 o, float3 d, float tmin, float tmax, float &t, float &u, float &v) {
    const float3 p0 = f3(t0.x, t0.y, t0.z), e1 = f3(t0.w, t1.x, t1.y) - p0, e2 = f3(t1.z, t1.w, t2.x) - p0;
    const float3 pvec = cross(d, e2);
    const float det = dot(e1, pvec);
    if (det == 0.f) return 0;
    const float inv = 1.0f / det;
    const float3 tvec = o - p0;
    u = dot(tvec, pvec) * inv;
    const float3 qvec = cross(tvec, e1);
    v = dot(d, qvec) * inv;
    t = dot(e2, qvec) * inv;
    const float tolT = 1e-5f + 1e-5f * fabsf(t);
    if (!(u >= -DR_TRI_TOL && v >= -DR_TRI_TOL && u + v <= 1.0f + DR_TRI_TOL && t >= tmin - tolT && t <= tmax + tolT)) return 0;
    const bool clear = u >= DR_TRI_TOL && v >= DR_TRI_TOL && u + v <= 1.0f - DR_TRI_TOL && t >= tmin + tolT && t <= tmax - tolT;
    return clear ? 1 : 2;
}
// ... and the deciding test in double on the un-rounded ray (o, d, tmin, tmax = rd[0..7]): the hit / miss decision, t
// and the barycentrics are those of the reference's double-precision triangle test (triaccel.h:91-157: inclusive
// [mint, maxt], u, v >= 0, u + v <= 1), independent of float32 rounding in the traversal.
DR_D bool tri_verify(const float4 t0, const float4 t1, const float4 t2, const double *rd, double tmax, double &t, double &u, double &v) {
    const double p0x = t0.x, p0y = t0.y, p0z = t0.z;
    const double e1x = (double) t0.w - p0x, e1y = (double) t1.x - p0y, e1z = (double) t1.y - p0z;
    const double e2x = (double) t1.z - p0x, e2y = (double) t1.w - p0y, e2z = (double) t2.x - p0z;
    const double ox = rd[0], oy = rd[1], oz = rd[2], dx = rd[3], dy = rd[4], dz = rd[5];
    const double pvx = dy * e2z - dz * e2y, pvy = dz * e2x - dx * e2z, pvz = dx * e2y - dy * e2x;
    const double det = e1x * pvx + e1y * pvy + e1z * pvz;
    if (det == 0.0) return false;
    const double inv = 1.0 / det;
    const double tvx = ox - p0x, tvy = oy - p0y, tvz = oz - p0z;
    u = (tvx * pvx + tvy * pvy + tvz * pvz) * inv;
    const double qx = tvy * e1z - tvz * e1y, qy = tvz * e1x - tvx * e1z, qz = tvx * e1y - tvy * e1x;
    v = (dx * qx + dy * qy + dz * qz) * inv;
    t = (e2x * qx + e2y * qy + e2z * qz) * inv;
    return u >= 0.0 && v >= 0.0 && u + v <= 1.0 && t >= rd[6] && t <= tmax;
}

// One ray's traversal state; step() processes ONE inner node or ONE leaf, so that a persistent kernel can
// interleave "fetch a new ray for the lanes that finished" with the traversal of the others.
struct Traversal {
    bool anyhit;              // shadow ray: the first hit ends the traversal
    float3 o, d, inv, oi;
    float tmin, tmax;
    const double *rd;         // un-rounded ray: o, d, tmin, tmax
    double bestT;
    bool bestExact;
    Hit hit;
    int *stack;               // DR_STACK entries of thread-private memory, owned by the caller
    int sp, cur;              // cur: inner node index, or leaf code (< 0)
    bool done;

    DR_D void begin(int *stack_, bool anyhit_, float3 o_, float3 d_, float tmin_, float tmax_, const double *rd_) {
        stack = stack_; anyhit = anyhit_;
        o = o_; d = d_; tmin = tmin_; tmax = tmax_; rd = rd_;
        hit.tri = -1; hit.t = hit.u = hit.v = 0.f;
        bestT = 0.0; bestExact = true;                // no hit yet: the limit of the exact test is rd[7]
        // zero direction components: clamp so that 1/d stays finite (no 0 * inf = NaN in the slab test)
        inv = f3(1.0f / (fabsf(d.x) > 1e-20f ? d.x : copysignf(1e-20f, d.x)), 1.0f / (fabsf(d.y) > 1e-20f ? d.y : copysignf(1e-20f, d.y)),
                 1.0f / (fabsf(d.z) > 1e-20f ? d.z : copysignf(1e-20f, d.z)));
        oi = f3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
        sp = 0; cur = 0;
        done = !(tmax > tmin);
    }
    DR_D void pop() { if (sp == 0) done = true; else cur = stack[--sp]; }
    DR_D void step(const DevScene &sc) { if (cur >= 0) node_step(sc); else leaf_step(sc); }
#ifdef DR_BVH4
    // 4-wide node: seven 16-byte loads (six box planes x four children, four child codes), four independent slab tests,
    // the hit children sorted by entry distance: the nearest is visited next, the others are pushed far to near.
    DR_D void node_step(const DevScene &sc) {
        const float4 *n = sc.nodes + 8 * (size_t) cur;
        const float4 lx = ldg4(n), ly = ldg4(n + 1), lz = ldg4(n + 2), hx = ldg4(n + 3), hy = ldg4(n + 4), hz = ldg4(n + 5), cc = ldg4(n + 6);
        float key[4]; int code[4];
#define DR_SLAB(i, c)                                                                                         \
        {                                                                                                     \
            const float a0 = lx.c * inv.x - oi.x, b0 = hx.c * inv.x - oi.x;                                   \
            const float a1 = ly.c * inv.y - oi.y, b1 = hy.c * inv.y - oi.y;                                   \
            const float a2 = lz.c * inv.z - oi.z, b2 = hz.c * inv.z - oi.z;                                   \
            const float nr = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fmaxf(fminf(a2, b2), tmin));          \
            const float fr = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fminf(fmaxf(a2, b2), tmax));          \
            code[i] = __float_as_int(cc.c);                                                                   \
            key[i] = (nr <= fr * 1.000002f && code[i] != DR_NO_CHILD) ? nr : INFINITY;                        \
        }
        DR_SLAB(0, x) DR_SLAB(1, y) DR_SLAB(2, z) DR_SLAB(3, w)
#undef DR_SLAB
#define DR_CSWAP(i, j) { if (key[j] < key[i]) { const float tk = key[i]; key[i] = key[j]; key[j] = tk; const int tc = code[i]; code[i] = code[j]; code[j] = tc; } }
        DR_CSWAP(0, 1) DR_CSWAP(2, 3) DR_CSWAP(0, 2) DR_CSWAP(1, 3) DR_CSWAP(1, 2)
#undef DR_CSWAP
        if (key[0] == INFINITY) { pop(); return; }
        if (key[3] != INFINITY) stack[sp++] = code[3];
        if (key[2] != INFINITY) stack[sp++] = code[2];
        if (key[1] != INFINITY) stack[sp++] = code[1];
        cur = code[0];
    }
#else
    DR_D void node_step(const DevScene &sc) {
        {
            const float4 *n = sc.nodes + 4 * (size_t) cur;
            const float4 n0 = ldg4(n), n1 = ldg4(n + 1), n2 = ldg4(n + 2), n3 = ldg4(n + 3);
            // child 0: lo = (n0.x n0.y n0.z) hi = (n0.w n1.x n1.y); child 1: lo = (n1.z n1.w n2.x) hi = (n2.y n2.z n2.w)
            float a0 = n0.x * inv.x - oi.x, b0 = n0.w * inv.x - oi.x;
            float a1 = n0.y * inv.y - oi.y, b1 = n1.x * inv.y - oi.y;
            float a2 = n0.z * inv.z - oi.z, b2 = n1.y * inv.z - oi.z;
            const float near0 = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fmaxf(fminf(a2, b2), tmin));
            const float far0 = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fminf(fmaxf(a2, b2), tmax));
            a0 = n1.z * inv.x - oi.x; b0 = n2.y * inv.x - oi.x;
            a1 = n1.w * inv.y - oi.y; b1 = n2.z * inv.y - oi.y;
            a2 = n2.x * inv.z - oi.z; b2 = n2.w * inv.z - oi.z;
            const float near1 = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fmaxf(fminf(a2, b2), tmin));
            const float far1 = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fminf(fmaxf(a2, b2), tmax));
            // conservative: the boxes are padded at upload and the comparison is widened by a few ulps, so that float
            // slab rounding (of the test and of the float-cast ray) never culls a true hit
            const bool h0 = near0 <= far0 * 1.000002f, h1 = near1 <= far1 * 1.000002f;
            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool swap = near1 < near0;
                stack[sp++] = swap ? c0 : c1;             // depth < DR_STACK is checked when the scene is created
                cur = swap ? c1 : c0;
            } else if (h0) cur = c0;
            else if (h1) cur = c1;
            else pop();
        }
    }
#endif
    DR_D void leaf_step(const DevScene &sc) {
        const int code = ~cur;
        const int first = code >> 2, count = (code & 3) + 1;
        for (int i = 0; i < count; ++i) {
            const float4 *tp = sc.tris + 3 * (size_t) (first + i);
            const float4 t0 = ldg4(tp), t1 = ldg4(tp + 1), t2 = ldg4(tp + 2);
            float tf, uf, vf;
            const int cls = tri_classify(t0, t1, t2, o, d, tmin, tmax, tf, uf, vf);
            if (cls == 1) {
                hit.t = tf; hit.u = uf; hit.v = vf; hit.tri = first + i;
                if (anyhit) { done = true; return; }
                bestT = (double) tf; bestExact = false;
                tmax = tf;
            } else if (cls == 2) {                       // rare: decide in double on the un-rounded ray
                double t, u, v;
                if (!anyhit && !bestExact && hit.tri >= 0 && tf >= tmax - (2e-5f + 2e-5f * fabsf(tf))) {
                    // the best hit so far was accepted on its float t, which is too close to this candidate's: make it exact
                    const float4 *bp = sc.tris + 3 * (size_t) hit.tri;
                    double bu, bv;
                    if (!tri_verify(ldg4(bp), ldg4(bp + 1), ldg4(bp + 2), rd, rd[7], bestT, bu, bv)) bestT = (double) hit.t;
                    bestExact = true;
                }
                if (tri_verify(t0, t1, t2, rd, hit.tri >= 0 ? bestT : rd[7], t, u, v)) {
                    hit.t = (float) t; hit.u = (float) u; hit.v = (float) v; hit.tri = first + i;
                    if (anyhit) { done = true; return; }
                    bestT = t; bestExact = true;
                    tmax = __double2float_ru(t);
                }
            }
        }
        pop();
    }
};

template <bool ANYHIT>
DR_D bool traverse(const DevScene &sc, float3 o, float3 d, float tmin, float tmax, const double *rd, Hit &hit) {
    int stack[DR_STACK];
    Traversal tr;
    tr.begin(stack, ANYHIT, o, d, tmin, tmax, rd);
    while (!tr.done) tr.step(sc);
    hit = tr.hit;
    return hit.tri >= 0;
}
