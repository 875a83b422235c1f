// traverse.cuh -- BVH2 ray traversal: closest hit and any hit.
//
// One ray per thread, ordered while-while traversal with a short per-thread stack; every node
// visit is four 16-byte loads (64 B), every triangle test three (48 B), through the read-only
// path (ld.global.nc).  Replaces ShapeKDTree::rayIntersect + rayIntersectHavran + TriAccel
// (src/librender/skdtree.cpp:111-215, include/mitsuba/render/sahkdtree3.h:179-320,
// include/mitsuba/render/triaccel.h:91-157) for this path: same closest hit in [mint, maxt]
// (inclusive), same adaptive ray epsilon.
#pragma once
#include "scene.h"

struct Hit {
    float t, u, v;
    int tri;        // leaf-order triangle index, -1 = miss
};

#define DR_STACK 48

DR_D float4 ldg4(const float4 *p) { return __ldg(p); }

// skdtree.cpp:126-129: a ray whose mint is exactly Epsilon gets mint scaled by the origin's magnitude
DR_D float adaptive_mint(const DevScene &sc, float3 o, float mint) {
    if (mint == sc.epsilon)
        mint *= fmaxf(fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z)), sc.epsilon);
    return mint;
}

DR_D bool tri_test(const float4 t0, const float4 t1, const float4 t2, float3 o, float3 d,
                   float tmin, float tmax, float &t, float &u, float &v) {
    const float3 p0 = f3(t0.x, t0.y, t0.z), e1 = f3(t0.w, t1.x, t1.y) - p0, e2 = f3(t1.z, t1.w, t2.x) - p0;
    const float3 pvec = cross(d, e2);
    const float det = dot(e1, pvec);
    if (det == 0.f) return false;
    const float inv = 1.0f / det;
    const float3 tvec = o - p0;
    u = dot(tvec, pvec) * inv;
    const float3 qvec = cross(tvec, e1);
    v = dot(d, qvec) * inv;
    t = dot(e2, qvec) * inv;
    return u >= 0.f && v >= 0.f && u + v <= 1.0f && t >= tmin && t <= tmax;
}

template <bool ANYHIT>
DR_D bool traverse(const DevScene &sc, float3 o, float3 d, float tmin, float tmax, Hit &hit, uint32_t *nodeVisits = nullptr) {
    hit.tri = -1;
    if (!(tmax > tmin)) return false;
    // zero direction components: clamp so that 1/d stays finite (no 0 * inf = NaN in the slab test)
    const float3 inv = f3(1.0f / (fabsf(d.x) > 1e-20f ? d.x : copysignf(1e-20f, d.x)), 1.0f / (fabsf(d.y) > 1e-20f ? d.y : copysignf(1e-20f, d.y)),
                          1.0f / (fabsf(d.z) > 1e-20f ? d.z : copysignf(1e-20f, d.z)));
    const float3 oi = f3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
    int stack[DR_STACK];
    int sp = 0;
    int cur = 0;   // inner node index, or leaf code (< 0)
    while (true) {
        if (cur >= 0) {
            const float4 *n = sc.nodes + 4 * (size_t) cur;
            const float4 n0 = ldg4(n), n1 = ldg4(n + 1), n2 = ldg4(n + 2), n3 = ldg4(n + 3);
            if (nodeVisits) ++*nodeVisits;
            // child 0: lo = (n0.x n0.y n0.z) hi = (n0.w n1.x n1.y); child 1: lo = (n1.z n1.w n2.x) hi = (n2.y n2.z n2.w)
            float a0 = n0.x * inv.x - oi.x, b0 = n0.w * inv.x - oi.x;
            float a1 = n0.y * inv.y - oi.y, b1 = n1.x * inv.y - oi.y;
            float a2 = n0.z * inv.z - oi.z, b2 = n1.y * inv.z - oi.z;
            float near0 = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fmaxf(fminf(a2, b2), tmin));
            float far0 = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fminf(fmaxf(a2, b2), tmax));
            a0 = n1.z * inv.x - oi.x; b0 = n2.y * inv.x - oi.x;
            a1 = n1.w * inv.y - oi.y; b1 = n2.z * inv.y - oi.y;
            a2 = n2.x * inv.z - oi.z; b2 = n2.w * inv.z - oi.z;
            float near1 = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fmaxf(fminf(a2, b2), tmin));
            float far1 = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fminf(fmaxf(a2, b2), tmax));
            // conservative: widen by a few ulps so that float slab rounding never culls a true hit
            const bool h0 = near0 <= far0 * 1.0000004f, h1 = near1 <= far1 * 1.0000004f;
            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool swap = near1 < near0;
                stack[sp++] = swap ? c0 : c1;
                cur = swap ? c1 : c0;
                continue;
            } else if (h0) { cur = c0; continue; }
            else if (h1) { cur = c1; continue; }
        } else {
            const int code = ~cur;
            const int first = code >> 2, count = (code & 3) + 1;
            for (int i = 0; i < count; ++i) {
                const float4 *tp = sc.tris + 3 * (size_t) (first + i);
                const float4 t0 = ldg4(tp), t1 = ldg4(tp + 1), t2 = ldg4(tp + 2);
                float t, u, v;
                if (tri_test(t0, t1, t2, o, d, tmin, tmax, t, u, v)) {
                    hit.t = t; hit.u = u; hit.v = v; hit.tri = first + i;
                    if (ANYHIT) return true;
                    tmax = t;
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return hit.tri >= 0;
}
