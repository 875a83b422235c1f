// scene.h -- GPU-resident scene: BVH, triangles, materials, emitters, camera.
//
// Data layout in HBM (all read-only during rendering, 16-byte vector loads):
//   nodes   : 128 B / inner node of the 4-wide BVH = 8 x float4, structure of arrays over the four children:
//               lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] child[4] (int bits) -
//             child >= 0: inner node index; child < 0: leaf, ~child = (firstTri << 2) | (count - 1), count <= 2;
//             DR_NO_CHILD: empty slot.  (Built as a binary SAH tree and collapsed, bvh_build.cpp; -DDR_BVH2 keeps the
//             binary layout: 64 B = (lo0.x lo0.y lo0.z hi0.x) (hi0.y hi0.z lo1.x lo1.y) (lo1.z hi1.x hi1.y hi1.z) (child0, child1, -, -).)
//   tris    : 48 B / triangle in LEAF ORDER = 3 x float4
//               t0 = (p0.x p0.y p0.z p1.x) t1 = (p1.y p1.z p2.x p2.y) t2 = (p2.z, prim, matflags, emitter)
//             the exact float vertices (edges are formed on the fly, so the double-precision
//             re-intersection of the shading stage sees the same triangle as the reference);
//             prim = index in the caller's triangle order, matflags = material | bsdf model << 24 | has texcoords << 29 |
//             UV tangents << 30 | smooth << 31
//   uvs     : 32 B / triangle in leaf order = (u0 v0 u1 v1) (u2 v2 - -): the vertices' texture coordinates (only scenes with texcoords)
//   texels  : 16 B / texel (r g b -), all bitmap textures back to back, row y = 0 first; textures: one DevTexture each
//   normals : 48 B / triangle in leaf order (only read for smooth triangles at the closest hit)
//   em_tris : 96 B / emitter triangle in EMITTER order: p0,p1,p2,smooth + n0,n1,n2 (position sampling)
//   em_cdf  : double prefix sums of the per-emitter triangle areas (pmf.h DiscreteDistribution)
// Replaces ShapeKDTree + TriAccel (src/librender/skdtree.cpp, include/mitsuba/render/triaccel.h) for this path.
#pragma once
#ifndef DR_BVH2
#define DR_BVH4 1            /* 4-wide nodes (collapse_bvh4, bvh_build.cpp); -DDR_BVH2 keeps the binary tree */
#endif
#include "common.cuh"
#include "../../include/drmlt_b200.h"
#include <atomic>
#include <vector>
#include <string>

struct DevMaterial {      // 64 B, mirrors dr_material
    int32_t type; uint32_t flags;
    float reflectance[3], transmittance[3], eta[3], k[3];
    float alpha; uint32_t table;
};

struct DevTexture {       // 64 B (src/textures/bitmap.cpp, include/mitsuba/render/mipmap.h: MIP level 0 only on this path)
    uint32_t w, h, wrapU, wrapV;
    uint32_t nearest, pad0;
    uint64_t first;       // offset of texel (0, 0) in DevScene::texels
    double scaleU, scaleV, offU, offV;   // Texture2D::m_uvScale / m_uvOffset
};
#define DR_MF_HAS_UV      0x20000000u
#define DR_MF_UV_TANGENTS 0x40000000u

struct DevEmitter {       // 64 B
    double area, invArea;
    double pdfDiscrete;   // emitterPDF[e]
    float radiance[3];
    uint32_t firstEmTri;  // offset into em_tris
    uint32_t nTris;
    uint32_t cdfOffset;   // offset into em_cdf (nTris + 1 entries)
    uint32_t pad[4];
};

struct DevCamera {        // derived in double from the float parameters of dr_camera (perspective.cpp:126-173)
    double m[12];         // rows of the 3x4 camera-to-world matrix
    double inv[9];        // inverse of its 3x3 block, rows (the reference inverts the matrix, it does not assume a rotation)
    double pos[3], dir[3];
    double tanHalf, aspect, nearClip, farClip;   // aspect of the FULL film
    double resX, resY;    // crop size = resolution of the sensor (perspective.cpp:126-130)
    double relOffX, relOffY, relSizeX, relSizeY;   // crop window as fractions of the film (:132-135)
    double rectMinX, rectMaxX, rectMinY, rectMaxY; // m_imageRect: the crop window on the plane z = 1 (:162-169)
    double normalization;
};
// film + crop window -> the sensor's derived quantities (PerspectiveCameraImpl::configure, perspective.cpp:126-173)
inline void camera_set_window(DevCamera &dc, int filmW, int filmH, int cropX, int cropY, int cropW, int cropH) {
    dc.aspect = (double) filmW / (double) filmH;
    dc.resX = cropW; dc.resY = cropH;
    dc.relSizeX = (double) cropW / (double) filmW; dc.relSizeY = (double) cropH / (double) filmH;
    dc.relOffX = (double) cropX / (double) filmW; dc.relOffY = (double) cropY / (double) filmH;
    const double x0 = (1.0 - 2.0 * dc.relOffX) * dc.tanHalf, x1 = (1.0 - 2.0 * (dc.relOffX + dc.relSizeX)) * dc.tanHalf;
    const double y0 = (1.0 - 2.0 * dc.relOffY) * dc.tanHalf / dc.aspect, y1 = (1.0 - 2.0 * (dc.relOffY + dc.relSizeY)) * dc.tanHalf / dc.aspect;
    dc.rectMinX = x0 < x1 ? x0 : x1; dc.rectMaxX = x0 < x1 ? x1 : x0;
    dc.rectMinY = y0 < y1 ? y0 : y1; dc.rectMaxY = y0 < y1 ? y1 : y0;
    dc.normalization = 1.0 / ((dc.rectMaxX - dc.rectMinX) * (dc.rectMaxY - dc.rectMinY));
}

struct DevScene {
    const float4 *nodes;
    const float4 *tris;
    const float4 *normals;
    const float4 *emTris;
    const double *emCdf;
    const double *emitterCdf;       // nEmitters + 1 (selection by sampling weight)
    const DevEmitter *emitters;
    const DevMaterial *materials;
    const double *roughTables;      // roughplastic: DR_ROUGH_TABLE_DOUBLES per table (include/drmlt_b200.h)
    const float4 *uvs;              // 2 per triangle, leaf order (placeholder when the scene has no texcoords)
    const float4 *texels;
    const DevTexture *textures;
    int nEmitters, nTris, nNodes;
    int rootIsLeaf;                  // degenerate scenes with <= 4 triangles
    DevCamera cam;
    float epsilon, shadowEpsilon;
};

// Host-side BVH build output (bvh_build.cpp)
struct BuiltBVH {
    std::vector<float4> nodes;       // 4 per node
    std::vector<uint32_t> order;     // leaf order -> caller's triangle index
    int rootIsLeaf = 0;
    int maxDepth = 1;                // inner-node levels (bounds the traversal stack)
};
void build_bvh(const float *positions, const uint32_t *indices, uint32_t nTris, BuiltBVH &out);
#define DR_NO_CHILD ((int) 0x80000000)       /* empty child slot of a 4-wide node */
void collapse_bvh4(BuiltBVH &bvh);           // BVH2 (4 float4 per node) -> BVH4 (8 float4 per node); bvh_build.cpp
// LBVH build + triangle packing on the current device (bvh_gpu.cu): the device-resident arrays of DevScene; false + tooDeep: fall back to build_bvh
struct GpuScene {
    float4 *nodes = nullptr, *tris = nullptr, *normals = nullptr;
    uint32_t *order = nullptr;
    uint32_t nNodes = 0;
    size_t normalsCount = 0;         // float4s in `normals` (0: no smooth triangle, a 16-byte placeholder is allocated)
    int stackBound = 0;
    void release();
};
bool build_scene_gpu(const dr_scene_desc *d, bool anySmooth, int stackLimit, GpuScene &out, bool *tooDeep);

struct dr_scene_t {
    int device = 0;
    DevScene dev;                    // pointers are device pointers
    std::vector<void *> allocations;
    int filmW = 0, filmH = 0;
    uint32_t nTris = 0, nNodes = 0;
    int bvhBuilder = 0, bvhDepth = 0;   // DR_SCENE_BVH_* that built the tree | its stack bound
    double bvhBuildMs = 0.0;
    unsigned typeMask = 0;           // BSDF models present (bit = dr_bsdf_type)
    std::atomic<int> cancel{0};       // set by dr_cancel from any thread, polled between graph replays
    size_t bytes = 0;
};

void dr_set_error(const char *fmt, ...);
