// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
//
// orc_scene.hpp: flattened scene, ray casting, emitter and sensor sampling.
// Restates: src/librender/skdtree.cpp:111-138 (closest hit, adaptive epsilon),
// include/mitsuba/render/triaccel.h:61-157 (Wald projection test),
// include/mitsuba/render/skdtree.h:343-426 (fillIntersectionRecord),
// src/librender/scene.cpp:879-904,1057-1087 (emitter sampling), src/emitters/area.cpp:98-200,
// src/librender/trimesh.cpp:429-440, src/librender/shape.cpp:102-126, src/libcore/triangle.cpp:24-60,
// include/mitsuba/core/pmf.h:60-200, src/sensors/perspective.cpp:126-428,
// src/textures/bitmap.cpp:432-455 + include/mitsuba/render/mipmap.h:503-596 + src/librender/texture.cpp:112-121 (bitmap lookup
// without ray differentials), src/librender/trimesh.cpp:708-760 (UV tangents).
// The acceleration structure is a plain SAH BVH: the reference's kd-tree only decides WHICH
// triangles are tested, the closest hit it returns is the same.
#pragma once
#include "orc_math.hpp"
#include "../include/drmlt_b200.h"
#include <stdexcept>

namespace orc {

// ------------------------------------------------------------------ DiscreteDistribution
// include/mitsuba/core/pmf.h:60-200
struct DiscreteDistribution {
    std::vector<Float> cdf;
    Float sum = 0, normalization = 0;
    bool normalized = false;
    DiscreteDistribution() { cdf.push_back(0.0); }
    void clear() { cdf.clear(); cdf.push_back(0.0); normalized = false; }
    void append(Float v) { cdf.push_back(cdf.back() + v); }
    size_t size() const { return cdf.size() - 1; }
    Float operator[](size_t i) const { return cdf[i + 1] - cdf[i]; }
    Float normalize() {
        sum = cdf.back();
        if (sum > 0) {
            normalization = 1.0 / sum;
            for (size_t i = 1; i < cdf.size(); ++i) cdf[i] *= normalization;
            cdf.back() = 1.0;
            normalized = true;
        } else {
            normalization = 0.0;
        }
        return sum;
    }
    size_t sample(Float v) const {
        auto entry = std::lower_bound(cdf.begin(), cdf.end(), v);
        size_t index = (size_t) std::min((ptrdiff_t) cdf.size() - 2,
                                         std::max((ptrdiff_t) 0, (ptrdiff_t) (entry - cdf.begin()) - 1));
        while ((*this)[index] == 0 && index < cdf.size() - 1) ++index;
        return index;
    }
    size_t sampleReuse(Float &v, Float &pdf) const {
        size_t index = sample(v);
        pdf = (*this)[index];
        v = (v - cdf[index]) / (cdf[index + 1] - cdf[index]);
        return index;
    }
    size_t sampleReuse(Float &v) const { Float pdf; return sampleReuse(v, pdf); }
};

// ------------------------------------------------------------------ geometry
struct Ray {
    Vec3 o, d;
    Float mint, maxt;
};

struct Intersection {
    Float t = INF;
    Vec3 p;
    Vec3 ng;          // geoFrame.n (flipped to agree with the shading normal)
    Frame sh;         // shFrame
    Vec2 uv;
    Vec3 wi;          // local incident direction
    int prim = -1;
    int material = -1;
    int emitter = -1;
    bool valid() const { return t != INF; }
    Vec3 toLocal(const Vec3 &v) const { return sh.toLocal(v); }
    Vec3 toWorld(const Vec3 &v) const { return sh.toWorld(v); }
};

struct TriAccel {   // include/mitsuba/render/triaccel.h:37-58
    int k;
    Float n_u, n_v, n_d, a_u, a_v, b_nu, b_nv, c_nu, c_nv;
};

struct BVHNode {
    Float lo[3], hi[3];
    int left, right;       // children (inner) ; left = -1 for leaf
    int first, count;      // leaf primitives (indices into order[])
};

struct Camera {
    Float toWorld[16];
    Float inv[9];           // inverse of the 3x3 block (the reference inverts the matrix: Transform::inverse, transform.h)
    Vec3 pos, dir;          // trafo(0), trafo((0,0,1))
    Float tanHalf, aspect, nearClip, farClip;
    Float resX, resY;       // crop size = the sensor's resolution (perspective.cpp:126-130)
    Float relOffX = 0, relOffY = 0, relSizeX = 1, relSizeY = 1;   // crop window as fractions of the film (:132-135)
    Float rectMinX, rectMaxX, rectMinY, rectMaxY;   // m_imageRect at z = 1 (:162-169)
    Float normalization;
    int filmW = 0, filmH = 0;   // film size given by dr_camera (a configuration may override it, see setWindow)
    // PerspectiveCameraImpl::configure (perspective.cpp:126-173) for a film / crop window (Film::Film, film.cpp:30-48)
    void setWindow(int fW, int fH, int cropX, int cropY, int cropW, int cropH) {
        aspect = (Float) fW / (Float) fH;
        resX = cropW; resY = cropH;
        relSizeX = (Float) cropW / (Float) fW; relSizeY = (Float) cropH / (Float) fH;
        relOffX = (Float) cropX / (Float) fW; relOffY = (Float) cropY / (Float) fH;
        Float x0 = (1 - 2 * relOffX) * tanHalf, x1 = (1 - 2 * (relOffX + relSizeX)) * tanHalf;
        Float y0 = (1 - 2 * relOffY) * tanHalf / aspect, y1 = (1 - 2 * (relOffY + relSizeY)) * tanHalf / aspect;
        rectMinX = std::min(x0, x1); rectMaxX = std::max(x0, x1);
        rectMinY = std::min(y0, y1); rectMaxY = std::max(y0, y1);
        normalization = 1.0 / ((rectMaxX - rectMinX) * (rectMaxY - rectMinY));   // :167-173
    }
    Vec3 xformDir(const Vec3 &v) const {
        return Vec3(toWorld[0] * v.x + toWorld[1] * v.y + toWorld[2] * v.z,
                    toWorld[4] * v.x + toWorld[5] * v.y + toWorld[6] * v.z,
                    toWorld[8] * v.x + toWorld[9] * v.y + toWorld[10] * v.z);
    }
    Vec3 invDir(const Vec3 &v) const {   // world -> camera direction
        return Vec3(inv[0] * v.x + inv[1] * v.y + inv[2] * v.z, inv[3] * v.x + inv[4] * v.y + inv[5] * v.z, inv[6] * v.x + inv[7] * v.y + inv[8] * v.z);
    }
    void invert() {
        const Float *a = toWorld;
        const Float c00 = a[5] * a[10] - a[6] * a[9], c01 = a[6] * a[8] - a[4] * a[10], c02 = a[4] * a[9] - a[5] * a[8];
        const Float id = 1.0 / (a[0] * c00 + a[1] * c01 + a[2] * c02);
        inv[0] = c00 * id; inv[1] = (a[2] * a[9] - a[1] * a[10]) * id; inv[2] = (a[1] * a[6] - a[2] * a[5]) * id;
        inv[3] = c01 * id; inv[4] = (a[0] * a[10] - a[2] * a[8]) * id; inv[5] = (a[2] * a[4] - a[0] * a[6]) * id;
        inv[6] = c02 * id; inv[7] = (a[1] * a[8] - a[0] * a[9]) * id; inv[8] = (a[0] * a[5] - a[1] * a[4]) * id;
    }
    // perspective.cpp:150-157 m_sampleToCamera applied to (sx, sy, 0), normalised
    Vec3 sampleToDir(Float sx, Float sy) const {
        Float fx = relOffX + sx * relSizeX, fy = relOffY + sy * relSizeY;   // sample over the crop window -> film fraction
        return normalize(Vec3((1 - 2 * fx) * tanHalf, (1 - 2 * fy) * tanHalf / aspect, 1.0));
    }
    // perspective.cpp:191-245
    Float importance(const Vec3 &d) const {
        Float cosTheta = d.z;
        if (cosTheta <= 0) return 0.0;
        Float invCosTheta = 1.0 / cosTheta;
        Float px = d.x * invCosTheta, py = d.y * invCosTheta;
        if (px < rectMinX || px > rectMaxX || py < rectMinY || py > rectMaxY) return 0.0;
        return normalization * invCosTheta * invCosTheta * invCosTheta;
    }
    // perspective.cpp:367-385 (dWorld need not be normalised)
    bool getSamplePosition(const Vec3 &dWorld, Vec2 &pos) const {
        Vec3 local = invDir(dWorld);
        if (local.z <= 0) return false;
        Float sx = (0.5 * (1 - local.x / (local.z * tanHalf)) - relOffX) / relSizeX;
        Float sy = (0.5 * (1 - local.y * aspect / (local.z * tanHalf)) - relOffY) / relSizeY;
        if (sx < 0 || sx > 1 || sy < 0 || sy > 1) return false;
        pos = Vec2(sx * resX, sy * resY);
        return true;
    }
};

struct EmitterRec {
    int firstTri, nTris;
    RGB radiance;
    Float area, invArea;
    DiscreteDistribution areaDistr;   // trimesh.cpp:405-420
    RGB power() const { return radiance * (PI * area); }   // area.cpp:205
};

namespace detail { inline void preparePlastic(dr_material &m); inline void prepareRoughPlastic(dr_material &m, const double *sceneTables); }   // orc_bsdf.hpp: constants SmoothPlastic::configure derives

// A bitmap texture: MIP level 0 of the reference's TMIPMap<Spectrum, Color3> (texels in Float) with its boundary conditions.
struct Texture {
    int w = 0, h = 0, wrapU = 0, wrapV = 0;
    bool nearest = false;
    Float scaleU = 1, scaleV = 1, offU = 0, offV = 0;
    std::vector<RGB> texels;
    static int modulo(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }   // math::modulo
    // mipmap.h:503-563 evalTexel
    RGB texel(int x, int y) const {
        if (x < 0 || x >= w) {
            switch (wrapU) {
                case DR_WRAP_REPEAT: x = modulo(x, w); break;
                case DR_WRAP_CLAMP: x = std::min(std::max(x, 0), w - 1); break;
                case DR_WRAP_MIRROR: x = modulo(x, 2 * w); if (x >= w) x = 2 * w - x - 1; break;
                case DR_WRAP_ZERO: return RGB(0.0f);
                default: return RGB(1.0f);
            }
        }
        if (y < 0 || y >= h) {
            switch (wrapV) {
                case DR_WRAP_REPEAT: y = modulo(y, h); break;
                case DR_WRAP_CLAMP: y = std::min(std::max(y, 0), h - 1); break;
                case DR_WRAP_MIRROR: y = modulo(y, 2 * h); if (y >= h) y = 2 * h - y - 1; break;
                case DR_WRAP_ZERO: return RGB(0.0f);
                default: return RGB(1.0f);
            }
        }
        return texels[(size_t) y * w + x];
    }
    // Texture2D::eval (texture.cpp:112-121) -> BitmapTexture::eval(uv) (bitmap.cpp:432-455) -> evalBilinear / evalBox (mipmap.h:566-596)
    RGB eval(const Vec2 &itsUv) const {
        const Vec2 uv(itsUv.x * scaleU + offU, itsUv.y * scaleV + offV);
        if (nearest) return texel((int) std::floor(uv.x * w), (int) std::floor(uv.y * h));
        if (!std::isfinite(uv.x) || !std::isfinite(uv.y)) return RGB(0.0f);
        Float u = uv.x * w - 0.5f, v = uv.y * h - 0.5f;
        int xPos = (int) std::floor(u), yPos = (int) std::floor(v);
        Float dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
        return texel(xPos, yPos) * dx2 * dy2 + texel(xPos, yPos + 1) * dx2 * dy1
             + texel(xPos + 1, yPos) * dx1 * dy2 + texel(xPos + 1, yPos + 1) * dx1 * dy1;
    }
};

struct Scene {
    std::vector<Vec3> P, N;
    std::vector<Vec2> UV;              // vertex texture coordinates (empty: its.uv is the barycentric pair)
    std::vector<Texture> textures;
    std::vector<uint32_t> idx, triMat, triFlags;
    std::vector<int32_t> triEmitter;
    std::vector<dr_material> mats;
    bool hasRoughDielectric = false;   // pssmlt_utils.h:35-45
    std::vector<EmitterRec> emitters;
    DiscreteDistribution emitterPDF;   // scene.cpp:380-383 (sampling weights)
    std::vector<TriAccel> accel;
    std::vector<BVHNode> nodes;
    std::vector<int> order;
    Camera cam;
    Float epsilon = 1e-7, shadowEpsilon = 1e-5;   // constants.h:25-27 (double build)
    mutable uint64_t dummy = 0;

    void load(const dr_scene_desc &d);
    void buildBVH();
    bool rayIntersect(const Ray &ray, Intersection &its, uint64_t *rayCounter = nullptr) const;
    bool rayIntersectShadow(const Ray &ray, uint64_t *rayCounter = nullptr) const;
    bool traverse(const Ray &ray, Float mint, Float maxt, bool shadow, Float &tOut, Float &uOut, Float &vOut, int &prim) const;
    void fillIntersection(const Ray &ray, Float t, Float u, Float v, int prim, Intersection &its) const;
    Float adaptiveMint(const Ray &ray) const {
        Float m = ray.mint;
        if (m == epsilon)   // skdtree.cpp:126-129
            m *= std::max(std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z)), epsilon);
        return m;
    }
    Ray makeRay(const Vec3 &o, const Vec3 &d) const { Ray r; r.o = o; r.d = d; r.mint = epsilon; r.maxt = INF; return r; }
};

inline int triLoad(TriAccel &ta, const Vec3 &A, const Vec3 &B, const Vec3 &C) {
    static const int waldModulo[4] = { 1, 2, 0, 1 };
    Vec3 b = C - A, c = B - A, N = cross(c, b);
    int k = 0;
    for (int j = 0; j < 3; j++)
        if (std::abs(N[j]) > std::abs(N[k])) k = j;
    int u = waldModulo[k], v = waldModulo[k + 1];
    const Float n_k = N[k], denom = b[u] * c[v] - b[v] * c[u];
    if (denom == 0) { ta.k = 3; return 1; }
    ta.k = k;
    ta.n_u = N[u] / n_k; ta.n_v = N[v] / n_k; ta.n_d = dot(A, N) / n_k;
    ta.b_nu = b[u] / denom; ta.b_nv = -b[v] / denom;
    ta.a_u = A[u]; ta.a_v = A[v];
    ta.c_nu = c[v] / denom; ta.c_nv = -c[u] / denom;
    return 0;
}

inline bool triIntersect(const TriAccel &ta, const Ray &ray, Float mint, Float maxt, Float &u, Float &v, Float &t) {
    Float o_u, o_v, o_k, d_u, d_v, d_k;
    switch (ta.k) {
        case 0: o_u = ray.o.y; o_v = ray.o.z; o_k = ray.o.x; d_u = ray.d.y; d_v = ray.d.z; d_k = ray.d.x; break;
        case 1: o_u = ray.o.z; o_v = ray.o.x; o_k = ray.o.y; d_u = ray.d.z; d_v = ray.d.x; d_k = ray.d.y; break;
        case 2: o_u = ray.o.x; o_v = ray.o.y; o_k = ray.o.z; d_u = ray.d.x; d_v = ray.d.y; d_k = ray.d.z; break;
        default: return false;
    }
    t = (ta.n_d - o_u * ta.n_u - o_v * ta.n_v - o_k) / (d_u * ta.n_u + d_v * ta.n_v + d_k);
    if (!(t >= mint && t <= maxt)) return false;   // NaN rejected as in "t < mint || t > maxt" + later tests
    const Float hu = o_u + t * d_u - ta.a_u;
    const Float hv = o_v + t * d_v - ta.a_v;
    u = hv * ta.b_nu + hu * ta.b_nv;
    v = hu * ta.c_nu + hv * ta.c_nv;
    return u >= 0 && v >= 0 && u + v <= 1.0;
}

inline void Scene::load(const dr_scene_desc &d) {
    P.resize(d.n_vertices);
    for (uint32_t i = 0; i < d.n_vertices; ++i)
        P[i] = Vec3(d.positions[3 * i], d.positions[3 * i + 1], d.positions[3 * i + 2]);
    if (d.normals) {
        N.resize(d.n_vertices);
        for (uint32_t i = 0; i < d.n_vertices; ++i)
            N[i] = Vec3(d.normals[3 * i], d.normals[3 * i + 1], d.normals[3 * i + 2]);
    }
    if (d.texcoords) {
        UV.resize(d.n_vertices);
        for (uint32_t i = 0; i < d.n_vertices; ++i) UV[i] = Vec2(d.texcoords[2 * i], d.texcoords[2 * i + 1]);
    }
    textures.resize(d.n_textures);
    for (uint32_t t = 0; t < d.n_textures; ++t) {
        const dr_texture &tx = d.textures[t];
        Texture &o = textures[t];
        o.w = (int) tx.width; o.h = (int) tx.height; o.wrapU = (int) tx.wrap_u; o.wrapV = (int) tx.wrap_v; o.nearest = tx.nearest != 0;
        o.scaleU = tx.uv_scale[0]; o.scaleV = tx.uv_scale[1]; o.offU = tx.uv_offset[0]; o.offV = tx.uv_offset[1];
        o.texels.resize((size_t) o.w * o.h);
        for (size_t i = 0; i < o.texels.size(); ++i) o.texels[i] = RGB(tx.texels[3 * i], tx.texels[3 * i + 1], tx.texels[3 * i + 2]);
    }
    idx.assign(d.indices, d.indices + 3 * (size_t) d.n_triangles);
    triMat.assign(d.tri_material, d.tri_material + d.n_triangles);
    triEmitter.assign(d.tri_emitter, d.tri_emitter + d.n_triangles);
    if (d.tri_flags) triFlags.assign(d.tri_flags, d.tri_flags + d.n_triangles);
    else triFlags.assign(d.n_triangles, 0);
    mats.assign(d.materials, d.materials + d.n_materials);
    for (dr_material &m : mats) detail::preparePlastic(m);
    for (dr_material &m : mats)
        if (m.type == DR_BSDF_ROUGHPLASTIC) {
            if (!d.rough_tables || m.table >= d.n_rough_tables) throw std::runtime_error("roughplastic material without a rough-transmittance table");
            detail::prepareRoughPlastic(m, d.rough_tables);
        }
    for (uint32_t i = 0; i < d.n_triangles; ++i) hasRoughDielectric |= mats[d.tri_material[i]].type == DR_BSDF_ROUGHDIELECTRIC;
    emitters.resize(d.n_emitters);
    emitterPDF.clear();
    for (uint32_t e = 0; e < d.n_emitters; ++e) {
        EmitterRec &er = emitters[e];
        er.firstTri = d.emitters[e].first_tri;
        er.nTris = d.emitters[e].n_tris;
        er.radiance = RGB(d.emitters[e].radiance[0], d.emitters[e].radiance[1], d.emitters[e].radiance[2]);
        for (int i = 0; i < er.nTris; ++i) {
            int tri = er.firstTri + i;
            const Vec3 &p0 = P[idx[3 * tri]], &p1 = P[idx[3 * tri + 1]], &p2 = P[idx[3 * tri + 2]];
            er.areaDistr.append(0.5 * length(cross(p1 - p0, p2 - p0)));   // triangle.cpp:62-68
        }
        er.area = er.areaDistr.normalize();
        er.invArea = 1.0 / er.area;
        emitterPDF.append(d.emitters[e].sampling_weight);
    }
    if (d.n_emitters) emitterPDF.normalize();
    accel.resize(d.n_triangles);
    for (uint32_t i = 0; i < d.n_triangles; ++i)
        triLoad(accel[i], P[idx[3 * i]], P[idx[3 * i + 1]], P[idx[3 * i + 2]]);

    const dr_camera &c = d.camera;
    for (int i = 0; i < 16; ++i) cam.toWorld[i] = c.to_world[i];
    cam.invert();
    cam.pos = Vec3(c.to_world[3], c.to_world[7], c.to_world[11]);
    cam.dir = cam.xformDir(Vec3(0, 0, 1));
    cam.tanHalf = std::tan(0.5 * (Float) c.xfov_deg * PI / 180.0);
    cam.nearClip = c.near_clip; cam.farClip = c.far_clip;
    cam.filmW = c.film_width; cam.filmH = c.film_height;
    cam.setWindow(c.film_width, c.film_height, 0, 0, c.film_width, c.film_height);
    buildBVH();
}

// Binned-SAH BVH (16 bins, leaves <= 4 primitives).
inline void Scene::buildBVH() {
    const int n = (int) accel.size();
    order.resize(n);
    std::vector<Vec3> lo(n), hi(n), cen(n);
    for (int i = 0; i < n; ++i) {
        order[i] = i;
        const Vec3 &a = P[idx[3 * i]], &b = P[idx[3 * i + 1]], &c = P[idx[3 * i + 2]];
        lo[i] = Vec3(std::min(a.x, std::min(b.x, c.x)), std::min(a.y, std::min(b.y, c.y)), std::min(a.z, std::min(b.z, c.z)));
        hi[i] = Vec3(std::max(a.x, std::max(b.x, c.x)), std::max(a.y, std::max(b.y, c.y)), std::max(a.z, std::max(b.z, c.z)));
        cen[i] = (lo[i] + hi[i]) * 0.5;
    }
    nodes.clear();
    nodes.reserve(2 * n / 2 + 16);
    struct Task { int node, first, count; };
    std::vector<Task> stack;
    nodes.push_back(BVHNode());
    stack.push_back({ 0, 0, n });
    auto area = [](const Float *l, const Float *h) {
        Float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
        return 2 * (dx * dy + dy * dz + dz * dx);
    };
    while (!stack.empty()) {
        Task tk = stack.back(); stack.pop_back();
        BVHNode nd;
        for (int a = 0; a < 3; ++a) { nd.lo[a] = INF; nd.hi[a] = -INF; }
        Float clo[3] = { INF, INF, INF }, chi[3] = { -INF, -INF, -INF };
        for (int i = tk.first; i < tk.first + tk.count; ++i) {
            int p = order[i];
            for (int a = 0; a < 3; ++a) {
                nd.lo[a] = std::min(nd.lo[a], lo[p][a]); nd.hi[a] = std::max(nd.hi[a], hi[p][a]);
                clo[a] = std::min(clo[a], cen[p][a]); chi[a] = std::max(chi[a], cen[p][a]);
            }
        }
        nd.left = -1; nd.right = -1; nd.first = tk.first; nd.count = tk.count;
        if (tk.count > 4) {
            const int NB = 16;
            int bestAxis = -1, bestSplit = -1; Float bestCost = INF;
            for (int a = 0; a < 3; ++a) {
                Float ext = chi[a] - clo[a];
                if (!(ext > 0)) continue;
                int cnt[NB] = { 0 }; Float blo[NB][3], bhi[NB][3];
                for (int b = 0; b < NB; ++b) for (int c = 0; c < 3; ++c) { blo[b][c] = INF; bhi[b][c] = -INF; }
                for (int i = tk.first; i < tk.first + tk.count; ++i) {
                    int p = order[i];
                    int b = std::min(NB - 1, (int) (NB * (cen[p][a] - clo[a]) / ext));
                    cnt[b]++;
                    for (int c = 0; c < 3; ++c) { blo[b][c] = std::min(blo[b][c], lo[p][c]); bhi[b][c] = std::max(bhi[b][c], hi[p][c]); }
                }
                Float rArea[NB]; int rCnt[NB];
                Float l[3] = { INF, INF, INF }, h[3] = { -INF, -INF, -INF }; int c2 = 0;
                for (int b = NB - 1; b > 0; --b) {
                    for (int c = 0; c < 3; ++c) { l[c] = std::min(l[c], blo[b][c]); h[c] = std::max(h[c], bhi[b][c]); }
                    c2 += cnt[b]; rCnt[b] = c2; rArea[b] = c2 ? area(l, h) : 0;
                }
                Float l2[3] = { INF, INF, INF }, h2[3] = { -INF, -INF, -INF }; int c1 = 0;
                for (int b = 0; b < NB - 1; ++b) {
                    for (int c = 0; c < 3; ++c) { l2[c] = std::min(l2[c], blo[b][c]); h2[c] = std::max(h2[c], bhi[b][c]); }
                    c1 += cnt[b];
                    if (c1 == 0 || rCnt[b + 1] == 0) continue;
                    Float cost = c1 * area(l2, h2) + rCnt[b + 1] * rArea[b + 1];
                    if (cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = b; }
                }
            }
            int mid;
            if (bestAxis >= 0) {
                Float ext = chi[bestAxis] - clo[bestAxis];
                int *beg = order.data() + tk.first, *end = beg + tk.count;
                int *m = std::partition(beg, end, [&](int p) {
                    int b = std::min(NB - 1, (int) (NB * (cen[p][bestAxis] - clo[bestAxis]) / ext));
                    return b <= bestSplit;
                });
                mid = (int) (m - order.data());
            } else {
                mid = tk.first + tk.count / 2;
            }
            if (mid == tk.first || mid == tk.first + tk.count) mid = tk.first + tk.count / 2;
            nd.left = (int) nodes.size(); nd.right = nd.left + 1;
            nodes.push_back(BVHNode()); nodes.push_back(BVHNode());
            stack.push_back({ nd.left, tk.first, mid - tk.first });
            stack.push_back({ nd.right, mid, tk.first + tk.count - mid });
        }
        nodes[tk.node] = nd;
    }
}

inline bool Scene::traverse(const Ray &ray, Float mint, Float maxt, bool shadow,
                            Float &tOut, Float &uOut, Float &vOut, int &prim) const {
    Float inv[3] = { 1.0 / ray.d.x, 1.0 / ray.d.y, 1.0 / ray.d.z };
    Float o[3] = { ray.o.x, ray.o.y, ray.o.z };
    int stack[128]; int sp = 0;
    stack[sp++] = 0;
    bool found = false;
    while (sp) {
        const BVHNode &nd = nodes[stack[--sp]];
        Float t0 = mint, t1 = maxt;
        bool miss = false;
        for (int a = 0; a < 3; ++a) {
            Float ta = (nd.lo[a] - o[a]) * inv[a], tb = (nd.hi[a] - o[a]) * inv[a];
            if (ta > tb) std::swap(ta, tb);
            // conservative slab test; NaN (0*inf) treated as "inside"
            if (ta > t0) t0 = ta;
            if (tb < t1) t1 = tb;
            if (t0 > t1 * (1 + 1e-12) + 1e-300) { miss = true; break; }
        }
        if (miss) continue;
        if (nd.left < 0) {
            for (int i = nd.first; i < nd.first + nd.count; ++i) {
                int p = order[i];
                Float u, v, t;
                if (triIntersect(accel[p], ray, mint, maxt, u, v, t)) {
                    if (shadow) { prim = p; return true; }
                    maxt = t; tOut = t; uOut = u; vOut = v; prim = p; found = true;
                }
            }
        } else {
            stack[sp++] = nd.left;
            stack[sp++] = nd.right;
        }
    }
    return found;
}

// skdtree.h:343-426
inline void Scene::fillIntersection(const Ray &ray, Float t, Float u, Float v, int prim, Intersection &its) const {
    const Vec3 b(1 - u - v, u, v);
    const uint32_t i0 = idx[3 * prim], i1 = idx[3 * prim + 1], i2 = idx[3 * prim + 2];
    const Vec3 &p0 = P[i0], &p1 = P[i1], &p2 = P[i2];
    its.t = t;
    its.p = p0 * b.x + p1 * b.y + p2 * b.z;
    Vec3 side1 = p1 - p0, side2 = p2 - p0;
    Vec3 faceNormal = cross(side1, side2);
    Float len = length(faceNormal);
    if (len != 0) faceNormal /= len;
    Vec3 shN;
    if ((triFlags[prim] & DR_TRI_SMOOTH) && !N.empty()) {
        shN = normalize(N[i0] * b.x + N[i1] * b.y + N[i2] * b.z);
        if (dot(faceNormal, shN) < 0) faceNormal = -faceNormal;
    } else {
        shN = faceNormal;
    }
    its.ng = faceNormal;
    Vec3 dpdu = side1;
    if (!UV.empty() && !(triFlags[prim] & DR_TRI_NO_TEXCOORDS)) {
        const Vec2 &t0 = UV[i0], &t1 = UV[i1], &t2 = UV[i2];
        its.uv = Vec2(t0.x * b.x + t1.x * b.y + t2.x * b.z, t0.y * b.x + t1.y * b.y + t2.y * b.z);
        if (triFlags[prim] & DR_TRI_UV_TANGENTS) {   // trimesh.cpp:741-759 (computed per mesh there, per hit here)
            const Vec2 dUV1(t1.x - t0.x, t1.y - t0.y), dUV2(t2.x - t0.x, t2.y - t0.y);
            const Float determinant = dUV1.x * dUV2.y - dUV1.y * dUV2.x;
            if (determinant == 0) {
                Vec3 n = cross(side1, side2), dpdv;
                coordinateSystem(n / length(n), dpdu, dpdv);
            } else {
                const Float invDet = 1.0f / determinant;
                dpdu = (side1 * dUV2.y - side2 * dUV1.y) * invDet;
            }
        }
    } else {
        its.uv = Vec2(b.y, b.z);
    }
    its.prim = prim;
    its.material = (int) triMat[prim];
    its.emitter = triEmitter[prim];
    computeShadingFrame(shN, dpdu, its.sh);
    its.wi = its.toLocal(-ray.d);
}

inline bool Scene::rayIntersect(const Ray &ray, Intersection &its, uint64_t *rayCounter) const {
    if (rayCounter) ++*rayCounter;
    its.t = INF;
    Float mint = adaptiveMint(ray), maxt = ray.maxt;
    if (!(maxt > mint)) return false;
    Float t, u, v; int prim;
    if (traverse(ray, mint, maxt, false, t, u, v, prim)) {
        fillIntersection(ray, t, u, v, prim, its);
        return true;
    }
    return false;
}

inline bool Scene::rayIntersectShadow(const Ray &ray, uint64_t *rayCounter) const {
    if (rayCounter) ++*rayCounter;
    Float mint = adaptiveMint(ray), maxt = ray.maxt;
    if (!(maxt > mint)) return false;
    Float t, u, v; int prim;
    return traverse(ray, mint, maxt, true, t, u, v, prim);
}

} // namespace orc
